"""The host-side mirror of models.mocodad.MoCoDAD: constructor surface, state_dict layout, error
behaviour.  CPU only -- forward needs the GPU and says so."""
import argparse

import pytest
import torch

from mocodad_b200 import MoCoDAD, state_dict_spec
from oracle import synth

BASE = dict(seg_len=6, num_coords=2, headless=False, kp18_format=False, embedding_dim=16, dropout=0.0,
            conditioning_strategy="inject", conditioning_indices=[0, 1, 2], conditioning_architecture="AE",
            h_dim=32, latent_dim=16, channels=[32, 16, 32], opt_lr=1e-4, loss_fn="smooth_l1", rec_weight=0.01,
            noise_steps=10, aggregation_strategy="best", n_generated_samples=50, model_return_value="loss",
            gt_path="", split="test", use_hr=False, ckpt_dir="", save_tensors=False, num_transform=5,
            pad_size=12, filter_kernel_size=30, frames_shift=6, dataset_choice="HR-Avenue", seed=999)


def make(**kw):
    cfg = dict(BASE)
    cfg.update(kw)
    return MoCoDAD(argparse.Namespace(**cfg))


def test_state_dict_layout_matches_reference_checkpoint():
    m = make()
    spec = synth.state_dict_spec(T=3, T_cond=3)  # asserted equal to the real module by oracle/make_golden.py
    sd = m.state_dict()
    assert list(sd.keys()) == list(spec.keys())
    assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in spec)
    assert len(sd) == 335
    assert sum(v.numel() for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k) == 142294
    assert dict(state_dict_spec(T=3)) == dict(spec)
    m.load_state_dict(synth.synth_state_dict(spec, seed=0), strict=True)
    assert (m.n_frames_condition, m.n_frames_corrupt, m.input_n_frames) == (3, 3, 3)


def test_variants_of_the_constructor():
    m = make(seg_len=27)
    assert m.input_n_frames == 24 and len(m.state_dict()) == 335
    m = make(conditioning_indices=2)  # integer form: first seg_len // 2 frames condition
    assert m.n_frames_condition == 3
    m = make(conditioning_indices=[3, 4, 5])
    assert m._cond_first is False
    m = make(conditioning_strategy="no_condition", seg_len=3)
    assert m.n_frames_condition == 0 and not any(k.startswith("condition_encoder") for k in m.state_dict())
    m = make(conditioning_architecture="E")
    assert not any("decoder" in k for k in m.state_dict())
    assert torch.equal(m._alpha_, 1 - m._beta_)


def test_out_of_scope_configurations_raise_like_the_reference():
    with pytest.raises(NotImplementedError):
        make(conditioning_strategy="concat")
    with pytest.raises(NotImplementedError):
        make(conditioning_architecture="E_unet")
    with pytest.raises(KeyError):
        make(conditioning_strategy="bogus")
    with pytest.raises(AssertionError):
        make(conditioning_indices=[1, 2, 3])
    with pytest.raises(NotImplementedError):
        make().training_step(None, 0)


def test_forward_refuses_to_run_without_cuda():
    m = make()
    batch = synth.synth_batch(4, 6)
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.forward(batch)


def test_trajectory_entry_points_validate_before_touching_the_gpu(tmp_path):
    with pytest.raises(NotImplementedError):
        make(aggregation_strategy="mean").score_trajectories(str(tmp_path), [640, 360])
    with pytest.raises(ValueError, match="data_dir and vid_res"):
        make().score_trajectories()
    m = make(data_dir=str(tmp_path), vid_res=[640, 360])          # the YAML keys the reference passes to its dataset
    assert m.data_dir == str(tmp_path) and m.vid_res == [640, 360]
    with pytest.raises(FileNotFoundError):
        m.score_trajectories()                                     # no {data_dir}/testing/trajectories


def test_product_package_never_touches_the_oracle_or_a_cpu_path():
    """The oracle is test infrastructure: nothing under mocodad_b200/ may import it, and the CUDA library is the only
    implementation (a missing library or a CPU device is an error, never a fallback)."""
    import ast
    import os
    import mocodad_b200
    pkg = os.path.dirname(mocodad_b200.__file__)
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), f"{fn} imports the oracle"
    from mocodad_b200 import ScoringEngine
    with pytest.raises(RuntimeError, match="no CPU path"):
        ScoringEngine(seg_len=6, n_frames_cond=3, noise_steps=10, device="cpu")


def test_fit_trajectory_scaler_validates_its_inputs(tmp_path):
    with pytest.raises(ValueError, match="data_dir and vid_res"):
        make().fit_trajectory_scaler()
    with pytest.raises(FileNotFoundError):
        make(data_dir=str(tmp_path), vid_res=[640, 360]).fit_trajectory_scaler()   # no {data_dir}/training/trajectories
