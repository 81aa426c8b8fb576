"""SURVEY.md 8 row f4 (the latent variant, models/mocodad_latent.py): the CPU restatement oracle/latent_port.py against the
fixture the UNMODIFIED MoCoDADlatent produced (oracle/make_latent_golden.py asserted bit-equality where it was generated; a
small fp32 tolerance here allows for another CPU / BLAS build), and the host-side surface of the product's MoCoDADlatent
(state_dict layout, constructor contract, no CPU fallback).  The CUDA parity tests are in tests/test_latent_gpu.py."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import latent_port, ref_port, synth

TOL = dict(rtol=0, atol=5e-6)


def _setup(golden):
    g = golden("latent_T3")
    seg_len, N, G, B, latent = (int(v) for v in g["meta"][:5])
    hidden = [int(v) for v in g["meta"][5:]]
    spec = OrderedDict((str(n), tuple(int(d) for d in s.split(",")) if s else ()) for n, s in zip(g["spec_names"], g["spec_shapes"]))
    sd = synth.synth_state_dict(spec, seed=0)
    batch = synth.synth_batch(B, seg_len, seed=1)
    return g, sd, batch, torch.from_numpy(g["noise"]), (seg_len, N, G, B, latent, hidden)


def test_state_dict_layout_of_the_latent_module(golden):
    g, sd, *_ = _setup(golden)
    assert len(sd) == 292
    assert sd["model.to_time_dim.weight"].shape == (64, 64 * 3 * 10)          # STSE_Unet out layer: C*T*V(10) -> latent
    assert sd["denoiser.net.0.0.weight"].shape == (64, 64) and sd["denoiser.net.3.weight"].shape == (64, 128)
    assert sd["denoiser.cond_layers.1.weight"].shape == (128, 16)
    assert not any(k.startswith("model.st_gcnnsu") or k.startswith("model.up") for k in sd)   # down half only


def test_latent_code_and_denoiser_call(golden):
    g, sd, batch, noise, (seg_len, N, G, B, latent, hidden) = _setup(golden)
    with torch.no_grad():
        cond, corrupt = ref_port.select_frames(batch[0], (0, 1, 2))
        emb = ref_port.cond_encode(sd, cond)
        taps = {}
        code = latent_port.latent_encode(sd, corrupt, emb, taps=taps)
        eps = latent_port.denoiser_forward(sd, noise[0, 0], torch.full((B,), 7, dtype=torch.long), emb, len(hidden))
    np.testing.assert_allclose(code.numpy(), g["latent_code"], **TOL)
    np.testing.assert_allclose(taps["st_gcnnsd3.1"].numpy(), g["tap_sd3_1"], **TOL)
    np.testing.assert_allclose(eps.numpy(), g["eps_t7"], **TOL)
    assert code.shape == (B, latent) and eps.shape == (B, hidden[-1])


@pytest.mark.parametrize("strategy", ["best", "mean", "median"])
def test_latent_reverse_diffusion_losses(golden, strategy):
    g, sd, batch, noise, (seg_len, N, G, B, latent, hidden) = _setup(golden)
    with torch.no_grad():
        loss, sel, _ = latent_port.latent_reverse_diffusion(sd, batch[0], noise_steps=N, n_generated_samples=G, n_layers=len(hidden),
                                                            noise=noise, strategy=strategy)
    np.testing.assert_allclose(loss.numpy(), g["loss_" + strategy], rtol=0, atol=2e-5)
    if strategy == "best":
        np.testing.assert_allclose(sel.numpy(), g["latent_sel"], rtol=0, atol=1e-4)


def _latent_args(**over):
    import argparse
    from test_module import BASE
    cfg = dict(BASE, diffusion_on_latent=True, stage="diffusion", latent_embedding_dim=64, hidden_sizes=[64, 128, 128, 64],
               pretrained_model_ckpt_path="", n_generated_samples=3)
    cfg.update(over)
    return argparse.Namespace(**cfg)


def test_product_module_has_the_reference_state_dict_layout(golden):
    """MoCoDADlatent(args).state_dict() == the unmodified reference module's, name by name and shape by shape."""
    from mocodad_b200 import MoCoDADlatent
    g = golden("latent_T3")
    m = MoCoDADlatent(_latent_args())
    sd = m.state_dict()
    assert list(sd.keys()) == [str(n) for n in g["spec_names"]]
    assert [",".join(map(str, v.shape)) for v in sd.values()] == [str(x) for x in g["spec_shapes"]]
    m.load_state_dict(synth.synth_state_dict(OrderedDict((k, tuple(v.shape)) for k, v in sd.items()), seed=0), strict=True)


def test_product_module_contract_without_a_gpu(tmp_path):
    from mocodad_b200 import MoCoDADlatent
    with pytest.raises(NotImplementedError, match="pretrain"):
        MoCoDADlatent(_latent_args(stage="pretrain"))
    with pytest.raises(ValueError):
        MoCoDADlatent(_latent_args(stage="nope"))
    with pytest.raises(ValueError, match="hidden_sizes"):
        MoCoDADlatent(_latent_args(hidden_sizes=[64, 128, 32]))
    with pytest.raises(AttributeError):          # a missing YAML key is an AttributeError, as upstream (mocodad_latent.py:24-27)
        ns = _latent_args()
        del ns.hidden_sizes
        MoCoDADlatent(ns)
    # the pretraining checkpoint is loaded with strict=False at construction (mocodad_latent.py:222-227)
    m0 = MoCoDADlatent(_latent_args())
    part = {k: torch.full_like(v, 0.5) for k, v in m0.state_dict().items() if k.startswith("model.st_gcnnsd1.0.tcn.0")}
    ck = tmp_path / "pretrain.ckpt"
    torch.save({"state_dict": part}, ck)
    m1 = MoCoDADlatent(_latent_args(pretrained_model_ckpt_path=str(ck)))
    assert all(torch.equal(m1.state_dict()[k], v) for k, v in part.items())
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):   # no CPU fallback
            m1.forward(synth.synth_batch(4, 6, seed=1))
