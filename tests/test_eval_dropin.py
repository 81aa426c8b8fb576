"""eval_MoCoDAD.py on the B200 path, with not a line of the script changed (SURVEY.md 8 row b; BASELINE.json configs[0], [3]).

Fixtures tests/golden/eval_*.npz come from the UNMODIFIED reference script run with the reference's own models / utils on a
synthetic dataset in the reference's on-disk format (oracle/make_eval_golden.py): shipped YAML values, trajectory rows, ground
truth (+ HR-UBnormal masks), scaler, and the per-window losses / AUC it produced.  Here the same YAML, the same files and the
same checkpoint go through `python -m mocodad_b200.dropin <script> -c cfg.yaml`, which resolves the script's imports
(`models.mocodad`, `models.mocodad_latent`, `utils.argparser`, `utils.dataset`, `pytorch_lightning`) to this repository.
  gpu tests       the whole epoch on the GPU: losses within 1e-4, |AUC difference| <= 1e-3 (0.1 percentage points)
  non-gpu tests   the saved-tensor branch of the script (no GPU needed) -- with the reference's real eval_MoCoDAD.py when the
                  checkout is present -- and the loud failure of the scoring branch on a machine without CUDA."""
import os
import pickle
import sys
from collections import OrderedDict

import numpy as np
import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_SCRIPT = os.path.join(os.environ.get("MOCODAD_REFERENCE", "/root/reference"), "eval_MoCoDAD.py")
CALLER = os.path.join(ROOT, "tests", "eval_caller.py")
CASES = ["avenue_N2", "ubnormal_hr", "ubnormal_latent"]
NOISE_SEED = 7000


def _materialise(g, root, extra_yaml=None):
    """The reference's on-disk layout under `root` (the YAMLs use ./data/... and ./checkpoints relative paths)."""
    from sklearn.preprocessing import RobustScaler
    from mocodad_b200 import synthetic as synth
    cfg = yaml.safe_load(str(g["yaml"]))
    cfg.update(extra_yaml or {})
    row0 = 0
    for k, n in enumerate(g["lengths"]):
        scene, clip, person = (int(v) for v in g["ids"][k])
        folder = os.path.join(root, cfg["data_dir"], "testing", "trajectories", f"{scene:02d}-{clip:04d}")
        os.makedirs(folder, exist_ok=True)
        rows = np.concatenate([g["frames_rows"][row0:row0 + n, None].astype(np.float64), g["coords"][row0:row0 + n].astype(np.float64)], axis=1)
        np.savetxt(os.path.join(folder, f"{person:04d}.csv"), rows, delimiter=",", fmt=["%d"] + ["%.2f"] * 34)
        row0 += n
    os.makedirs(os.path.join(root, cfg["test_path"]))
    for scene, clip in g["gt_keys"]:
        np.save(os.path.join(root, cfg["test_path"], f"{scene:02d}_{clip:04d}.npy"), g[f"gt_{scene}_{clip}"])
    for scene, clip in g["hr_keys"]:
        d = os.path.join(root, "data", "UBnormal", "hr_bool_masks", "testing", "test_frame_mask")
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, f"{scene}_{clip}.npy"), g[f"hr_{scene}_{clip}"])
    ckpt_dir = os.path.join(root, cfg["exp_dir"], cfg["dataset_choice"], cfg["dir_name"])
    os.makedirs(ckpt_dir)
    sk = RobustScaler(quantile_range=(10.0, 90.0))
    sk.center_, sk.scale_ = g["center"], g["scale"]
    with open(os.path.join(ckpt_dir, "local_robust.pickle"), "wb") as fh:
        pickle.dump(sk, fh)
    latent = "diffusion_on_latent" in cfg
    spec = synth.state_dict_spec(T=3, T_cond=3, latent_embedding_dim=cfg.get("latent_embedding_dim", 0) if latent else 0,
                                 hidden_sizes=cfg.get("hidden_sizes", ()) if latent else ())
    assert len(spec) == int(g["n_state_dict"])
    torch.save({"state_dict": synth.synth_state_dict(spec, seed=0)}, os.path.join(ckpt_dir, cfg["load_ckpt"]))
    if latent:
        os.makedirs(os.path.dirname(os.path.join(root, cfg["pretrained_model_ckpt_path"])))
        torch.save({"state_dict": {}}, os.path.join(root, cfg["pretrained_model_ckpt_path"]))
    with open(os.path.join(root, "cfg.yaml"), "w") as fh:
        yaml.safe_dump(cfg, fh)
    return cfg, ckpt_dir


def _run_launcher(script, monkeypatch, root):
    """`python -m mocodad_b200.dropin <script> -c cfg.yaml`, in process (so that the script's globals can be inspected)."""
    from mocodad_b200.dropin.__main__ import main
    monkeypatch.chdir(root)
    monkeypatch.setattr(sys, "path", list(sys.path))
    monkeypatch.setattr(sys, "argv", list(sys.argv))
    real_listdir = os.listdir
    monkeypatch.setattr(os, "listdir", lambda p=".": sorted(real_listdir(p)))     # dataset order independent of the file system
    saved = {k: v for k, v in sys.modules.items() if k in ("models", "utils", "pytorch_lightning") or k.startswith(("models.", "utils."))}
    try:
        return main([script, "-c", "cfg.yaml"])
    finally:
        for k in [k for k in sys.modules if k in ("models", "utils") or k.startswith(("models.", "utils."))]:
            del sys.modules[k]
        sys.modules.update(saved)


def _noise_feed():
    state = {"n": 0}
    real_randn = torch.randn

    def feed(*shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        gen = torch.Generator().manual_seed(NOISE_SEED + state["n"])
        state["n"] += 1
        return real_randn(*shape, generator=gen).to(kw.get("device", "cpu") or "cpu")
    return feed, state


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_eval_script_flow_on_the_gpu_matches_the_unmodified_reference_run(name, tmp_path, monkeypatch):
    g = np.load(os.path.join(GOLDEN, f"eval_{name}.npz"))
    # the shipped YAML + the one optional knob of this implementation: draw the noise with torch.randn in the reference's order
    cfg, ckpt_dir = _materialise(g, str(tmp_path), {"b200_rng": "torch"})
    assert cfg["accelerator"] == "gpu" and cfg["load_tensors"] is False
    feed, state = _noise_feed()
    monkeypatch.setattr(torch, "randn", feed)
    script = REF_SCRIPT if os.path.exists(REF_SCRIPT) else CALLER
    glob = _run_launcher(script, monkeypatch, str(tmp_path))
    model = glob["model"]
    assert type(model).__module__ == "mocodad_b200.mocodad" and type(model).__name__ == ("MoCoDADlatent" if "latent" in name else "MoCoDAD")
    assert state["n"] == int(g["noise_calls"])                        # same number of draws as the reference made
    auc = model._logged["AUC"] if hasattr(model, "_logged") else float(glob["out"][0]["AUC"])
    saved = os.path.join(ckpt_dir, f"saved_tensors_test_{cfg['aggregation_strategy']}_{cfg['n_generated_samples']}")   # save_tensors: true
    t = {f.split(".")[0]: torch.load(os.path.join(saved, f), weights_only=False) for f in os.listdir(saved)}
    assert np.array_equal(np.asarray(t["trans"]), g["trans"]) and np.array_equal(np.asarray(t["metadata"]), g["meta"])
    assert np.array_equal(np.asarray(t["frames"]), g["frames"])
    tol = 1e-4 * max(1.0, float(np.abs(g["out"]).max()))               # 1e-4 on O(1) losses; relative for the latent's O(1e3) ones
    np.testing.assert_allclose(np.asarray(t["prediction"]), g["out"], rtol=0, atol=tol)
    assert abs(auc - float(g["auc"])) <= 1e-3, (auc, float(g["auc"]))


def _write_saved_tensors(g, cfg, ckpt_dir):
    d = os.path.join(ckpt_dir, f"saved_tensors_{cfg['split']}_{cfg['aggregation_strategy']}_{cfg['n_generated_samples']}")
    os.makedirs(d)
    n = len(g["out"])
    for name, arr in (("prediction", g["out"]), ("gt_data", np.zeros((n, 2, 6, 17), np.float32)), ("trans", g["trans"]),
                      ("metadata", g["meta"]), ("frames", g["frames"])):
        torch.save(arr, os.path.join(d, name + ".pt"))


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("script", ["reference", "caller"])
def test_saved_tensor_branch_of_the_script_runs_unchanged(name, script, tmp_path, monkeypatch):
    """eval_MoCoDAD.py:26-28 (`load_tensors: true`): scores come from disk, the AUC tail runs on the host -- no GPU involved, so
    the REAL script (when the reference checkout is here) runs end to end through the overlay on this machine."""
    path = REF_SCRIPT if script == "reference" else CALLER
    if not os.path.exists(path):
        pytest.skip("reference checkout not present on this machine")
    g = np.load(os.path.join(GOLDEN, f"eval_{name}.npz"))
    cfg, ckpt_dir = _materialise(g, str(tmp_path), {"load_tensors": True})
    _write_saved_tensors(g, cfg, ckpt_dir)
    glob = _run_launcher(path, monkeypatch, str(tmp_path))
    model = glob["model"]
    assert type(model).__module__ == "mocodad_b200.mocodad"
    auc = model.test_on_saved_tensors(split_name=cfg["split"])
    assert abs(auc - float(g["auc"])) < 1e-9


def test_scoring_branch_fails_loudly_without_cuda(tmp_path, monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    g = np.load(os.path.join(GOLDEN, "eval_avenue_N2.npz"))
    _materialise(g, str(tmp_path))
    path = REF_SCRIPT if os.path.exists(REF_SCRIPT) else CALLER
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _run_launcher(path, monkeypatch, str(tmp_path))


def test_launcher_command_line(tmp_path):
    """The documented command itself, in a fresh interpreter."""
    import subprocess
    g = np.load(os.path.join(GOLDEN, "eval_avenue_N2.npz"))
    cfg, ckpt_dir = _materialise(g, str(tmp_path), {"load_tensors": True})
    _write_saved_tensors(g, cfg, ckpt_dir)
    path = REF_SCRIPT if os.path.exists(REF_SCRIPT) else CALLER
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run([sys.executable, "-m", "mocodad_b200.dropin", path, "-c", "cfg.yaml"], cwd=str(tmp_path), env=env,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert f"AUC score: {float(g['auc']):.6f}" in res.stdout
