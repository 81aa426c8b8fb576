"""Runs the UNMODIFIED reference script eval_MoCoDAD.py (with the reference's own models/ and utils/) on a synthetic dataset in
the reference's on-disk format and records what it produced -- per-window losses, window identities and the AUC -- as
tests/golden/eval_*.npz, together with everything the drop-in needs to repeat the run: the shipped YAML (values only), the
trajectory rows, ground-truth masks, scaler and the seeds of the checkpoint and of the injected noise.

    python oracle/make_eval_golden.py          # build container only (needs /root/reference)

TEST INFRASTRUCTURE ONLY.  Cases (BASELINE.json `configs`):
  avenue_N2        config/Avenue/mocodad_test.yaml with noise_steps: 2 -- configs[0], the CPU "plumbing" run
  ubnormal_hr      config/UBnormal/mocodad_test.yaml as shipped (use_hr: true, noise_steps 10, 50 samples) + HR mask files
  ubnormal_latent  config/UBnormal/mocodad-latent_test.yaml as shipped (MoCoDADlatent, 10 samples)
Overrides for the reference run on this GPU-less machine: accelerator 'cpu', num_workers 0; the fixture keeps the shipped
values.  batch_size is lowered to 256 in the UBnormal cases so that the epoch has several batches.
`pytorch_lightning` is mocodad_b200/lightning_standin (Trainer.test: load checkpoint, eval, test_step per batch, epoch hooks);
`matplotlib` an empty stub.  torch.randn_like / torch.randn are replaced, for the duration of the run, by a feed of seeded
tensors indexed by call number -- the drop-in test feeds the same tensors to its `b200_rng: torch` draws -- and os.listdir is
sorted so that the dataset order does not depend on the file system."""
import argparse
import os
import pickle
import runpy
import sys
import tempfile
import types
from collections import OrderedDict

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("MOCODAD_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from mocodad_b200 import ingest, synthetic  # noqa: E402  (host-side CSV reader / deterministic checkpoints: no arithmetic)

NOISE_SEED = 7000

CASES = {
    # name: (yaml, overrides, tree seed)
    "avenue_N2": ("config/Avenue/mocodad_test.yaml", {"noise_steps": 2}, 11),
    "ubnormal_hr": ("config/UBnormal/mocodad_test.yaml", {"batch_size": 256}, 12),
    "ubnormal_latent": ("config/UBnormal/mocodad-latent_test.yaml", {"batch_size": 256}, 13),
}


def noise_feed():
    """(callable, counter): call c returns randn(shape) from Generator(NOISE_SEED + c) -- the tensors of both arms."""
    state = {"n": 0}
    real_randn = torch.randn   # the feed itself is installed as torch.randn

    def feed(*shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        elif len(shape) == 1 and torch.is_tensor(shape[0]):
            shape = tuple(shape[0].shape)                                   # randn_like(tensor)
        g = torch.Generator().manual_seed(NOISE_SEED + state["n"])
        state["n"] += 1
        return real_randn(*shape, generator=g).to(kw.get("device", "cpu") or "cpu")
    return feed, state


def synth_person(rng, n_frames, first_frame, cx0, cy0, size, res):
    shape = rng.normal(0.0, 1.0, size=(17, 2)) * np.array([0.25, 0.5]) * size
    rows, f = [], first_frame
    for i in range(n_frames):
        c = np.array([cx0 + 2.5 * i, cy0 + 0.7 * i])
        kp = np.clip(c + shape + rng.normal(0, 0.02 * size, size=(17, 2)), 1.0, [res[0] - 2.0, res[1] - 2.0])
        kp = np.round(kp, 2)
        kp[rng.random(17) < 0.08] = 0.0
        rows.append(np.concatenate([[f], kp.ravel()]))
        f += 1 if rng.random() > 0.1 else 2
    return np.asarray(rows)


def write_tree(data_dir, rng, split_dir, n_clips, clip0, res):
    base = os.path.join(data_dir, split_dir, "trajectories")
    for c in range(n_clips):
        folder = os.path.join(base, f"{1 + c % 2:02d}-{clip0 + c:04d}")
        os.makedirs(folder)
        for p in range(1, 3 + (c % 2)):
            n = int(rng.integers(12, 40))
            traj = synth_person(rng, n, int(rng.integers(1, 30)), rng.uniform(60, 480), rng.uniform(60, 240), rng.uniform(40, 120), res)
            np.savetxt(os.path.join(folder, f"{p:04d}.csv"), traj, delimiter=",", fmt=["%d"] + ["%.2f"] * 34)


def install_stubs():
    sys.path.insert(0, os.path.join(ROOT, "mocodad_b200", "lightning_standin"))   # import pytorch_lightning -> the stand-in Trainer
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    np.int = int   # alias numpy 2 removed; utils/dataset_utils.py touches it at import time (not on this path)
    sys.path.insert(0, REF)


def run_case(name, yaml_rel, overrides, seed):
    cfg = yaml.load(open(os.path.join(REF, yaml_rel)), Loader=yaml.FullLoader)
    cfg.update(overrides)
    shipped_yaml = yaml.safe_dump(cfg, sort_keys=True)
    latent = "diffusion_on_latent" in cfg
    rng = np.random.default_rng(seed)
    rec = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as root:
        os.chdir(root)
        try:
            data_dir, res = cfg["data_dir"], cfg["vid_res"]
            write_tree(data_dir, rng, "training", 4, 100, res)
            write_tree(data_dir, rng, "testing", 3, 200, res)
            ckpt_dir = os.path.join(cfg["exp_dir"], cfg["dataset_choice"], cfg["dir_name"])
            os.makedirs(ckpt_dir)
            # the training run's scaler: the reference's own train split fits and pickles it (get_robust_data.py:115-119)
            from utils.dataset import PoseDatasetRobust  # type: ignore  (the reference)
            from utils.dataset_utils import ae_trans_list  # type: ignore
            real_listdir = os.listdir
            os.listdir = lambda p=".": sorted(real_listdir(p))
            PoseDatasetRobust(path_to_data=data_dir, exp_dir=ckpt_dir, include_global=False, split="train",
                              transform_list=ae_trans_list[:5], return_indices=False, return_metadata=True, debug=False, headless=False,
                              seg_len=cfg["seg_len"], seg_stride=1, normalize_pose=True, kp18_format=False, vid_res=res, num_coords=2,
                              sub_mean=False, return_mean=False, symm_range=False, hip_center=False, normalization_strategy="robust",
                              ckpt=ckpt_dir, scaler=None, kp_threshold=0, double_item=False)
            with open(os.path.join(ckpt_dir, "local_robust.pickle"), "rb") as fh:
                sk = pickle.load(fh)
            ts = ingest.load_trajectories(os.path.join(data_dir, "testing", "trajectories"))
            # ground truth per clip (+ HR masks for UBnormal with use_hr)
            os.makedirs(cfg["test_path"])
            gts, hr = {}, {}
            for scene, clip in sorted({(int(a), int(b)) for a, b, _ in ts.ids}):
                n = int(ts.frames[np.repeat((ts.ids[:, 0] == scene) & (ts.ids[:, 1] == clip), ts.lengths)].max()) + 6
                g = np.zeros(n, dtype=np.int64)
                a = int(rng.integers(5, n - 15))
                g[a:a + int(rng.integers(6, 14))] = 1
                gts[(scene, clip)] = g
                np.save(os.path.join(cfg["test_path"], f"{scene:02d}_{clip:04d}.npy"), g)
                if cfg["use_hr"] and cfg["dataset_choice"] == "UBnormal" and clip != 201:      # one clip without a mask file
                    keep = np.ones(n, dtype=bool)
                    b = int(rng.integers(0, n - 12))
                    keep[b:b + int(rng.integers(4, 10))] = False
                    hr[(scene, clip)] = keep
                    d = "./data/UBnormal/hr_bool_masks/testing/test_frame_mask"
                    os.makedirs(d, exist_ok=True)
                    np.save(os.path.join(d, f"{scene}_{clip}.npy"), keep)
            # checkpoint: seeded synthetic weights under the reference module's own state_dict names
            if latent:
                from models.mocodad_latent import MoCoDADlatent as Model  # type: ignore
                os.makedirs(os.path.dirname(cfg["pretrained_model_ckpt_path"]))
                torch.save({"state_dict": {}}, cfg["pretrained_model_ckpt_path"])
            else:
                from models.mocodad import MoCoDAD as Model  # type: ignore
            probe_args = argparse.Namespace(**dict(cfg, gt_path=cfg["test_path"], ckpt_dir=ckpt_dir))
            spec = OrderedDict((k, tuple(v.shape)) for k, v in Model(probe_args).state_dict().items())
            torch.save({"state_dict": synthetic.synth_state_dict(spec, seed=0)}, os.path.join(ckpt_dir, cfg["load_ckpt"]))
            # ---- the unmodified script ------------------------------------------------------------------------------
            run_cfg = dict(cfg, accelerator="cpu", num_workers=0)
            with open("run.yaml", "w") as fh:
                yaml.safe_dump(run_cfg, fh)
            import models.mocodad as ref_mocodad  # type: ignore
            captured = {}
            real_pp = ref_mocodad.MoCoDAD.post_processing

            def spy(self, out, gt_data, trans, meta, frames):
                captured.update(out=np.asarray(out), trans=np.asarray(trans), meta=np.asarray(meta), frames=np.asarray(frames))
                return real_pp(self, out, gt_data, trans, meta, frames)
            ref_mocodad.MoCoDAD.post_processing = spy
            feed, state = noise_feed()
            real = (torch.randn, torch.randn_like, sys.argv)
            torch.randn, torch.randn_like = feed, feed
            sys.argv = [os.path.join(REF, "eval_MoCoDAD.py"), "-c", "run.yaml"]
            try:
                glob = runpy.run_path(os.path.join(REF, "eval_MoCoDAD.py"), run_name="__main__")
            finally:
                torch.randn, torch.randn_like, sys.argv = real
                ref_mocodad.MoCoDAD.post_processing = real_pp
                os.listdir = real_listdir
            auc = float(glob["out"][0]["AUC"])
            n_items = len(captured["out"])
            per_call = cfg["n_generated_samples"] * max(cfg["noise_steps"] - 1, 1)
            n_batches = -(-n_items // cfg["batch_size"])
            assert state["n"] == per_call * n_batches, (state["n"], per_call, n_batches)
            assert type(glob["model"]).__name__ == ("MoCoDADlatent" if latent else "MoCoDAD")
            saved = os.path.join(ckpt_dir, f"saved_tensors_test_{cfg['aggregation_strategy']}_{cfg['n_generated_samples']}")
            assert sorted(os.listdir(saved)) == ["frames.pt", "gt_data.pt", "metadata.pt", "prediction.pt", "trans.pt"]
            rec.update(yaml=np.array(shipped_yaml), coords=ts.coords, frames_rows=ts.frames, lengths=ts.lengths, ids=ts.ids,
                       center=np.asarray(sk.center_), scale=np.asarray(sk.scale_),
                       gt_keys=np.asarray(sorted(gts), dtype=np.int64), hr_keys=np.asarray(sorted(hr), dtype=np.int64).reshape(-1, 2),
                       out=captured["out"].astype(np.float32), trans=captured["trans"].astype(np.int64),
                       meta=captured["meta"].astype(np.int64), frames=captured["frames"].astype(np.int64), auc=np.float64(auc),
                       noise_calls=np.int64(state["n"]), n_state_dict=np.int64(len(spec)))
            for k, g in gts.items():
                rec[f"gt_{k[0]}_{k[1]}"] = g
            for k, m in hr.items():
                rec[f"hr_{k[0]}_{k[1]}"] = m
            print(f"[eval golden] {name}: unmodified eval_MoCoDAD.py -> {type(glob['model']).__name__}, {n_items} items in {n_batches} "
                  f"batch(es), {state['n']} noise draws, AUC {auc:.6f}, scaler dtypes {sk.center_.dtype}/{sk.scale_.dtype}")
        finally:
            os.chdir(cwd)
    path = os.path.join(ROOT, "tests", "golden", f"eval_{name}.npz")
    np.savez_compressed(path, **rec)
    print("   wrote", path, os.path.getsize(path), "bytes")


def main():
    install_stubs()
    for name, (yaml_rel, overrides, seed) in CASES.items():
        run_case(name, yaml_rel, overrides, seed)


if __name__ == "__main__":
    main()
