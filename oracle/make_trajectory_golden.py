"""Pins oracle/trajectories.py (and the host window table of mocodad_b200/ingest.py) against the UNMODIFIED reference
(utils/dataset.py PoseDatasetRobust -> utils/get_robust_data.py -> utils/data.py, utils/preprocessing.py) and writes
tests/golden/trajectories.npz.  Run in the build container (needs /root/reference): python oracle/make_trajectory_golden.py

A synthetic trajectory tree in the reference's on-disk format is written to a temporary directory; the reference's
'train' split fits and pickles the RobustScaler, its 'test' split loads it -- exactly the shipped flow -- and the resulting
``segs_data_np`` / ``segs_meta`` / ``segs_ids`` and the transformed dataset items are the fixture."""
import os
import pickle
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
np.int = int  # alias numpy 2 removed, used by utils/dataset_utils.py at import time; not on the restated path

from utils.dataset import PoseDatasetRobust  # noqa: E402  (the reference)
from utils.dataset_utils import ae_trans_list  # noqa: E402
from oracle import trajectories as otr  # noqa: E402
from mocodad_b200 import ingest  # noqa: E402  (host-side table builder; no CUDA needed)

VID_RES = [640, 360]


def synth_person(rng, n_frames, first_frame, cx0, cy0, size, border=False):
    """One trajectory [n, 35]: frame number + 17 (x, y) pairs; pixel coordinates with 2 decimals like a pose estimator dump."""
    shape = rng.normal(0.0, 1.0, size=(17, 2)) * np.array([0.25, 0.5]) * size
    rows = []
    f = first_frame
    for i in range(n_frames):
        c = np.array([cx0 + 2.5 * i, cy0 + 0.7 * i])
        kp = c + shape + rng.normal(0, 0.02 * size, size=(17, 2))
        if border:
            kp = np.clip(kp, 0.0, [VID_RES[0] + 8.0, VID_RES[1] + 8.0])     # detections on / past the image border: the clip is active
        else:
            kp = np.clip(kp, 1.0, None)
        kp = np.round(kp, 2)
        miss = rng.random(17) < 0.08
        kp[miss] = 0.0                                                     # missing joints are exact (0, 0)
        rows.append(np.concatenate([[f], kp.ravel()]))
        f += 1 if rng.random() > 0.1 else 2                                # trackers skip frames now and then
    return np.asarray(rows)


def write_tree(root, rng, split_dir, n_clips, seed_shift):
    base = os.path.join(root, split_dir, "trajectories")
    for c in range(n_clips):
        folder = os.path.join(base, f"{1 + c % 2:02d}-{seed_shift + c:04d}")
        os.makedirs(folder)
        for p in range(1, 3 + (c % 2)):
            n = int(rng.integers(9, 22))
            traj = synth_person(rng, n, int(rng.integers(1, 30)), rng.uniform(60, 480), rng.uniform(60, 240), rng.uniform(40, 120),
                                border=(p == 2))
            if c == 0 and p == 1:
                traj[3, 1:] = 0.0                       # a fully missing frame inside a trajectory
                traj[5, 1::2] = 0.0                     # every x missing, y present -> compute_bounding_box's ValueError branch
                traj[7, 1:] = traj[7, 1:2].repeat(34)   # degenerate skeleton: all joints on one point of the x = y diagonal
            np.savetxt(os.path.join(folder, f"{p:04d}.csv"), traj, delimiter=",", fmt=["%d"] + ["%.2f"] * 34)
        if c == 1:                                      # a trajectory shorter than the window: contributes nothing
            np.savetxt(os.path.join(folder, "0009.csv"), synth_person(rng, 4, 3, 200, 100, 60), delimiter=",", fmt=["%d"] + ["%.2f"] * 34)


def run_reference(root, exp_dir, split, seg_len, seg_stride):
    ds = PoseDatasetRobust(path_to_data=root, exp_dir=exp_dir, include_global=False, split=split,
                           transform_list=ae_trans_list[:5], return_indices=False, return_metadata=True, debug=False,
                           headless=False, seg_len=seg_len, seg_stride=seg_stride, normalize_pose=True, kp18_format=False,
                           vid_res=VID_RES, num_coords=2, sub_mean=False, return_mean=False, symm_range=False, hip_center=False,
                           normalization_strategy="robust", ckpt=exp_dir, scaler=None, kp_threshold=0, double_item=False)
    return ds


def check_stress_rows():
    """oracle == reference on the random sweep the GPU test uses (utils/data.py:165-187 called directly)."""
    from utils.data import Trajectory
    rows = otr.stress_rows(11)
    for res in ([640, 360], [856, 480], [1080, 720]):   # the three shipped video resolutions (config/*/mocodad_test.yaml: vid_res)
        ref = Trajectory._from_image_to_centre_bounding_box(rows.copy(), video_resolution=np.array(res, dtype=np.float32))
        got = otr.bbox_centre_normalize(rows, res)
        assert ref.dtype == got.dtype and ref.tobytes() == got.tobytes(), (res, np.abs(ref - got).max())
    print("bbox-centre normalisation: oracle bit-identical to the reference on", 3 * len(rows), "stress rows")


def main():
    check_stress_rows()
    rng = np.random.default_rng(20240607)
    out = {}
    with tempfile.TemporaryDirectory() as root:
        exp_dir = os.path.join(root, "exp")
        os.makedirs(exp_dir)
        write_tree(root, rng, "training", 4, 100)
        write_tree(root, rng, "testing", 3, 200)
        run_reference(root, exp_dir, "train", 6, 1)                      # fits + pickles local_robust.pickle
        center, scale = ingest.load_robust_scaler(exp_dir)
        with open(os.path.join(exp_dir, "local_robust.pickle"), "rb") as fh:
            sk = pickle.load(fh)
        print("scaler attribute dtypes:", sk.center_.dtype, sk.scale_.dtype)
        # the train-split fit restated: rows of the training tree -> oracle normalisation -> ingest.fit_robust_scaler
        tr = ingest.load_trajectories(os.path.join(root, "training", "trajectories"))
        fitted = ingest.fit_robust_scaler(otr.bbox_centre_normalize(tr.coords, VID_RES), tr.lengths, 6, 1)
        assert fitted.center_.dtype == sk.center_.dtype and fitted.center_.tobytes() == sk.center_.tobytes(), "center_ differs"
        assert fitted.scale_.dtype == sk.scale_.dtype and fitted.scale_.tobytes() == sk.scale_.tobytes(), "scale_ differs"
        print("RobustScaler fit on", len(tr), "training trajectories: bit-identical to the reference's pickle")
        out.update(train_coords=tr.coords, train_lengths=tr.lengths)
        # a STRIDED train split (config/STC/mocodad_train.yaml: seg_stride 6): remove_short_trajectories runs with
        # input_gap = seg_stride - 1 before the fit (get_robust_data.py:44,58), so fewer trajectories reach the scaler
        exp_s = os.path.join(root, "exp_s3")
        os.makedirs(exp_s)
        run_reference(root, exp_s, "train", 6, 3)
        with open(os.path.join(exp_s, "local_robust.pickle"), "rb") as fh:
            sk3 = pickle.load(fh)
        fitted3 = ingest.fit_robust_scaler(otr.bbox_centre_normalize(tr.coords, VID_RES), tr.lengths, 6, 3)
        assert fitted3.center_.tobytes() == sk3.center_.tobytes() and fitted3.scale_.tobytes() == sk3.scale_.tobytes(), "stride-3 fit differs"
        assert sk3.center_.tobytes() != sk.center_.tobytes(), "the strided split must drop at least one trajectory for this case to bite"
        kept = int((tr.lengths >= 6 + 2 * 5).sum())
        print(f"RobustScaler fit at seg_stride 3 ({kept} of {len(tr)} trajectories long enough): bit-identical to the reference's pickle")
        out.update(center_s3=np.asarray(sk3.center_, dtype=np.float64), scale_s3=np.asarray(sk3.scale_, dtype=np.float64))
        ts = ingest.load_trajectories(os.path.join(root, "testing", "trajectories"))
        out.update(coords=ts.coords, frames=ts.frames, lengths=ts.lengths, ids=ts.ids, center=center, scale=scale,
                   vid_res=np.asarray(VID_RES, dtype=np.float32))
        for tag, seg_len, seg_stride in (("L6", 6, 1), ("L27", 27, 1), ("L6s2", 6, 2)):
            if seg_len == 27:   # long windows need long trajectories: a second test tree
                root2 = os.path.join(root, "long")
                rng2 = np.random.default_rng(5)
                folder = os.path.join(root2, "testing", "trajectories", "03-0007")
                os.makedirs(folder)
                for p in (1, 2):
                    np.savetxt(os.path.join(folder, f"{p:04d}.csv"), synth_person(rng2, 27 + 4 * p, 5, 150 * p, 90, 80, border=(p == 2)),
                               delimiter=",", fmt=["%d"] + ["%.2f"] * 34)
                ds = run_reference(root2, exp_dir, "test", seg_len, seg_stride)
                tsx = ingest.load_trajectories(os.path.join(root2, "testing", "trajectories"))
                out.update(coords_long=tsx.coords, frames_long=tsx.frames, lengths_long=tsx.lengths, ids_long=tsx.ids)
            else:
                # the reference normalises the Trajectory objects in place and parses afresh per dataset, so one call per case
                ds = run_reference(root, exp_dir, "test", seg_len, seg_stride)
                tsx = ts
            ref_base = ds.segs_data_np                                        # [N,3,L,17] float32
            ref_meta, ref_ids = np.asarray(ds.segs_meta), np.asarray(ds.segs_ids)
            # ---- pin the host window table and the oracle restatement
            starts, meta, frames = ingest.window_table(tsx, seg_len, seg_stride)
            o_starts, o_meta, o_frames = otr.window_table(tsx.lengths, tsx.frames, tsx.ids, seg_len, seg_stride)
            assert np.array_equal(starts, o_starts) and np.array_equal(meta, o_meta) and np.array_equal(frames, o_frames)
            assert np.array_equal(meta, ref_meta), "window meta differs from the reference"
            assert np.array_equal(frames, ref_ids), "window frame ids differ from the reference"
            got = otr.base_windows(tsx.coords, starts, seg_len, seg_stride, sk.center_, sk.scale_, VID_RES)
            assert got.dtype == ref_base.dtype == np.float32 and got.shape == ref_base.shape, (got.shape, ref_base.shape)
            assert got.tobytes() == ref_base.tobytes(), ("base windows differ from the reference", np.abs(got - ref_base).max())
            got64 = otr.base_windows(tsx.coords, starts, seg_len, seg_stride, center, scale, VID_RES)
            assert got64.tobytes() == ref_base.tobytes(), "float64 scaler attributes change the result"
            n = len(ds) // 5
            assert n == len(starts)
            pick = sorted(set(int(i) for i in np.random.default_rng(3).integers(0, len(ds), size=24)) | {0, len(ds) - 1})
            items = np.stack([ds[i][0] for i in pick]).astype(np.float32)    # PoseDataset.__getitem__: transform idx // N of window idx % N
            out.update({f"base_{tag}": ref_base[:, :2].copy(), f"meta_{tag}": ref_meta, f"ids_{tag}": ref_ids.astype(np.int32),
                        f"item_idx_{tag}": np.asarray(pick, dtype=np.int64), f"items_{tag}": items})
            nz = float((ref_base[:, :2] == 0).mean())
            print(f"{tag}: {len(starts)} windows bit-identical to the reference ({nz:.1%} zeros), meta + frame ids equal")
    path = os.path.join(ROOT, "tests", "golden", "trajectories.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
