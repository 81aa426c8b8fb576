"""Pin mocodad_b200/postproc.py against the UNMODIFIED reference's MoCoDAD.post_processing (mocodad.py:337-430).

Run in the build container only (needs /root/reference):   python oracle/make_postproc_golden.py
Writes tests/golden/postproc.npz: the AUCs the reference computes on synthetic test epochs that
``mocodad_b200.synthetic.synth_scored_dataset`` regenerates bit-identically anywhere.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.make_golden import load_reference  # noqa: E402  (installs the pytorch_lightning / matplotlib stubs)
from mocodad_b200 import postproc, synthetic  # noqa: E402

CASES = {
    # name: (clips, dataset_choice, pad_size, frames_shift, filter_kernel_size, num_transform)
    "stc_like": ({(1, 14): 265, (1, 15): 433, (6, 3): 337}, "STC", -1, 9, 15, 2),
    "avenue_like": ({(1, 4): 947, (1, 5): 1007}, "HR-Avenue", 12, 6, 30, 2),
    "hr_avenue_masked": ({(1, 3): 923, (1, 16): 740}, "HR-Avenue", 12, 6, 30, 3),
    "ubnormal_like": ({(3, 7): 301, (12, 1): 451}, "UBnormal", -1, 18, 30, 2),
}


# HR-UBnormal (config/UBnormal/mocodad_test.yaml: use_hr true): boolean keep-masks per clip, read by the reference from the
# cwd-relative ./data/UBnormal/hr_bool_masks/{testing,validating}/test_frame_mask/{scene}_{clip}.npy (utils/eval_utils.py:169-185)
HR_CASES = {
    # name: (clips, clips that have a mask file, split, pad_size, frames_shift, filter_kernel_size, num_transform)
    "hr_ubnormal_test": ({(3, 7): 301, (12, 1): 451, (5, 2): 223}, [(3, 7), (12, 1)], "test", -1, 18, 30, 2),
    "hr_ubnormal_validation": ({(2, 9): 260, (4, 4): 340}, [(4, 4)], "validation", -1, 18, 30, 3),
}


def hr_masks(clips, masked, seed=77):
    """Seeded boolean keep-masks (about 70 % of the frames kept, in runs) for the clips that have a mask file."""
    rng = np.random.default_rng(seed)
    res = {}
    for key in masked:
        n = clips[key]
        keep = np.ones(n, dtype=bool)
        for _ in range(4):
            a = int(rng.integers(0, n - 20))
            keep[a:a + int(rng.integers(10, n // 6))] = False
        res[key] = keep
    return res


def write_hr_masks(root, split, masks):
    sub = "testing" if "test" in split else "validating"
    d = os.path.join(root, "data", "UBnormal", "hr_bool_masks", sub, "test_frame_mask")
    os.makedirs(d)
    for (scene, clip), keep in masks.items():
        np.save(os.path.join(d, f"{scene}_{clip}.npy"), keep)


def main():
    MoCoDAD = load_reference()
    rec = {}
    for name, (clips, masked, split, pad, shift, ksize, ntr) in HR_CASES.items():
        out, trans, meta, frames, gt = synthetic.synth_scored_dataset(clips, num_transform=ntr, seed=len(name))
        with tempfile.TemporaryDirectory() as d:
            gt_dir = os.path.join(d, "gt")
            os.makedirs(gt_dir)
            for (scene, clip), g in gt.items():
                np.save(os.path.join(gt_dir, f"{scene:02d}_{clip:04d}.npy"), g)
            write_hr_masks(d, split, hr_masks(clips, masked))
            cwd = os.getcwd()
            os.chdir(d)
            try:
                me = types.SimpleNamespace(gt_path=gt_dir, split=split, use_hr=True, dataset_name="UBnormal", num_transforms=ntr,
                                           anomaly_score_pad_size=pad, anomaly_score_frames_shift=shift,
                                           anomaly_score_filter_kernel_size=ksize)
                ref_auc = MoCoDAD.post_processing(me, out, np.zeros((len(out), 2, 6, 17), np.float32), trans, meta, frames)
                me.use_hr = False
                ref_auc_all = MoCoDAD.post_processing(me, out, np.zeros((len(out), 2, 6, 17), np.float32), trans, meta, frames)
                mine = postproc.dataset_auc(out, trans, meta, frames, postproc.load_ground_truth(gt_dir), num_transform=ntr,
                                            pad_size=pad, frames_shift=shift, filter_kernel_size=ksize,
                                            clip_masks=postproc.hr_ubnormal_masks(split))
            finally:
                os.chdir(cwd)
        print(f"[postproc golden] {name}: reference AUC (use_hr) {ref_auc:.12f}  restated {mine:.12f}  (all frames: {ref_auc_all:.12f})")
        assert abs(ref_auc - mine) < 1e-12 and abs(ref_auc - ref_auc_all) > 1e-6, name
        rec[name] = np.float64(ref_auc)
    for name, (clips, dataset, pad, shift, ksize, ntr) in CASES.items():
        out, trans, meta, frames, gt = synthetic.synth_scored_dataset(clips, num_transform=ntr)
        with tempfile.TemporaryDirectory() as d:
            for (scene, clip), g in gt.items():
                np.save(os.path.join(d, f"{scene:02d}_{clip:04d}.npy"), g)
            me = types.SimpleNamespace(gt_path=d, split="test", use_hr=False, dataset_name=dataset, num_transforms=ntr,
                                       anomaly_score_pad_size=pad, anomaly_score_frames_shift=shift,
                                       anomaly_score_filter_kernel_size=ksize)
            ref_auc = MoCoDAD.post_processing(me, out, np.zeros((len(out), 2, 6, 17), np.float32), trans, meta, frames)
            mine = postproc.dataset_auc(out, trans, meta, frames, postproc.load_ground_truth(d), num_transform=ntr,
                                        pad_size=pad, frames_shift=shift, filter_kernel_size=ksize,
                                        avenue_masks=postproc.avenue_hr_mask() if dataset == "HR-Avenue" else None)
        print(f"[postproc golden] {name}: reference AUC {ref_auc:.12f}  restated {mine:.12f}  ({len(out)} windows)")
        assert abs(ref_auc - mine) < 1e-12, name
        rec[name] = np.float64(ref_auc)
    np.savez(os.path.join(ROOT, "tests", "golden", "postproc.npz"), **rec)


if __name__ == "__main__":
    main()
