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


def main():
    MoCoDAD = load_reference()
    rec = {}
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
