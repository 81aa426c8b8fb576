"""Score assembly -> AUC (mocodad_b200/postproc.py) against the AUCs the unmodified reference's
``MoCoDAD.post_processing`` produced on the same synthetic test epochs (oracle/make_postproc_golden.py)."""
import numpy as np
import pytest

from mocodad_b200 import postproc, synthetic

CASES = {
    "stc_like": ({(1, 14): 265, (1, 15): 433, (6, 3): 337}, "STC", -1, 9, 15, 2),
    "avenue_like": ({(1, 4): 947, (1, 5): 1007}, "HR-Avenue", 12, 6, 30, 2),
    "hr_avenue_masked": ({(1, 3): 923, (1, 16): 740}, "HR-Avenue", 12, 6, 30, 3),
    "ubnormal_like": ({(3, 7): 301, (12, 1): 451}, "UBnormal", -1, 18, 30, 2),
}


@pytest.mark.parametrize("name", list(CASES))
def test_auc_matches_reference(golden, name):
    clips, dataset, pad, shift, ksize, ntr = CASES[name]
    out, trans, meta, frames, gt = synthetic.synth_scored_dataset(clips, num_transform=ntr)
    auc = postproc.dataset_auc(out, trans, meta, frames, gt, num_transform=ntr, pad_size=pad, frames_shift=shift,
                               filter_kernel_size=ksize,
                               avenue_masks=postproc.avenue_hr_mask() if dataset == "HR-Avenue" else None)
    assert abs(auc - float(golden("postproc")[name])) < 1e-9


def test_avenue_mask_lengths():
    m = postproc.avenue_hr_mask()
    assert {k: len(v) for k, v in m.items()} == {1: 1439, 2: 1211, 3: 923, 6: 1283, 16: 740}  # eval_utils.py:153-157


def test_person_scores_and_padding_edge_cases():
    # a window's loss covers all its frames; overlapping windows keep the maximum; absent frames stay 0
    s = postproc.person_frame_scores(np.array([0.5, 0.2, 0.9]), np.array([[1, 2, 3], [2, 3, 4], [8, 9, 10]]), 12)
    assert s.tolist() == [0.5, 0.5, 0.5, 0.2, 0, 0, 0, 0.9, 0.9, 0.9, 0, 0]
    padded = postproc.pad_absences(s, 12, 2)
    # interior absence [4,6] widens by 2 on both sides (right end exclusive: frames 2..7); the trailing absence
    # [10,10] reaches n_gt-2, so it is widened on the left only (frames 8..9) -- same as the reference's pad_scores
    assert padded.tolist() == [0.5, 0.5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
    # a person never absent, and one absent everywhere, are left untouched
    full = np.ones(6)
    assert postproc.pad_absences(full, 6, 3).tolist() == full.tolist()
    assert postproc.pad_absences(np.zeros(6), 6, 3).tolist() == [0] * 6


def test_module_post_processing_reads_gt_files(tmp_path, golden):
    """Through the module surface: MoCoDAD.post_processing with ground truth on disk (mocodad.py:351-376 layout)."""
    import argparse
    from mocodad_b200 import MoCoDAD
    from test_module import BASE
    clips, dataset, pad, shift, ksize, ntr = CASES["avenue_like"]
    out, trans, meta, frames, gt = synthetic.synth_scored_dataset(clips, num_transform=ntr)
    for (scene, clip), g in gt.items():
        np.save(tmp_path / f"{scene:02d}_{clip:04d}.npy", g)
    cfg = dict(BASE, gt_path=str(tmp_path), dataset_choice=dataset, pad_size=pad, frames_shift=shift,
               filter_kernel_size=ksize, num_transform=ntr)
    m = MoCoDAD(argparse.Namespace(**cfg))
    auc = m.post_processing(out, None, trans, meta, frames)
    assert abs(auc - float(golden("postproc")["avenue_like"])) < 1e-9
