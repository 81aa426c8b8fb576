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


def _frame_scores_host(loss, frames, row, row_len, stride):
    """The contract of the ``frame_scores`` accelerator (mcd_frame_scores), restated with the pinned per-person function."""
    frames = np.asarray(frames).reshape(len(loss), -1)
    out = np.zeros((len(row_len), stride), dtype=np.float32)
    for r in range(len(row_len)):
        sel = np.asarray(row) == r
        out[r, :row_len[r]] = postproc.person_frame_scores(np.asarray(loss)[sel], frames[sel], int(row_len[r]))
    return out


@pytest.mark.parametrize("name", list(CASES))
def test_auc_with_the_frame_score_accelerator_is_bit_identical(golden, name):
    """Host logic of the accelerated first stage: row assignment per (transformation, scene, clip, person), clips without
    windows or ground truth, padding on the accelerator's rows -- same scores to the last bit as the all-host path."""
    clips, dataset, pad, shift, ksize, ntr = CASES[name]
    out, trans, meta, frames, gt = synthetic.synth_scored_dataset(clips, num_transform=ntr)
    # windows of a clip that has no ground truth file must be ignored by both paths
    extra = slice(0, 7)
    out = np.concatenate([out, out[extra]]); trans = np.concatenate([trans, trans[extra]])
    m2 = meta[extra].copy(); m2[:, 1] = 9999
    meta = np.concatenate([meta, m2]); frames = np.concatenate([frames, frames[extra]])
    kw = dict(num_transform=ntr, pad_size=pad, frames_shift=shift, filter_kernel_size=ksize, return_scores=True,
              avenue_masks=postproc.avenue_hr_mask() if dataset == "HR-Avenue" else None)
    auc0, pds0, gt0 = postproc.dataset_auc(out, trans, meta, frames, gt, **kw)
    calls = []

    def accel(*a):
        calls.append(a[3])
        return _frame_scores_host(*a)
    auc1, pds1, gt1 = postproc.dataset_auc(out, trans, meta, frames, gt, frame_scores=accel, **kw)
    assert len(calls) == 1 and auc0 == auc1 and np.array_equal(pds0, pds1) and np.array_equal(gt0, gt1)
    assert abs(auc1 - float(golden("postproc")[name])) < 1e-9


@pytest.mark.parametrize("seed", range(6))
def test_accelerated_first_stage_on_random_epochs(seed):
    """Property: for random epochs (ragged persons per clip, clips without windows for some transformation, shuffled item
    order, padding on / off) the accelerated path and the all-host path produce the same per-frame scores bit for bit."""
    rng = np.random.default_rng(100 + seed)
    ntr, L = int(rng.integers(1, 4)), int(rng.integers(2, 7))
    gt = {(int(s), int(c)): (rng.random(int(rng.integers(30, 90))) < 0.3).astype(np.int64)
          for s, c in zip(rng.integers(1, 4, size=5), rng.permutation(20)[:5])}
    for g in gt.values():
        g[0], g[-1] = 0, 1                                 # both classes present
    items = []
    for tr in range(ntr):
        for (scene, clip), g in gt.items():
            for person in rng.permutation(6)[: int(rng.integers(1, 4))]:
                for _ in range(int(rng.integers(1, 12))):
                    f0 = int(rng.integers(1, len(g) - L + 1))
                    items.append((tr, scene, clip, int(person), f0))
    order = rng.permutation(len(items))
    items = [items[i] for i in order]
    trans = np.array([i[0] for i in items]); meta = np.array([[i[1], i[2], i[3], i[4]] for i in items])
    frames = np.array([np.arange(i[4], i[4] + L) for i in items]); out = rng.random(len(items)).astype(np.float32)
    kw = dict(num_transform=ntr, pad_size=int(rng.choice([-1, 2, 5])), frames_shift=3, filter_kernel_size=4, return_scores=True)
    a0, p0, g0 = postproc.dataset_auc(out, trans, meta, frames, gt, **kw)
    a1, p1, g1 = postproc.dataset_auc(out, trans, meta, frames, gt, frame_scores=_frame_scores_host, **kw)
    assert a0 == a1 and np.array_equal(p0, p1) and np.array_equal(g0, g1)


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


# ---- HR-UBnormal (use_hr: true, the shipped UBnormal test config): per-clip boolean keep-masks from
# ./data/UBnormal/hr_bool_masks/{testing,validating}/test_frame_mask (utils/eval_utils.py:169-185, mocodad.py:403-406).
# Same seeded masks as oracle/make_postproc_golden.py wrote for the unmodified reference.
HR_CASES = {
    "hr_ubnormal_test": ({(3, 7): 301, (12, 1): 451, (5, 2): 223}, [(3, 7), (12, 1)], "test", -1, 18, 30, 2),
    "hr_ubnormal_validation": ({(2, 9): 260, (4, 4): 340}, [(4, 4)], "validation", -1, 18, 30, 3),
}


def _hr_masks(clips, masked, seed=77):
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


@pytest.mark.parametrize("name", list(HR_CASES))
def test_hr_ubnormal_masks_through_the_module(tmp_path, monkeypatch, golden, name):
    import argparse
    from mocodad_b200 import MoCoDAD
    from test_module import BASE
    clips, masked, split, pad, shift, ksize, ntr = HR_CASES[name]
    out, trans, meta, frames, gt = synthetic.synth_scored_dataset(clips, num_transform=ntr, seed=len(name))
    gt_dir = tmp_path / "gt"
    gt_dir.mkdir()
    for (scene, clip), g in gt.items():
        np.save(gt_dir / f"{scene:02d}_{clip:04d}.npy", g)
    sub = "testing" if "test" in split else "validating"
    mdir = tmp_path / "data" / "UBnormal" / "hr_bool_masks" / sub / "test_frame_mask"
    mdir.mkdir(parents=True)
    for (scene, clip), keep in _hr_masks(clips, masked).items():
        np.save(mdir / f"{scene}_{clip}.npy", keep)
    monkeypatch.chdir(tmp_path)                                    # the reference reads the masks relative to the cwd
    want = float(golden("postproc")[name])
    cfg = dict(BASE, gt_path=str(gt_dir), dataset_choice="UBnormal", use_hr=True, split=split, pad_size=pad, frames_shift=shift,
               filter_kernel_size=ksize, num_transform=ntr)
    m = MoCoDAD(argparse.Namespace(**cfg))
    assert abs(m.post_processing(out, None, trans, meta, frames) - want) < 1e-9
    # the masks matter: without use_hr the AUC is another number
    m_all = MoCoDAD(argparse.Namespace(**dict(cfg, use_hr=False)))
    assert abs(m_all.post_processing(out, None, trans, meta, frames) - want) > 1e-6
    masks = postproc.hr_ubnormal_masks(split)
    assert sorted(masks) == sorted(masked) and all(v.dtype == bool for v in masks.values())
