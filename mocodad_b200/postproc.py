"""Score assembly -> frame-level AUC: the host-side tail of the test epoch.

Restates ``MoCoDAD.post_processing`` (models/mocodad.py:337-430) and the helpers it uses from
``utils/eval_utils.py`` (``compute_var_matrix`` :27-34, ``pad_scores`` :133-149, ``score_process`` :100-106,
``get_avenue_mask`` :152-166) as vectorised numpy over arrays already in memory, so that
``mocodad_b200.MoCoDAD.on_test_epoch_end`` works without the reference checkout.  The arithmetic that decides the
AUC is kept operation for operation (per-person max over windows, zero padding around absences, mean + log-range
mix over persons, shift + ``scipy.ndimage.gaussian_filter1d``, mean over the affine transformations,
``sklearn.metrics.roc_auc_score``); ``oracle/make_postproc_golden.py`` pins it against the unmodified reference.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def avenue_hr_mask() -> Dict[int, np.ndarray]:
    """Frames kept for HR-Avenue, by clip id (utils/eval_utils.py:152-166): run lengths of kept(1)/dropped(0) frames."""
    runs = {
        1: [(1, 75), (0, 46), (1, 269), (0, 47), (1, 427), (0, 47), (1, 20), (0, 70), (1, 438)],   # 1439 frames
        2: [(1, 272), (0, 48), (1, 403), (0, 41), (1, 447)],                                        # 1211
        3: [(1, 293), (0, 48), (1, 582)],                                                           # 923
        6: [(1, 561), (0, 64), (1, 189), (0, 193), (1, 276)],                                       # 1283
        16: [(1, 728), (0, 12)],                                                                    # 740
    }
    return {k: np.concatenate([np.full(n, v, dtype=np.int64) for v, n in r]) for k, r in runs.items()}


def person_frame_scores(loss: np.ndarray, frames: np.ndarray, n_frames: int) -> np.ndarray:
    """Per-frame score of one person in one clip: every window spreads its loss over the frames it covers and a
    frame keeps the maximum over the windows that contain it; frames never covered stay 0
    (compute_var_matrix + nanmax, mocodad.py:390-391)."""
    out = np.zeros(n_frames, dtype=np.float64)
    if len(loss) == 0:
        return out
    idx = (np.asarray(frames, dtype=np.int64) - 1).reshape(len(loss), -1)  # frame numbers are 1-based
    # numpy fancy assignment (the reference's pose[n, frames-1] = loss[n]) wraps negative indices
    idx = np.where(idx < 0, idx + n_frames, idx)
    vals = np.repeat(np.asarray(loss, dtype=np.float64), idx.shape[1])
    np.maximum.at(out, idx.reshape(-1), vals)
    return out


def pad_absences(score: np.ndarray, n_gt: int, pad_size: int) -> np.ndarray:
    """Zero the score ``pad_size`` frames around every interval in which the person is absent
    (pad_scores, utils/eval_utils.py:133-149, including its off-by-one conventions: absences are searched in
    frames [0, n_gt-2], an interval that spans that whole range is left alone, right ends are exclusive)."""
    score = score.copy()
    last = n_gt - 2
    absent = np.zeros(max(n_gt - 1, 0), dtype=bool)
    absent[:] = True
    nz = np.nonzero(score)[0]
    absent[nz[nz < n_gt - 1]] = False
    if not absent.any():
        return score
    # maximal runs [start, end] of absent frames -- the reference's ``ranges`` merges consecutive integers
    d = np.diff(np.concatenate([[0], absent.astype(np.int8), [0]]))
    starts, ends = np.nonzero(d == 1)[0], np.nonzero(d == -1)[0] - 1
    for start, end in zip(starts, ends):
        if start == 0 and end == last:
            continue
        if start == 0:
            lo, hi = start, min(end + pad_size, n_gt)
        elif end == last:
            lo, hi = max(start - pad_size, 0), end
        else:
            lo, hi = max(start - pad_size, 0), min(end + pad_size, n_gt)
        score[lo:hi] = 0
    return score


def smooth_scores(score: np.ndarray, shift: int, kernel_size: float) -> np.ndarray:
    """Shift right by ``shift`` frames (zero fill) then Gaussian-filter (score_process, eval_utils.py:100-106)."""
    from scipy.ndimage import gaussian_filter1d
    shifted = np.zeros_like(score)
    shifted[shift:] = score[:-shift]
    return gaussian_filter1d(shifted, kernel_size)


def clip_scores(out: np.ndarray, meta: np.ndarray, frames: np.ndarray, gt: np.ndarray, pad_size: int) -> np.ndarray:
    """Frame scores of one clip (one transformation) from its windows: persons are scored separately, then mixed as
    mean + (max - min of log1p) over persons (mocodad.py:379-401)."""
    n_frames = gt.shape[0]
    per_person = []
    for person in sorted(set(meta[:, 2].tolist())):
        sel = meta[:, 2] == person
        s = person_frame_scores(out[sel], frames[sel], n_frames)
        if pad_size != -1:
            s = pad_absences(s, n_frames, pad_size)
        per_person.append(s)
    return mix_persons(np.stack(per_person, axis=0))


def mix_persons(ps: np.ndarray) -> np.ndarray:
    """mean + (max - min of log1p) over the persons of a clip (mocodad.py:393-401); ``ps`` [persons, frames] float64."""
    logs = np.log1p(ps)
    return ps.mean(axis=0) + (logs.max(axis=0) - logs.min(axis=0))


def dataset_auc(out: np.ndarray, trans: np.ndarray, meta: np.ndarray, frames: np.ndarray,
                gt_by_clip: "Dict[Tuple[int, int], np.ndarray]", *, num_transform: int, pad_size: int,
                frames_shift: int, filter_kernel_size: float,
                clip_masks: Optional[Dict[Tuple[int, int], np.ndarray]] = None,
                avenue_masks: Optional[Dict[int, np.ndarray]] = None,
                return_scores: bool = False, frame_scores=None):
    """mocodad.py:337-430 with the ground truth passed as a dict {(scene, clip): 0/1 per frame}, visited in the
    reference's order (sorted file names ``{scene:02d}_{clip:04d}.npy`` == sorted (scene, clip) for zero-padded
    names).  ``clip_masks`` are the HR-UBnormal boolean masks, ``avenue_masks`` the HR-Avenue keep masks.

    ``frame_scores``: optional accelerator for the first stage -- a callable ``(loss [N], frames [N, L], row [N], row_len [rows],
    stride) -> float32 [rows, stride]`` with the semantics of ``person_frame_scores`` per row (``ScoringEngine.frame_scores_host``
    = the CUDA kernel behind ``mcd_frame_scores``; a frame maximum has no rounding, so the AUC is the same to the last bit)."""
    from sklearn.metrics import roc_auc_score
    clip_masks = clip_masks or {}
    avenue_masks = avenue_masks or {}
    out, trans, meta, frames = np.asarray(out), np.asarray(trans), np.asarray(meta), np.asarray(frames)
    # One stable sort by (transformation, scene, clip, person) replaces the reference's boolean mask over the whole epoch
    # per (transformation, clip) -- O(n log n) instead of O(transformations x clips x n); inside a group only maxima over
    # windows are taken, so the order of the windows does not matter.
    order = np.lexsort((meta[:, 2], meta[:, 1], meta[:, 0], trans))
    out, trans, meta, frames = out[order], trans[order], meta[order], frames[order]
    key = np.stack([trans.astype(np.int64), meta[:, 0].astype(np.int64), meta[:, 1].astype(np.int64)], axis=1)
    bounds = np.concatenate([[0], np.flatnonzero(np.any(key[1:] != key[:-1], axis=1)) + 1, [len(key)]]).astype(np.int64)
    groups = {tuple(key[b].tolist()): (int(b), int(e)) for b, e in zip(bounds[:-1], bounds[1:]) if e > b}
    rows_of, row_scores = {}, None
    if frame_scores is not None and len(out) and out.dtype == np.float32:   # (other dtypes: the all-host path keeps them exact)
        # one output row per (transformation, scene, clip, person) of a clip that has ground truth: consecutive in the sorted order
        pkey = np.concatenate([key, meta[:, 2:3].astype(np.int64)], axis=1)
        pb = np.concatenate([[0], np.flatnonzero(np.any(pkey[1:] != pkey[:-1], axis=1)) + 1, [len(pkey)]]).astype(np.int64)
        row = np.full(len(out), -1, dtype=np.int64)
        row_len: List[int] = []
        for b, e in zip(pb[:-1], pb[1:]):
            tr, scene, clip, _ = pkey[b].tolist()
            gt = gt_by_clip.get((scene, clip))
            if gt is None or tr >= num_transform:
                continue
            rows_of.setdefault((tr, scene, clip), []).append(len(row_len))
            row[b:e] = len(row_len)
            row_len.append(int(gt.shape[0]))
        if row_len:
            row_scores = np.asarray(frame_scores(out.astype(np.float32, copy=False), frames, row, np.asarray(row_len, dtype=np.int32),
                                                 int(max(row_len))))
    per_transform, gt_all = [], None
    for tr in range(num_transform):
        scores, gts = [], []
        for (scene, clip), gt in gt_by_clip.items():
            lo, hi = groups.get((tr, int(scene), int(clip)), (0, 0))
            rws = rows_of.get((tr, int(scene), int(clip)))
            if row_scores is not None and rws:
                ps = row_scores[rws[0]:rws[-1] + 1, :gt.shape[0]].astype(np.float64)
                if pad_size != -1:
                    ps = np.stack([pad_absences(p, gt.shape[0], pad_size) for p in ps], axis=0)
                s = mix_persons(ps)
            else:
                s = clip_scores(out[lo:hi], meta[lo:hi], frames[lo:hi], gt, pad_size)
            g = gt
            if (scene, clip) in clip_masks:
                keep = clip_masks[(scene, clip)]
                s, g = s[keep], g[keep]
            if clip in avenue_masks:
                keep = np.asarray(avenue_masks[clip]) == 1
                s, g = s[keep], g[keep]
            scores.append(smooth_scores(s, frames_shift, filter_kernel_size))
            gts.append(g)
        per_transform.append(np.concatenate(scores))
        if tr == 0:
            gt_all = np.concatenate(gts)
    pds = np.mean(np.stack(per_transform, 0), 0)
    auc = roc_auc_score(gt_all, pds)
    return (auc, pds, gt_all) if return_scores else auc


def load_ground_truth(gt_path: str) -> "Dict[Tuple[int, int], np.ndarray]":
    """{(scene, clip): labels} from ``{scene}_{clip}.npy`` files, in the reference's (sorted file name) order
    (mocodad.py:351-353)."""
    files = sorted(f for f in os.listdir(gt_path) if f.endswith(".npy"))
    res = {}
    for f in files:
        scene, clip = f.split("_")[0], f.split("_")[1].split(".")[0]
        res[(int(scene), int(clip))] = np.load(os.path.join(gt_path, f))
    return res


def hr_ubnormal_masks(split: str, root: str = "./data/UBnormal/hr_bool_masks") -> "Dict[Tuple[int, int], np.ndarray]":
    """utils/eval_utils.py:169-185"""
    from glob import glob
    sub = "testing" if "test" in split else "validating"
    res = {}
    for p in glob(os.path.join(root, sub, "test_frame_mask", "*")):
        scene, clip = map(int, os.path.basename(p).split(".")[0].split("_"))
        res[(scene, clip)] = np.load(p)
    return res
