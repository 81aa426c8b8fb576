"""CPU restatement of the reference's trajectory -> base-window stage (second slice): what
``PoseDatasetRobust.gen_dataset`` does between the parsed CSV rows and ``segs_data_np``.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and by oracle/make_trajectory_golden.py, which pins it against the
unmodified reference); nothing under mocodad_b200/ may import it.

Follows, in the reference tree:
  utils/data.py:11-44        compute_bounding_box (min/max over non-zero coordinates, 10 % margin, clip, round-half-even)
  utils/data.py:165-187      Trajectory._from_image_to_centre_bounding_box (per frame; missing joints -> the centre -> 0)
  utils/preprocessing.py:4-10, 55-86   remove_short_trajectories, _aggregate_rnn_autoencoder_data (windows start at every
                             frame row, rows start, start + gap + 1, ...; pred_length = 0)
  utils/preprocessing.py:14-52         aggregate_rnn_autoencoder_data(return_ids=True): meta = [scene, clip, person, first frame
                             number], ids = the frame numbers of the window
  utils/data.py:296-313, 345-354       scale_trajectories(strategy='robust'): 0 -> nan, RobustScaler.transform, nan -> 0
  utils/dataset.py:222-260   PoseDatasetRobust.gen_dataset: reshape to [N, L, 17, 2], third channel 1.0, transpose to [N, 3, L, 17]
The reference's per-frame Python loop is restated vectorised over frames, every operation in the dtype numpy 2 gives the
reference (all float32: Python scalars are weak under NEP 50; the rounded box corners are Python ints).
Parity status: PINNED (oracle/make_trajectory_golden.py runs the reference's PoseDatasetRobust on a synthetic trajectory tree,
asserts bit-identity and writes tests/golden/trajectories.npz).
"""
from __future__ import annotations

import numpy as np

N_JOINTS = 17


def bbox_centre_normalize(coords: np.ndarray, vid_res) -> np.ndarray:
    """data.py:165-187 + 11-44 for frame rows [F, 34] float32 (x1, y1, ..., x17, y17) in image coordinates."""
    c = np.asarray(coords, dtype=np.float32)
    F = c.shape[0]
    width, height = np.asarray(vid_res, dtype=np.float32)
    out = np.zeros_like(c)
    x, y = c[:, 0::2], c[:, 1::2]
    nzx, nzy = x != 0.0, y != 0.0
    # compute_bounding_box raises (-> box 0,0,0,0 -> zero width/height -> zeros) when x or y has no non-zero entry;
    # an all-zero frame is left as it is: zeros either way
    ok = nzx.any(axis=1) & nzy.any(axis=1)
    inf = np.float32(np.inf)
    left, right = np.where(nzx, x, inf).min(axis=1), np.where(nzx, x, -inf).max(axis=1)
    top, bottom = np.where(nzy, y, inf).min(axis=1), np.where(nzy, y, -inf).max(axis=1)
    left, right, top, bottom = (np.where(ok, v, np.float32(0)) for v in (left, right, top, bottom))
    one, tenth, zero = np.float32(1), np.float32(0.1), np.float32(0)
    ew, eh = tenth * (right - left + one), tenth * (bottom - top + one)
    wl, hl = width - one, height - one
    L, R = np.rint(np.clip(left - ew, zero, wl)), np.rint(np.clip(right + ew, zero, wl))      # Python round() = half to even
    Tp, Bt = np.rint(np.clip(top - eh, zero, hl)), np.rint(np.clip(bottom + eh, zero, hl))
    assert L.dtype == np.float32
    cx, cy = ((L.astype(np.float64) + R) / 2).astype(np.float32), ((Tp.astype(np.float64) + Bt) / 2).astype(np.float32)  # exact
    bw, bh = (R - L).astype(np.float32), (Bt - Tp).astype(np.float32)
    xs = np.where(nzx, x, cx[:, None]) - cx[:, None]
    ys = np.where(nzy, y, cy[:, None]) - cy[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        xs = np.where((bw != 0)[:, None], xs / bw[:, None], zero)
        ys = np.where((bh != 0)[:, None], ys / bh[:, None], zero)
    keep = ok[:, None]
    out[:, 0::2] = np.where(keep, xs, zero)
    out[:, 1::2] = np.where(keep, ys, zero)
    assert out.dtype == np.float32 and out.shape == (F, 2 * N_JOINTS)
    return out


def window_table(lengths, frame_numbers, ids, seg_len: int, seg_stride: int = 1):
    """preprocessing.py:4-10, 14-52, 55-86 over trajectories stored back to back.
    lengths [K] rows per trajectory; frame_numbers [F] int32; ids [K,3] (scene, clip, person).
    Returns (first row of every window [N] int64, meta [N,4] int64, window frame numbers [N, seg_len] int32)."""
    gap = seg_stride - 1
    span = seg_len + gap * (seg_len - 1)
    starts, meta, frames = [], [], []
    row0 = 0
    for k, n in enumerate(lengths):
        n = int(n)
        for s in range(0, n - span + 1):          # empty for short trajectories (remove_short_trajectories)
            rows = row0 + s + np.arange(seg_len) * (gap + 1)
            starts.append(row0 + s)
            meta.append([int(ids[k][0]), int(ids[k][1]), int(ids[k][2]), int(frame_numbers[row0 + s])])
            frames.append(frame_numbers[rows])
        row0 += n
    return (np.asarray(starts, dtype=np.int64), np.asarray(meta, dtype=np.int64).reshape(-1, 4),
            np.asarray(frames, dtype=np.int32).reshape(-1, seg_len))


def robust_scale(X: np.ndarray, center: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """data.py:345-354 with a fitted scaler: sklearn's RobustScaler.transform is `X -= center_; X /= scale_` in place on a
    float32 copy (whatever dtype the fitted attributes have), zeros are 'missing' and stay zero."""
    Xs = np.where(X == 0.0, np.nan, X).astype(np.float32, copy=True)
    Xs -= center
    Xs /= scale
    return np.where(np.isnan(Xs), 0.0, Xs).astype(np.float32)


def base_windows(coords: np.ndarray, starts: np.ndarray, seg_len: int, seg_stride: int, center, scale, vid_res) -> np.ndarray:
    """Raw frame rows [F,34] -> the dataset's base windows [N, 3, seg_len, 17] float32 (x, y, 1)."""
    local = bbox_centre_normalize(coords, vid_res)
    rows = starts[:, None] + np.arange(seg_len)[None, :] * seg_stride
    X = local[rows]                                                   # [N, L, 34]
    X = robust_scale(X.reshape(-1, 2 * N_JOINTS), center, scale).reshape(len(starts), seg_len, N_JOINTS, 2)
    out = np.empty((len(starts), seg_len, N_JOINTS, 3))
    out[..., :2] = X
    out[..., 2] = 1.0
    return np.transpose(out, (0, 3, 1, 2)).astype(np.float32)


def stress_rows(seed: int = 11) -> np.ndarray:
    """Bounded random sweep of frame rows for the normalisation: boxes on / past the border, single joints, x-only and
    y-only rows, degenerate boxes, half-integer coordinates (rounding ties).  oracle/make_trajectory_golden.py checks the
    restatement against the reference on exactly these rows; the GPU test checks the kernel against the restatement."""
    rng = np.random.default_rng(seed)
    rows = rng.uniform(0.0, 700.0, size=(4096, 2 * N_JOINTS)).astype(np.float32).round(2)
    rows[rng.random(rows.shape) < 0.15] = 0.0
    rows[:64] = 0.0
    rows[64:128, 0::2] = 0.0
    rows[128:192, 1::2] = 0.0
    rows[192:256] = np.repeat(rng.uniform(1, 300, size=(64, 1)).astype(np.float32), 2 * N_JOINTS, axis=1)
    rows[256:320, 2:] = 0.0
    rows[320:1320] = (rng.integers(0, 1280, size=(1000, 2 * N_JOINTS)) * 0.5).astype(np.float32)
    return rows
