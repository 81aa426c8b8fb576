"""Window ingest: the reference's on-disk trajectory format -> dataset windows on the device.

The reference builds its test set on the host in three numpy passes (``PoseDatasetRobust.gen_dataset``,
utils/dataset.py:213-268 -> ``data_of_combined_model``, utils/get_robust_data.py:25-189): a per-frame Python loop for
the bounding-box-centre coordinates, a materialised [N, seg_len, 34] sliding-window copy, and a scaler pass; then every
dataset item is transformed again in ``__getitem__``.  Here the host only parses the CSV files and builds the integer
window table; the frame rows cross PCIe ONCE (each row is shared by up to seg_len windows x num_transform items) and
``mcd_normalize_frames`` / ``mcd_build_items`` produce the transformed dataset items in HBM (``ScoringEngine.score_trajectories_host``).

Host-side pieces (this file) and the reference code they stand for:
  load_trajectories   utils/data.py:233-251  (folders ``{scene}-{clip}``, files ``{person}.csv``, rows ``frame,x1,y1,...,x17,y17``)
  window_table        utils/preprocessing.py:4-10, 14-52, 55-86  (window starts, [scene, clip, person, first frame] meta, frame ids)
  load_robust_scaler  utils/get_robust_data.py:18-22, 115-127  (``{exp_dir}/local_robust.pickle`` must exist for the test split)
  fit_robust_scaler   utils/get_robust_data.py:115-119, utils/data.py:345-349  (train split: fit + pickle, on device-normalised rows)
The arithmetic itself lives in csrc/mcd_kernels.cuh; there is no host implementation of it in this package.
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

N_JOINTS = 17
ROW = 2 * N_JOINTS


@dataclass
class TrajectorySet:
    """Trajectories stored back to back.  coords [F,34] float32 image coordinates, frames [F] int32 frame numbers,
    lengths [K] rows per trajectory, ids [K,3] int64 (scene, clip, person), names [K] ``{scene}-{clip}_{person}``."""
    coords: np.ndarray
    frames: np.ndarray
    lengths: np.ndarray
    ids: np.ndarray
    names: List[str]

    def __len__(self) -> int:
        return len(self.lengths)


def split_subfolder(split: str) -> str:
    """get_robust_data.py:33-38."""
    if "train" in split:
        return "training"
    if "test" in split:
        return "testing"
    return "validating"


def load_trajectories(trajectories_path: str, debug: bool = False) -> TrajectorySet:
    """utils/data.py:233-251: same directory walk (``os.listdir`` order, like the reference) and the same parser call, so
    the trajectory order -- hence the dataset index of every window -- is the reference's."""
    coords, frames, lengths, ids, names = [], [], [], [], []
    folder_names = os.listdir(trajectories_path)
    if debug:
        folder_names = folder_names[:5]
    for folder_name in folder_names:
        for csv_file_name in os.listdir(os.path.join(trajectories_path, folder_name)):
            rows = np.loadtxt(os.path.join(trajectories_path, folder_name, csv_file_name), dtype=np.float32, delimiter=",", ndmin=2)
            if rows.shape[1] != 1 + ROW:
                raise ValueError(f"{folder_name}/{csv_file_name}: expected {1 + ROW} columns (frame + 17 x,y pairs), got {rows.shape[1]}")
            person_id = csv_file_name.split(".")[0]
            scene_id, clip_id = (int(s) for s in folder_name.split("-"))   # preprocessing.py:25
            coords.append(rows[:, 1:])
            frames.append(rows[:, 0].astype(np.int32))
            lengths.append(rows.shape[0])
            ids.append([scene_id, clip_id, int(person_id)])
            names.append(folder_name + "_" + person_id)
    if not lengths:
        return TrajectorySet(np.zeros((0, ROW), np.float32), np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros((0, 3), np.int64), [])
    return TrajectorySet(np.ascontiguousarray(np.concatenate(coords), dtype=np.float32), np.concatenate(frames),
                         np.asarray(lengths, dtype=np.int64), np.asarray(ids, dtype=np.int64), names)


def window_table(ts: TrajectorySet, seg_len: int, seg_stride: int = 1) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Every window of ``seg_len`` rows (row step ``seg_stride``; the test split always uses 1, dataset.py:308) that
    fits inside one trajectory, in trajectory order (preprocessing.py:55-86; short trajectories yield none, :4-10).
    Returns (starts [N] int64 = index of the window's first row in ``ts.coords``, meta [N,4] int64 = [scene, clip, person,
    first frame number] (:26-27), frame numbers [N, seg_len] int32 (:28))."""
    if seg_len < 1 or seg_stride < 1:
        raise ValueError("window_table: seg_len and seg_stride must be positive")
    span = seg_len + (seg_stride - 1) * (seg_len - 1)
    per = np.maximum(ts.lengths - span + 1, 0)
    row0 = np.concatenate([[0], np.cumsum(ts.lengths)[:-1]]) if len(ts) else np.zeros(0, np.int64)
    traj = np.repeat(np.arange(len(ts)), per)
    first = np.concatenate([[0], np.cumsum(per)[:-1]]) if len(ts) else np.zeros(0, np.int64)
    starts = (row0[traj] + (np.arange(int(per.sum())) - first[traj])).astype(np.int64)
    rows = starts[:, None] + np.arange(seg_len, dtype=np.int64)[None, :] * seg_stride
    frames = ts.frames[rows].astype(np.int32).reshape(-1, seg_len)
    meta = np.concatenate([ts.ids[traj].reshape(-1, 3), frames[:, :1].astype(np.int64)], axis=1)
    return starts, meta, frames


def load_robust_scaler(exp_dir: str, strategy: str = "robust") -> Tuple[np.ndarray, np.ndarray]:
    """(center_, scale_) float64 [34] of the scaler the reference's training run pickled
    (get_robust_data.py:115-127: the test split only ever loads ``local_{strategy}.pickle``)."""
    path = os.path.join(exp_dir, f"local_{strategy}.pickle")
    with open(path, "rb") as fh:
        scaler = pickle.load(fh)
    return scaler_arrays(scaler)


def scaler_arrays(scaler) -> Tuple[np.ndarray, np.ndarray]:
    """center_ / scale_ of a fitted sklearn RobustScaler as float64 [34]: the transform is ``X -= center_; X /= scale_`` on
    float32 rows; float32 attributes widen exactly and one double operation rounded to float32 equals the float32 one."""
    center = np.asarray(getattr(scaler, "center_", None), dtype=np.float64)
    scale = np.asarray(getattr(scaler, "scale_", None), dtype=np.float64)
    if center.shape != (ROW,) or scale.shape != (ROW,):
        raise ValueError(f"robust scaler must be fitted with centering and scaling on {ROW} columns")
    return center, scale


def fit_robust_scaler(norm_rows: np.ndarray, lengths: np.ndarray, seg_len: int, seg_stride: int = 1, exp_dir: str = None,
                      strategy: str = "robust"):
    """The train-split half of the scaler handling (get_robust_data.py:115-119 -> scale_trajectories_robust, utils/data.py:345-349):
    fit sklearn's ``RobustScaler(quantile_range=(10, 90))`` -- the reference's own estimator, like ``roc_auc_score`` in the AUC
    tail -- on the bounding-box-centre rows of every trajectory long enough to yield a window (``remove_short_trajectories``
    runs first, preprocessing.py:4-10), zeros counted as missing.  ``norm_rows`` [F,34] are the rows ``mcd_normalize_frames``
    produced WITHOUT a scaler, back to back like ``TrajectorySet.coords``.  With ``exp_dir`` the estimator is pickled as
    ``local_{strategy}.pickle``, the file the test split loads.  Returns the fitted estimator."""
    from sklearn.preprocessing import RobustScaler
    norm_rows = np.asarray(norm_rows, dtype=np.float32)
    lengths = np.asarray(lengths, dtype=np.int64)
    if norm_rows.shape != (int(lengths.sum()), ROW):
        raise ValueError(f"norm_rows: expected [{int(lengths.sum())},{ROW}], got {norm_rows.shape}")
    span = seg_len + (seg_stride - 1) * (seg_len - 1)
    keep = np.repeat(lengths >= span, lengths)
    X = norm_rows[keep]
    if X.shape[0] == 0:
        raise ValueError("fit_robust_scaler: no trajectory is long enough for one window")
    scaler = RobustScaler(quantile_range=(10.0, 90.0))
    scaler.fit(np.where(X == 0.0, np.nan, X))
    if exp_dir is not None:
        with open(os.path.join(exp_dir, f"local_{strategy}.pickle"), "wb") as fh:
            pickle.dump(scaler, fh)
    return scaler
