"""CPU restatement of the dataset-item expansion of the reference (the first "next" row):
every dataset item is an affine transform of a base window.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and by oracle/make_transform_golden.py, which pins it against the
unmodified reference); nothing under mocodad_b200/ may import it.

Follows, in the reference tree:
  utils/dataset_utils.py:255-270   get_aff_trans_mat  (float64 cos/sin -> float32 torch matrices, flip @ (rot @ trans_scale))
  utils/dataset_utils.py:273-290   apply_pose_transform ('ktv,ck->ctv' over (x, y, 1), confidence row passed through)
  utils/dataset_utils.py:308-314   ae_trans_list (identity, flip, rot 90, rot 90 + flip, rot 45)
  utils/dataset.py:67-76           PoseDataset.__getitem__: item idx -> sample idx % N, transform idx // N, then [:num_coords]
Parity status: PINNED (oracle/make_transform_golden.py asserts bit-identity with the reference on seeded windows and
writes tests/golden/transforms.npz).
"""
import math

import numpy as np

AE_TRANSFORMS = [  # (rot degrees, flip) of ae_trans_list; sx = sy = 1, tx = ty = 0 throughout
    (0, False), (0, True), (90, False), (90, True), (45, False)]


def aff_trans_mat(rot: float = 0.0, flip: bool = False, sx: float = 1.0, sy: float = 1.0, tx: float = 0.0, ty: float = 0.0) -> np.ndarray:
    """get_aff_trans_mat, dataset_utils.py:255-270, in float32 numpy (torch.matmul on 3x3 float32 = plain fp32 dot products)."""
    cos_r, sin_r = math.cos(math.radians(rot)), math.sin(math.radians(rot))
    flip_mat = np.eye(3, dtype=np.float32)
    if flip:
        flip_mat[0, 0] = -1.0
    trans_scale = np.array([[sx, 0, tx], [0, sy, ty], [0, 0, 1]], dtype=np.float32)
    rot_mat = np.array([[cos_r, -sin_r, 0], [sin_r, cos_r, 0], [0, 0, 1]], dtype=np.float32)
    return (flip_mat @ (rot_mat @ trans_scale)).astype(np.float32)


def ae_trans_mats(n: int = 5) -> np.ndarray:
    return np.stack([aff_trans_mat(rot, flip) for rot, flip in AE_TRANSFORMS[:n]])


def apply_pose_transform(pose: np.ndarray, trans_mat: np.ndarray) -> np.ndarray:
    """dataset_utils.py:273-290 for a [3, T, V] window (x, y, confidence)."""
    conf = np.expand_dims(pose[2], axis=0)
    pose_w_ones = np.concatenate([pose[:2], np.ones_like(conf)], axis=0)
    out = np.einsum('ktv,ck->ctv', pose_w_ones, trans_mat)
    return np.concatenate([out[:2], conf], axis=0)


def dataset_item(base: np.ndarray, idx: int, mats: np.ndarray, num_coords: int = 2) -> np.ndarray:
    """PoseDataset.__getitem__, dataset.py:67-76: base [N, 3, T, V] -> transformed window [num_coords, T, V]."""
    n = base.shape[0]
    sample, trans = idx % n, idx // n
    return apply_pose_transform(np.array(base[sample]), mats[trans])[:num_coords]
