"""Dataset + loader of the reference's scripts, on the device (SURVEY.md 8 row f1 behind ``get_dataset_and_loader``).

``utils/dataset.py:286-330`` builds a ``PoseDatasetRobust`` (CSV trajectories -> bounding-box-centre coordinates -> RobustScaler
-> sliding windows, all on the host, ``utils/get_robust_data.py:25-189``) and wraps it in a ``DataLoader`` whose workers apply
one affine transform per item (``utils/dataset.py:67-76``).  Here the host parses the CSV files and builds the integer window
table (``mocodad_b200.ingest``); the frame rows are uploaded once, normalised and scaled in place by ``mcd_normalize_frames``,
and every batch of dataset items is built in HBM by ``mcd_build_items`` -- the loader yields the reference's batch layout
``[data [B,2,seg_len,17] float32 (already on the GPU), transformation index [B], metadata [B,4], frame numbers [B,seg_len]]`` in
dataset order (item idx = transform idx // N of window idx % N).  No CPU path: without a CUDA device construction raises."""
from __future__ import annotations

import os
import pickle
from typing import Iterator, List, Optional

import numpy as np
import torch

from . import engine as _engine
from . import ingest


def _n_cond(args) -> int:
    idx = args.conditioning_indices
    if getattr(args, "conditioning_strategy", "inject") in ("no_condition", "none"):
        return 0
    return args.seg_len // idx if isinstance(idx, int) else len(list(idx))


class TrajectoryWindowDataset:
    """The reference's ``PoseDatasetRobust`` for one split: ``len()`` = num_transform x windows, ``[i]`` = the reference's item."""

    def __init__(self, args, split: str = "test", device: Optional[torch.device] = None) -> None:
        if getattr(args, "normalization_strategy", "robust") != "robust":
            raise NotImplementedError("only normalization_strategy 'robust' (every shipped config) has a device ingest path")
        if args.num_coords != 2 or args.headless or args.kp18_format or getattr(args, "hip_center", False):
            raise NotImplementedError("device ingest covers the 17-joint (x, y) layout of the shipped configs")
        if int(args.num_transform) < 1:
            raise NotImplementedError("num_transform < 1: the reference's untransformed path applies a random temporal crop per item "
                                      "(utils/dataset.py:77-83), which has no device counterpart")
        if getattr(args, "use_fitted_scaler", False):
            raise NotImplementedError("use_fitted_scaler: the reference loads robust.pkl but PoseDatasetRobust ignores it")
        if not torch.cuda.is_available():
            raise RuntimeError("mocodad_b200 builds the dataset on a CUDA device; there is no CPU fallback")
        if device is None:
            devs = getattr(args, "devices", [0])
            device = torch.device("cuda", int(devs[0]) if isinstance(devs, (list, tuple)) and devs else 0)
        self.split, self.seg_len = split, int(args.seg_len)
        self.seg_stride = int(args.seg_stride) if split == "train" else 1          # utils/dataset.py:308
        self.num_transform = int(args.num_transform)
        self.vid_res = [float(v) for v in args.vid_res]
        self.engine = _engine.ScoringEngine(seg_len=self.seg_len, n_frames_cond=_n_cond(args), noise_steps=int(args.noise_steps),
                                            device=device)   # ingest only: this handle never receives weights
        ts = ingest.load_trajectories(os.path.join(args.data_dir, ingest.split_subfolder(split), "trajectories"),
                                      debug=bool(getattr(args, "debug", False)))
        exp_dir = args.ckpt_dir
        if split == "train":                                                        # get_robust_data.py:115-119
            scaler = self.engine.fit_scaler_host(ts.coords, ts.lengths, self.vid_res, seg_stride=self.seg_stride, exp_dir=exp_dir)
            center, scale = ingest.scaler_arrays(scaler)
        elif split == "validation" and "UBnormal" not in args.data_dir:             # :120-123
            scaler = self.engine.fit_scaler_host(ts.coords, ts.lengths, self.vid_res, seg_stride=1)
            with open(os.path.join(exp_dir, "local_robust_val.pickle"), "wb") as fh:
                pickle.dump(scaler, fh)
            center, scale = ingest.scaler_arrays(scaler)
        else:                                                                       # :124-125
            center, scale = ingest.load_robust_scaler(exp_dir)
        starts, self.segs_meta, self.segs_ids = ingest.window_table(ts, self.seg_len, self.seg_stride)
        self.num_samples = len(starts)
        self.mats = _engine.pose_transform_matrices(self.num_transform)
        d_rows = torch.from_numpy(ts.coords).to(device)
        self._rows = self.engine.normalize_frames(d_rows, self.vid_res, out=d_rows, center=center, scale=scale)
        self._starts = torch.from_numpy(starts).to(device)
        self.device = device

    def __len__(self) -> int:
        return self.num_transform * self.num_samples

    def batch(self, first_item: int, n_items: int) -> List[torch.Tensor]:
        idx = np.arange(first_item, first_item + n_items)
        data = self.engine.build_items(self._rows, self._starts, mats=self.mats, first_item=first_item, n_items=n_items,
                                       row_step=self.seg_stride)
        w = idx % self.num_samples
        return [data, torch.from_numpy(idx // self.num_samples), torch.from_numpy(self.segs_meta[w]),
                torch.from_numpy(self.segs_ids[w])]

    def __getitem__(self, index: int):
        if not 0 <= index < len(self):
            raise IndexError(index)
        data, trans, meta, ids = self.batch(int(index), 1)
        return [data[0].cpu().numpy(), int(trans[0]), meta[0].numpy(), ids[0].numpy()]


class DeviceBatchLoader:
    """``DataLoader(dataset, batch_size, shuffle=False)``: consecutive batches of dataset items, built on the device."""

    def __init__(self, dataset: TrajectoryWindowDataset, batch_size: int, shuffle: bool = False) -> None:
        if shuffle:
            raise NotImplementedError("shuffled (training) batches are outside the B200 scoring path")
        self.dataset, self.batch_size = dataset, int(batch_size)

    def __len__(self) -> int:
        return -(-len(self.dataset) // self.batch_size)

    def __iter__(self) -> Iterator[List[torch.Tensor]]:
        n = len(self.dataset)
        for i0 in range(0, n, self.batch_size):
            yield self.dataset.batch(i0, min(self.batch_size, n - i0))


def get_dataset_and_loader(args, split: str = "train", validation: bool = False):
    """utils/dataset.py:286-330: (dataset, loader, val_dataset, val_loader)."""
    dataset = TrajectoryWindowDataset(args, split)
    loader = DeviceBatchLoader(dataset, args.batch_size)
    val_dataset = val_loader = None
    if validation:
        val_dataset = TrajectoryWindowDataset(args, "validation")
        val_loader = DeviceBatchLoader(val_dataset, args.batch_size)
    return dataset, loader, val_dataset, val_loader
