"""Deterministic synthetic checkpoints and skeleton windows (bench.py, smoke(), tests).

The reference ships no checkpoints or datasets, so measurement and parity
runs manufacture a *trained-looking* checkpoint: every entry of the reference's ``state_dict``
(:func:`mocodad_b200.params.state_dict_spec`) is filled from ``numpy.random.default_rng`` keyed by
(seed, crc32(name)) -- PCG64 streams are stable across numpy versions and machines, so the GPU box
regenerates bit-identical weights without shipping them.  BatchNorm statistics and affine terms
are randomised so that BN folding is exercised (a fresh module has BN == identity).
Input generation only: nothing here computes any part of the scoring path.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np
import torch

from .params import state_dict_spec  # noqa: F401  (re-exported)


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def synth_state_dict(spec: Dict[str, Tuple[int, ...]], seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Fill ``spec`` with trained-looking fp32 values (int64 for num_batches_tracked)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in spec.items():
        r = _rng(seed, name)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            sd[name] = torch.tensor(100, dtype=torch.int64)
            continue
        if leaf in ("A", "T"):
            bound = 1.0 / np.sqrt(shape[1])
            v = r.uniform(-bound, bound, size=shape)
        elif leaf == "running_mean":
            v = 0.1 * r.standard_normal(size=shape)
        elif leaf == "running_var":
            v = r.uniform(0.5, 1.5, size=shape)
        elif name.endswith("prelu.weight"):
            v = r.uniform(0.1, 0.4, size=shape)
        elif ".tcn.1." in name or ".residual.1." in name or ".block.1." in name:  # BN affine
            v = r.uniform(0.6, 1.4, size=shape) if leaf == "weight" else r.uniform(-0.2, 0.2, size=shape)
        else:  # conv / linear weight or bias: U(+-1/sqrt(fan_in)), fan_in from the weight's dim 1
            wname = name[: -len(leaf)] + "weight"
            fan_in = int(np.prod(spec[wname][1:]))
            bound = 1.0 / np.sqrt(fan_in)
            v = r.uniform(-bound, bound, size=shape)
        sd[name] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32).reshape(shape))
    return sd


def synth_batch(B: int, seg_len: int, V: int = 17, seed: int = 1, zero_frac: float = 0.1
                ) -> List[torch.Tensor]:
    """A dataloader batch in the reference's format (utils/dataset.py:67-110 ->
    models/mocodad.py:843-858): [data f32 [B,2,seg_len,V], trans [B] i64, meta [B,4] i64,
    frames [B,seg_len] i64].  ``zero_frac`` of the joints are exact zeros, as robust-scaled
    real poses have for missing detections."""
    r = _rng(seed, f"batch{B}x{seg_len}x{V}")
    data = r.standard_normal(size=(B, 2, seg_len, V)).astype(np.float32)
    drop = r.uniform(size=(B, 1, seg_len, V)) < zero_frac
    data = np.where(drop, np.float32(0), data).astype(np.float32)
    trans = np.arange(B, dtype=np.int64) % 5
    meta = np.stack([np.ones(B, np.int64), 1 + (np.arange(B) // 64), 1 + (np.arange(B) % 7),
                     1 + np.arange(B)], axis=1).astype(np.int64)
    frames = (1 + np.arange(B)[:, None] + np.arange(seg_len)[None, :]).astype(np.int64)
    return [torch.from_numpy(data), torch.from_numpy(trans), torch.from_numpy(meta),
            torch.from_numpy(frames)]


def synth_noise(G: int, noise_steps: int, B: int, T: int, V: int = 17, seed: int = 2) -> torch.Tensor:
    """Pre-drawn N(0,1) noise [G, noise_steps-1, B, 2, T, V]; slot 0 = x_T (mocodad.py:162),
    slot k = z after the k-th denoiser call (mocodad.py:176)."""
    r = _rng(seed, f"noise{G}x{noise_steps}x{B}x{T}x{V}")
    n = r.standard_normal(size=(G, max(noise_steps - 1, 1), B, 2, T, V)).astype(np.float32)
    return torch.from_numpy(n)
