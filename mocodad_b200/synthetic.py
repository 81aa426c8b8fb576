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


def synth_scored_dataset(clips: Dict[Tuple[int, int], int], *, seg_len: int = 6, num_transform: int = 2,
                         persons_per_clip: int = 3, seed: int = 5):
    """A synthetic *test epoch* in the shape ``MoCoDAD.post_processing`` consumes (mocodad.py:337-430): for every
    clip ``(scene, clip) -> n_frames`` a few persons, each present over 1-2 frame intervals, sliding windows of
    ``seg_len`` frames with stride 1 over every interval, every window replicated for each affine transformation.
    Returns (out [N] positive scores, trans [N], meta [N,4] = scene, clip, person, first frame, frames [N,seg_len]
    1-based frame numbers, gt {(scene, clip): 0/1 per frame})."""
    r = _rng(seed, f"scored{sorted(clips.items())}x{seg_len}x{num_transform}")
    out, trans, meta, frames, gt = [], [], [], [], {}
    for (scene, clip), n_frames in clips.items():
        g = np.zeros(n_frames, dtype=np.int64)
        for _ in range(max(1, n_frames // 200)):
            a = int(r.integers(0, n_frames - 10))
            g[a:a + int(r.integers(5, max(6, n_frames // 6)))] = 1
        gt[(scene, clip)] = g
        for person in range(1, persons_per_clip + 1):
            cuts = np.sort(r.choice(np.arange(1, n_frames - 1), size=3, replace=False))
            spans = [(1, int(cuts[0])), (int(cuts[1]), int(cuts[2]))] if person % 2 else [(int(cuts[0]), n_frames)]
            for lo, hi in spans:  # inclusive 1-based frame range in which the person is tracked
                for start in range(lo, hi - seg_len + 2):
                    anomalous = g[start - 1:start - 1 + seg_len].mean()
                    for tr in range(num_transform):
                        out.append(float(r.gamma(2.0, 0.01) * (1.0 + 2.0 * anomalous)))
                        trans.append(tr)
                        meta.append([scene, clip, person, start])
                        frames.append(np.arange(start, start + seg_len))
    return (np.asarray(out, dtype=np.float32), np.asarray(trans, dtype=np.int64), np.asarray(meta, dtype=np.int64),
            np.stack(frames).astype(np.int64), gt)
