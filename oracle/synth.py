"""Deterministic synthetic checkpoints and inputs for the parity harness.

TEST INFRASTRUCTURE ONLY (see oracle/ref_port.py header).

The reference ships no checkpoints, datasets or golden vectors, so the
harness manufactures a *trained-looking* checkpoint: every entry of the reference's
``state_dict`` (same names, same shapes) is filled from ``numpy.random.default_rng`` keyed by
(seed, crc32(name)) -- PCG64 streams are stable across numpy versions and machines, so the
GPU box regenerates bit-identical weights without shipping them.  BatchNorm statistics and
affine terms are randomised so that BN folding is exercised (a fresh module has BN == identity).

``state_dict_spec`` restates the parameter layout of
  models/stsae/stsae_unet.py:50-157, 283-357   (denoiser U-Net)
  models/gcae/stsgcn.py:47-91, 135-140, 176-184 (ST_GCNN_layer / ConvTemporalGraphical / CNN_layer)
  models/stsae/stsae.py:44-55, 136-146          (STSE / STSAE bottlenecks)
  models/common/components.py:41-66, 123-148    (Encoder / Decoder layer stacks)
and ``oracle/make_golden.py`` asserts it equals the real reference module's state_dict.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

JOINTS = {"a": 17, "b": 12, "c": 10}
DOWN = (16, 32, 32, 64, 64, 128, 64)
UP = (64, 32, 32, 2)


def _stgcn_entries(spec, prefix: str, cin: int, cout: int, T: int, V: int, emb: int | None):
    spec[prefix + "gcn.A"] = (T, V, V)
    spec[prefix + "gcn.T"] = (V, T, T)
    spec[prefix + "tcn.0.weight"] = (cout, cin, 1, 1)
    spec[prefix + "tcn.0.bias"] = (cout,)
    for k in ("weight", "bias", "running_mean", "running_var"):
        spec[prefix + "tcn.1." + k] = (cout,)
    spec[prefix + "tcn.1.num_batches_tracked"] = ()
    if cin != cout:
        spec[prefix + "residual.0.weight"] = (cout, cin, 1, 1)
        spec[prefix + "residual.0.bias"] = (cout,)
        for k in ("weight", "bias", "running_mean", "running_var"):
            spec[prefix + "residual.1." + k] = (cout,)
        spec[prefix + "residual.1.num_batches_tracked"] = ()
    spec[prefix + "prelu.weight"] = (1,)
    if emb is not None:
        spec[prefix + "emb_layer.1.weight"] = (cout, emb)
        spec[prefix + "emb_layer.1.bias"] = (cout,)


def _cnn_entries(spec, prefix: str, vin: int, vout: int):
    spec[prefix + "block.0.weight"] = (vout, vin, 1, 1)
    spec[prefix + "block.0.bias"] = (vout,)
    for k in ("weight", "bias", "running_mean", "running_var"):
        spec[prefix + "block.1." + k] = (vout,)
    spec[prefix + "block.1.num_batches_tracked"] = ()


def state_dict_spec(T: int, T_cond: int = 3, *, num_coords: int = 2, embedding_dim: int = 16,
                    h_dim: int = 32, latent_dim: int = 16, channels: Sequence[int] = (32, 16, 32),
                    conditioning_architecture: str | None = "AE", n_joints: int = 17
                    ) -> "OrderedDict[str, Tuple[int, ...]]":
    """Name -> shape for MoCoDAD(args).state_dict() under the 'inject' strategy.
    Key order follows module registration order in the reference (condition_encoder first is
    NOT the case: MoCoDAD.build_model assigns condition_encoder then model, mocodad.py:126)."""
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    a, b, c = JOINTS["a"], JOINTS["b"], JOINTS["c"]
    E = embedding_dim
    if conditioning_architecture in ("AE", "E"):
        p = "condition_encoder."
        cin = num_coords
        for i, ch in enumerate(list(channels) + [h_dim]):
            _stgcn_entries(spec, f"{p}encoder.model_layers.{i}.", cin, ch, T_cond, n_joints, None)
            cin = ch
        spec[p + "btlnk.weight"] = (latent_dim, h_dim * T_cond * n_joints)
        spec[p + "btlnk.bias"] = (latent_dim,)
        if conditioning_architecture == "AE":
            cin = h_dim
            for i, ch in enumerate(list(channels)[::-1] + [num_coords]):
                _stgcn_entries(spec, f"{p}decoder.model_layers.{i}.", cin, ch, T_cond, n_joints, None)
                cin = ch
            spec[p + "rev_btlnk.weight"] = (h_dim * T_cond * n_joints, latent_dim)
            spec[p + "rev_btlnk.bias"] = (h_dim * T_cond * n_joints,)
    m = "model."
    _stgcn_entries(spec, m + "st_gcnnsp1a.0.", num_coords, DOWN[0], T, a, E)
    _stgcn_entries(spec, m + "st_gcnnsd1.0.", DOWN[0], DOWN[1], T, a, E)
    _stgcn_entries(spec, m + "st_gcnnsd1.1.", DOWN[1], DOWN[2], T, a, E)
    _stgcn_entries(spec, m + "st_gcnnsd2.0.", DOWN[2], DOWN[3], T, b, E)
    _stgcn_entries(spec, m + "st_gcnnsd2.1.", DOWN[3], DOWN[4], T, b, E)
    _stgcn_entries(spec, m + "st_gcnnsd3.0.", DOWN[4], DOWN[5], T, c, E)
    _stgcn_entries(spec, m + "st_gcnnsd3.1.", DOWN[5], DOWN[6], T, c, E)
    _cnn_entries(spec, m + "down1.", a, b)
    _cnn_entries(spec, m + "down2.", b, c)
    _stgcn_entries(spec, m + "st_gcnnsu4.0.", DOWN[6], UP[0], T, b, E)
    _stgcn_entries(spec, m + "st_gcnnsu4.1.", UP[0], UP[1], T, b, E)
    _stgcn_entries(spec, m + "st_gcnnsu3.0.", UP[1], UP[2], T, a, E)
    _stgcn_entries(spec, m + "st_gcnnsu3.1.", UP[2], UP[3], T, a, E)
    _cnn_entries(spec, m + "up2.", b, a)
    _cnn_entries(spec, m + "up3.", c, b)
    return spec


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
