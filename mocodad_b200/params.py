"""Parameter layout of the reference checkpoint for the scoring path.

``state_dict_spec`` lists every ``state_dict`` entry of the reference's ``MoCoDAD(args)`` module
(name -> shape, in registration order) for the 'inject' / 'no_condition' strategies, so that
Lightning checkpoints load into :class:`mocodad_b200.mocodad.MoCoDAD` unchanged.  Layout follows
  models/gcae/stsgcn.py:47-91,135-140,176-184   ST_GCNN_layer / ConvTemporalGraphical / CNN_layer
  models/stsae/stsae_unet.py:11,50-157,283-357  denoiser (joint pyramid 17/12/10, channel plan)
  models/stsae/stsae.py:44-55,136-146           condition encoder bottlenecks
  models/common/components.py:41-66,123-148     Encoder / Decoder stacks
(335 entries / 142 294 parameters for the shipped configs.)
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Sequence, Tuple

JOINT_PYRAMID = (17, 12, 10)
DOWN_CHANNELS = (16, 32, 32, 64, 64, 128, 64)
UP_CHANNELS = (64, 32, 32, 2)
BN_FIELDS = ("weight", "bias", "running_mean", "running_var")

Spec = "OrderedDict[str, Tuple[int, ...]]"


def _st_gcnn(spec, prefix: str, cin: int, cout: int, T: int, V: int, emb: Optional[int]) -> None:
    spec[prefix + "gcn.A"] = (T, V, V)
    spec[prefix + "gcn.T"] = (V, T, T)
    spec[prefix + "tcn.0.weight"] = (cout, cin, 1, 1)
    spec[prefix + "tcn.0.bias"] = (cout,)
    for k in BN_FIELDS:
        spec[prefix + "tcn.1." + k] = (cout,)
    spec[prefix + "tcn.1.num_batches_tracked"] = ()
    if cin != cout:  # stsgcn.py:69-80: 1x1 conv + BN on the skip when the width changes
        spec[prefix + "residual.0.weight"] = (cout, cin, 1, 1)
        spec[prefix + "residual.0.bias"] = (cout,)
        for k in BN_FIELDS:
            spec[prefix + "residual.1." + k] = (cout,)
        spec[prefix + "residual.1.num_batches_tracked"] = ()
    spec[prefix + "prelu.weight"] = (1,)
    if emb is not None:
        spec[prefix + "emb_layer.1.weight"] = (cout, emb)
        spec[prefix + "emb_layer.1.bias"] = (cout,)


def _cnn_layer(spec, prefix: str, vin: int, vout: int) -> None:
    spec[prefix + "block.0.weight"] = (vout, vin, 1, 1)
    spec[prefix + "block.0.bias"] = (vout,)
    for k in BN_FIELDS:
        spec[prefix + "block.1." + k] = (vout,)
    spec[prefix + "block.1.num_batches_tracked"] = ()


def state_dict_spec(T: int, T_cond: int = 3, *, num_coords: int = 2, embedding_dim: int = 16, h_dim: int = 32,
                    latent_dim: int = 16, channels: Sequence[int] = (32, 16, 32),
                    conditioning_architecture: Optional[str] = "AE", n_joints: int = 17,
                    latent_embedding_dim: int = 0, hidden_sizes: Sequence[int] = ()):
    """name -> shape of ``MoCoDAD(args).state_dict()``; ``T`` = denoised frames, ``T_cond`` = conditioning
    frames; ``conditioning_architecture`` None for 'no_condition'.  With ``latent_embedding_dim`` > 0 the layout of
    ``MoCoDADlatent(args)`` at stage 'diffusion' (models/mocodad_latent.py:42-56: STSE_Unet with its out layer
    ``to_time_dim`` instead of the full U-Net, plus the MLP ``Denoiser``, models/common/components.py:231-245)."""
    spec = OrderedDict()
    a, b, c = JOINT_PYRAMID
    E = embedding_dim
    if conditioning_architecture in ("AE", "E"):
        p = "condition_encoder."
        cin = num_coords
        for i, ch in enumerate(list(channels) + [h_dim]):
            _st_gcnn(spec, f"{p}encoder.model_layers.{i}.", cin, ch, T_cond, n_joints, None)
            cin = ch
        spec[p + "btlnk.weight"] = (latent_dim, h_dim * T_cond * n_joints)
        spec[p + "btlnk.bias"] = (latent_dim,)
        if conditioning_architecture == "AE":  # decoder half: in the checkpoint, unused at inference
            cin = h_dim
            for i, ch in enumerate(list(channels)[::-1] + [num_coords]):
                _st_gcnn(spec, f"{p}decoder.model_layers.{i}.", cin, ch, T_cond, n_joints, None)
                cin = ch
            spec[p + "rev_btlnk.weight"] = (h_dim * T_cond * n_joints, latent_dim)
            spec[p + "rev_btlnk.bias"] = (h_dim * T_cond * n_joints,)
    m = "model."
    D, U = DOWN_CHANNELS, UP_CHANNELS
    _st_gcnn(spec, m + "st_gcnnsp1a.0.", num_coords, D[0], T, a, E)
    _st_gcnn(spec, m + "st_gcnnsd1.0.", D[0], D[1], T, a, E)
    _st_gcnn(spec, m + "st_gcnnsd1.1.", D[1], D[2], T, a, E)
    _st_gcnn(spec, m + "st_gcnnsd2.0.", D[2], D[3], T, b, E)
    _st_gcnn(spec, m + "st_gcnnsd2.1.", D[3], D[4], T, b, E)
    _st_gcnn(spec, m + "st_gcnnsd3.0.", D[4], D[5], T, c, E)
    _st_gcnn(spec, m + "st_gcnnsd3.1.", D[5], D[6], T, c, E)
    _cnn_layer(spec, m + "down1.", a, b)
    _cnn_layer(spec, m + "down2.", b, c)
    if latent_embedding_dim:
        spec[m + "to_time_dim.weight"] = (latent_embedding_dim, D[6] * T * c)
        spec[m + "to_time_dim.bias"] = (latent_embedding_dim,)
        hidden = list(hidden_sizes)
        width = latent_embedding_dim
        for i, nxt in enumerate(hidden):
            p = f"denoiser.net.{i}."
            if i == len(hidden) - 1:
                spec[p + "weight"] = (nxt, width)
                spec[p + "bias"] = (nxt,)
            else:
                spec[p + "0.weight"] = (nxt, width)
                spec[p + "0.bias"] = (nxt,)
                for k in BN_FIELDS:
                    spec[p + "1." + k] = (nxt,)
                spec[p + "1.num_batches_tracked"] = ()
                width = nxt
        for i, nxt in enumerate(hidden):
            spec[f"denoiser.cond_layers.{i}.weight"] = (nxt, E)
            spec[f"denoiser.cond_layers.{i}.bias"] = (nxt,)
        return spec
    _st_gcnn(spec, m + "st_gcnnsu4.0.", D[6], U[0], T, b, E)
    _st_gcnn(spec, m + "st_gcnnsu4.1.", U[0], U[1], T, b, E)
    _st_gcnn(spec, m + "st_gcnnsu3.0.", U[1], U[2], T, a, E)
    _st_gcnn(spec, m + "st_gcnnsu3.1.", U[2], U[3], T, a, E)
    _cnn_layer(spec, m + "up2.", b, a)
    _cnn_layer(spec, m + "up3.", c, b)
    return spec
