"""CPU oracle for MoCoDAD's reverse-diffusion scoring path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mocodad_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline legs do.  It is the checker, never the product.

This is a *functional restatement* of the reference's PyTorch path: plain
functions over a ``state_dict`` (the reference's own key names), using the same
ATen operators the reference calls (``einsum`` / 1x1 ``conv2d`` / eval-mode
``batch_norm`` / ``prelu`` / ``linear``) in the same order, so that on CPU it is
arithmetically the reference.  Parity status: PINNED — ``oracle/make_golden.py``
executes the unmodified reference (imported from /root/reference behind a
``pytorch_lightning`` / ``matplotlib`` stub) on seeded synthetic weights and
inputs, asserts this port is bit-identical to it on the build machine, and
commits the reference's outputs under ``tests/golden/``; ``tests/test_oracle.py``
re-checks the port against those fixtures wherever the suite runs.

Every function cites the reference file:line it follows (paths relative to the
reference checkout).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

# models/stsae/stsae_unet.py:11 -- joint pyramid is a class attribute of STSE_Unet.
JOINT_PYRAMID = (17, 12, 10)
# models/stsae/stsae_unet.py:14,229-230 -- default channel plans used by MoCoDAD.build_model
UNET_DOWN_CHANNELS = (16, 32, 32, 64, 64, 128, 64)
UNET_UP_CHANNELS = (64, 32, 32, 2)
BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, models/gcae/stsgcn.py:65


# --------------------------------------------------------------------------- a1
def cosine_betas(noise_steps: int, max_beta: float = 0.999) -> np.ndarray:
    """utils/diffusion_utils.py:8-14 + :38-44 -- float64 cosine alpha-bar betas."""
    def alpha_bar(t: float) -> float:
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2

    out = []
    for i in range(noise_steps):
        t1 = i / noise_steps
        t2 = (i + 1) / noise_steps
        out.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return np.array(out)


def schedule(noise_steps: int) -> Tuple[Tensor, Tensor, Tensor]:
    """models/mocodad.py:799-808 -- (beta, alpha, alpha_hat) as fp32 tensors."""
    beta = torch.tensor(cosine_betas(noise_steps), dtype=torch.float32)
    alpha = 1.0 - beta
    alpha_hat = torch.cumprod(alpha, dim=0)
    return beta, alpha, alpha_hat


# --------------------------------------------------------------------------- a4
def pos_encoding(t: Tensor, channels: int) -> Tensor:
    """models/stsae/stsae_unet.py:161-179 -- t is [B,1] float."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, channels, 2).float() / channels))
    a = torch.sin(t.repeat(1, channels // 2) * inv_freq)
    b = torch.cos(t.repeat(1, channels // 2) * inv_freq)
    return torch.cat([a, b], dim=-1)


# ------------------------------------------------------------------------ a6/a7
def _bn_eval(sd: StateDict, prefix: str, x: Tensor) -> Tensor:
    return F.batch_norm(x, sd[prefix + "running_mean"], sd[prefix + "running_var"],
                        sd[prefix + "weight"], sd[prefix + "bias"], training=False, eps=BN_EPS)


def graph_mix(sd: StateDict, prefix: str, x: Tensor) -> Tensor:
    """models/gcae/stsgcn.py:143-156 -- ConvTemporalGraphical: T-mix then A-mix."""
    x = torch.einsum('nctv,vtq->ncqv', (x, sd[prefix + "T"])).contiguous()
    x = torch.einsum('nctv,tvw->nctw', (x, sd[prefix + "A"])).contiguous()
    return x


def st_gcnn_layer(sd: StateDict, prefix: str, x: Tensor, temb: Optional[Tensor]) -> Tensor:
    """models/gcae/stsgcn.py:94-116 (eval mode, dropout p=0)."""
    if (prefix + "residual.0.weight") in sd:
        res = F.conv2d(x, sd[prefix + "residual.0.weight"], sd[prefix + "residual.0.bias"])
        res = _bn_eval(sd, prefix + "residual.1.", res)
    else:
        res = x
    y = graph_mix(sd, prefix + "gcn.", x)
    y = F.conv2d(y, sd[prefix + "tcn.0.weight"], sd[prefix + "tcn.0.bias"])
    y = _bn_eval(sd, prefix + "tcn.1.", y)
    y = y + res
    y = F.prelu(y, sd[prefix + "prelu.weight"])
    if temb is not None and (prefix + "emb_layer.1.weight") in sd:
        emb = F.linear(F.silu(temb), sd[prefix + "emb_layer.1.weight"], sd[prefix + "emb_layer.1.bias"])
        return y + emb[:, :, None, None]
    return y


# --------------------------------------------------------------------------- a8
def joint_resample(sd: StateDict, prefix: str, x: Tensor) -> Tensor:
    """CNN_layer over the joint axis: models/gcae/stsgcn.py:187-199 wrapped by the
    permutes at models/stsae/stsae_unet.py:205,213,381,391.  x: [B,C,T,V] -> [B,C,T,V']."""
    z = x.permute(0, 3, 1, 2).contiguous()
    z = F.conv2d(z, sd[prefix + "block.0.weight"], sd[prefix + "block.0.bias"])
    z = _bn_eval(sd, prefix + "block.1.", z)
    return z.permute(0, 2, 3, 1).contiguous()


# --------------------------------------------------------------------------- a5
UNET_BLOCKS = ("st_gcnnsp1a.0", "st_gcnnsd1.0", "st_gcnnsd1.1", "st_gcnnsd2.0", "st_gcnnsd2.1",
               "st_gcnnsd3.0", "st_gcnnsd3.1", "st_gcnnsu4.0", "st_gcnnsu4.1", "st_gcnnsu3.0",
               "st_gcnnsu3.1")


def unet_forward(sd: StateDict, x: Tensor, t: Tensor, cond_emb: Optional[Tensor],
                 prefix: str = "model.", embedding_dim: int = 16,
                 taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """models/stsae/stsae_unet.py:406-438 with _downscale :182-219 and _upscale :365-403.

    x [B,2,T,V], t [B] int64, cond_emb [B,E] or None -> predicted noise [B,2,T,V].
    ``taps`` (optional dict) receives every intermediate activation by name.
    """
    temb = pos_encoding(t.unsqueeze(-1).type(torch.float), embedding_dim)
    if cond_emb is not None:
        temb = temb + cond_emb

    def blk(name: str, h: Tensor) -> Tensor:
        h = st_gcnn_layer(sd, prefix + name + ".", h, temb)
        if taps is not None:
            taps[name] = h
        return h

    def rs(name: str, h: Tensor) -> Tensor:
        h = joint_resample(sd, prefix + name + ".", h)
        if taps is not None:
            taps[name] = h
        return h

    h = blk("st_gcnnsp1a.0", x)
    h = blk("st_gcnnsd1.0", h)
    h = blk("st_gcnnsd1.1", h)
    d1 = h
    h = rs("down1", h)
    h = blk("st_gcnnsd2.0", h)
    h = blk("st_gcnnsd2.1", h)
    d2 = h
    h = rs("down2", h)
    h = blk("st_gcnnsd3.0", h)
    h = blk("st_gcnnsd3.1", h)
    h = rs("up3", h) + d2
    h = blk("st_gcnnsu4.0", h)
    h = blk("st_gcnnsu4.1", h)
    h = rs("up2", h) + d1
    h = blk("st_gcnnsu3.0", h)
    h = blk("st_gcnnsu3.1", h)
    return h + x


# --------------------------------------------------------------------------- a3
def cond_encode(sd: StateDict, cond: Tensor, prefix: str = "condition_encoder.",
                n_layers: int = 4) -> Tensor:
    """models/stsae/stsae.py:59-92 (STSE.encode, M=1 so the permute dance is the identity)
    + models/common/components.py:68-86 (Encoder.forward).  cond [B,2,Tc,V] -> [B,latent].
    The 'AE' decoder (stsae.py:149-170) runs in the reference but its output is dropped by
    MoCoDAD.forward (mocodad.py:157), so it is not part of the scoring path."""
    h = cond
    for i in range(n_layers):
        h = st_gcnn_layer(sd, f"{prefix}encoder.model_layers.{i}.", h, None)
    h = h.reshape(h.shape[0], -1)
    return F.linear(h, sd[prefix + "btlnk.weight"], sd[prefix + "btlnk.bias"])


# --------------------------------------------------------------------------- a2
def select_frames(data: Tensor, conditioning_indices: Sequence[int]) -> Tuple[Tensor, Tensor]:
    """models/mocodad.py:743-748 -- explicit-index branch."""
    n_frames = data.shape[2]
    cond_idx = torch.tensor(list(conditioning_indices))
    corrupt_idx = torch.tensor([i for i in range(n_frames) if i not in conditioning_indices])
    return torch.index_select(data, 2, cond_idx), torch.index_select(data, 2, corrupt_idx)


# --------------------------------------------------------------------------- a9
def ddpm_update(x: Tensor, eps: Tensor, noise: Tensor, alpha: Tensor, alpha_hat: Tensor,
                beta: Tensor) -> Tensor:
    """models/mocodad.py:172-178 -- alpha/alpha_hat/beta are [B,1,1,1] gathers at step i."""
    return (1 / torch.sqrt(alpha)) * (x - ((1 - alpha) / (torch.sqrt(1 - alpha_hat))) * eps) \
        + torch.sqrt(beta) * noise


# -------------------------------------------------------------------------- a11
def window_loss(x: Tensor, target: Tensor, loss_fn: str = "smooth_l1") -> Tensor:
    """models/mocodad.py:24,66,484 -- per-window mean of the elementwise loss over (C,T,V)."""
    fn = {"l1": F.l1_loss, "smooth_l1": F.smooth_l1_loss, "mse": F.mse_loss}[loss_fn]
    per = fn(x, target, reduction="none")
    return torch.mean(per.reshape(-1, int(np.prod(target.shape[1:]))), dim=-1)


def aggregate(generated: List[Tensor], target: Tensor, strategy: str = "best",
              loss_fn: str = "smooth_l1") -> Tuple[Optional[Tensor], Tensor]:
    """models/mocodad.py:454-520 -- every strategy the reference implements except 'random'."""
    B = target.shape[0]
    losses = [window_loss(x, target, loss_fn) for x in generated]
    if strategy == "all":
        sel = torch.stack(generated).permute(1, 0, 2, 3, 4)
        return sel, torch.stack(losses).permute(1, 0)
    if strategy == "mean":
        return None, torch.mean(torch.stack(losses), dim=0)
    if strategy == "mean_pose":
        sel = torch.mean(torch.stack(generated), dim=0)
        return sel, window_loss(sel, target, loss_fn)
    if strategy == "median":
        return None, torch.median(torch.stack(losses), dim=0)[0]
    if strategy == "median_pose":
        sel = torch.median(torch.stack(generated), dim=0)[0]
        return sel, window_loss(sel, target, loss_fn)
    if strategy in ("best", "worst"):
        better = (lambda a, b: a < b) if strategy == "best" else (lambda a, b: a > b)
        loss = torch.full((B,), 1e10 if strategy == "best" else -1.0)
        sel = torch.zeros_like(target)
        for g in range(len(generated)):
            m = better(losses[g], loss)
            loss[m] = losses[g][m]
            sel[m] = generated[g][m]
        return sel, loss
    if "quantile" in strategy:
        q = float(strategy.split(":")[-1])
        return None, torch.quantile(torch.stack(losses), q, dim=0)
    raise ValueError(f"Unknown aggregation strategy {strategy}")


# ------------------------------------------------------------------ the hot loop
NoiseFn = Callable[[Tensor], Tensor]


def reverse_diffusion(sd: StateDict, data: Tensor, *, noise_steps: int, n_generated_samples: int,
                      conditioning_indices: Sequence[int] = (0, 1, 2), embedding_dim: int = 16,
                      noise: Optional[Tensor] = None, randn_like: Optional[NoiseFn] = None,
                      inject: bool = True, strategy: str = "best", loss_fn: str = "smooth_l1",
                      return_samples: bool = False):
    """models/mocodad.py:129-184 for the 'inject' (and 'no_condition') strategies.

    Noise comes either from ``noise`` -- a pre-drawn tensor [G, noise_steps-1, B,2,T,V] whose
    slot 0 is x_T and slot k>=1 is the z added after the k-th denoiser call (the last step adds
    none, mocodad.py:176, so slot noise_steps-1 does not exist) -- or from ``randn_like`` called
    in exactly the reference's order (mocodad.py:162 then :176).
    Returns (loss[B], selected_x or None[, generated list]).
    """
    beta, alpha, alpha_hat = schedule(noise_steps)
    if inject:
        cond, corrupt = select_frames(data, conditioning_indices)
        cond_emb = cond_encode(sd, cond)
    else:
        corrupt, cond_emb = data, None
    B = data.shape[0]
    generated = []
    for g in range(n_generated_samples):
        x = noise[g, 0] if noise is not None else randn_like(corrupt)
        k = 0
        for i in reversed(range(1, noise_steps)):
            t = torch.full((B,), i, dtype=torch.long)
            eps = unet_forward(sd, x, t, cond_emb, embedding_dim=embedding_dim)
            a = alpha[t][:, None, None, None]
            ah = alpha_hat[t][:, None, None, None]
            b = beta[t][:, None, None, None]
            k += 1
            if i > 1:
                z = noise[g, k] if noise is not None else randn_like(x)
            else:
                z = torch.zeros_like(x)
            x = ddpm_update(x, eps, z, a, ah, b)
        generated.append(x)
    sel, loss = aggregate(generated, corrupt, strategy, loss_fn)
    if return_samples:
        return loss, sel, generated
    return loss, sel
