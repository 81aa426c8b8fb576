"""CPU restatement of the reference's LATENT scoring path (groundwork for the next round): what
``MoCoDADlatent.forward`` does at ``stage == 'diffusion'`` (the shipped config/UBnormal/mocodad-latent_test.yaml).

TEST INFRASTRUCTURE ONLY.  There is NO product (CUDA) path for this variant yet -- ``mocodad_b200`` raises for it; this file
and tests/golden/latent_T3.npz are the oracle the kernels of the next round will be checked against.

Follows, in the reference tree:
  models/mocodad_latent.py:69-132      forward: latent code of the corrupt frames once per batch (STSE_Unet at the constant step
                                       t = -1), then per generated sample x_T ~ N(0,1) [B, latent] and noise_steps-1 calls of the
                                       MLP denoiser with the DDPM update on vectors; aggregation against the latent code
  models/stsae/stsae_unet.py:182-246   STSE_Unet._downscale + forward (down half of the denoiser U-Net, flatten, to_time_dim)
  models/common/components.py:203-300  Denoiser: Linear(+BatchNorm1d+ReLU except the last layer) + Linear(cond) per layer,
                                       cond = pos_encoding(t) + condition embedding
Parity status: PINNED (oracle/make_latent_golden.py runs the unmodified MoCoDADlatent on a seeded synthetic checkpoint with
injected noise, asserts bit-identity and writes tests/golden/latent_T3.npz).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import ref_port

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]
DOWN_BLOCKS = ("st_gcnnsp1a.0", "st_gcnnsd1.0", "st_gcnnsd1.1", "down1", "st_gcnnsd2.0", "st_gcnnsd2.1", "down2",
               "st_gcnnsd3.0", "st_gcnnsd3.1")


def latent_encode(sd: StateDict, x: Tensor, cond_emb: Optional[Tensor], embedding_dim: int = 16, prefix: str = "model.",
                  taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """STSE_Unet.forward (stsae_unet.py:222-246) at the constant step t = -1 (mocodad_latent.py:96): [B,2,T,17] -> [B, latent]."""
    B = x.shape[0]
    t = torch.full((B,), -1, dtype=torch.long)
    temb = ref_port.pos_encoding(t.unsqueeze(-1).type(torch.float), embedding_dim)
    if cond_emb is not None:
        temb = temb + cond_emb
    h = x
    for name in DOWN_BLOCKS:
        if name.startswith("down"):
            h = ref_port.joint_resample(sd, prefix + name + ".", h)
        else:
            h = ref_port.st_gcnn_layer(sd, prefix + name + ".", h, temb)
        if taps is not None:
            taps[name] = h
    return F.linear(torch.flatten(h, 1), sd[prefix + "to_time_dim.weight"], sd[prefix + "to_time_dim.bias"])


def denoiser_forward(sd: StateDict, x: Tensor, t: Tensor, cond_emb: Optional[Tensor], n_layers: int, embedding_dim: int = 16,
                     prefix: str = "denoiser.") -> Tensor:
    """Denoiser.forward (components.py:264-291), eval mode: x [B, latent], t [B] int64 -> predicted noise [B, hidden_sizes[-1]]."""
    c = ref_port.pos_encoding(t.unsqueeze(-1).type(torch.float), embedding_dim)
    if cond_emb is not None:
        c = c + cond_emb
    for i in range(n_layers):
        if i == n_layers - 1:
            x = F.linear(x, sd[f"{prefix}net.{i}.weight"], sd[f"{prefix}net.{i}.bias"])
        else:
            x = F.linear(x, sd[f"{prefix}net.{i}.0.weight"], sd[f"{prefix}net.{i}.0.bias"])
            x = F.batch_norm(x, sd[f"{prefix}net.{i}.1.running_mean"], sd[f"{prefix}net.{i}.1.running_var"],
                             sd[f"{prefix}net.{i}.1.weight"], sd[f"{prefix}net.{i}.1.bias"], training=False, eps=1e-5)
            x = F.relu(x)
        x = x + F.linear(c, sd[f"{prefix}cond_layers.{i}.weight"], sd[f"{prefix}cond_layers.{i}.bias"])
    return x


def latent_reverse_diffusion(sd: StateDict, data: Tensor, *, noise_steps: int, n_generated_samples: int, n_layers: int,
                             noise: Tensor, conditioning_indices: Sequence[int] = (0, 1, 2), embedding_dim: int = 16,
                             strategy: str = "best", loss_fn: str = "smooth_l1") -> Tuple[Tensor, Optional[Tensor], Tensor]:
    """mocodad_latent.py:69-132, stage 'diffusion'.  ``noise`` [G, noise_steps-1, B, latent]: slot 0 is x_T (torch.randn,
    :104), slot k >= 1 the z added after the k-th denoiser call (torch.randn_like, :117; none after the last).
    Returns (loss [B], selected latent or None, latent code [B, latent])."""
    beta, alpha, alpha_hat = ref_port.schedule(noise_steps)
    cond, corrupt = ref_port.select_frames(data, conditioning_indices)
    cond_emb = ref_port.cond_encode(sd, cond)
    B = data.shape[0]
    code = latent_encode(sd, corrupt, cond_emb, embedding_dim)
    generated = []
    for g in range(n_generated_samples):
        x = noise[g, 0]
        k = 0
        for i in reversed(range(1, noise_steps)):
            t = torch.full((B,), i, dtype=torch.long)
            eps = denoiser_forward(sd, x, t, cond_emb, n_layers, embedding_dim)
            a, ah, b = alpha[t][:, None], alpha_hat[t][:, None], beta[t][:, None]
            k += 1
            z = noise[g, k] if i > 1 else torch.zeros_like(code)
            x = (1 / torch.sqrt(a)) * (x - ((1 - a) / (torch.sqrt(1 - ah))) * eps) + torch.sqrt(b) * z
        generated.append(x)
    losses = [ref_port.window_loss(x, code, loss_fn) for x in generated]
    if strategy == "best":   # mocodad.py:504-512 on [B, latent] vectors
        loss = torch.full((B,), 1e10)
        sel = torch.zeros_like(code)
        for g in range(len(generated)):
            m = losses[g] < loss
            loss[m] = losses[g][m]
            sel[m] = generated[g][m]
        return loss, sel, code
    if strategy == "mean":
        return torch.mean(torch.stack(losses), dim=0), None, code
    if strategy == "median":
        return torch.median(torch.stack(losses), dim=0)[0], None, code
    raise ValueError(f"latent port: aggregation strategy {strategy} not restated")
