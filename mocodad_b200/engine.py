"""Torch-facing wrapper over the C ABI: owns one ``mcd_model`` handle and hands the library raw
device pointers of torch tensors plus the current CUDA stream.  PyTorch is used here for device
memory and streams only -- every number is produced by the CUDA library (``_lib.load()`` raises
if it is absent; there is no other implementation).

Reference being replaced: the body of ``MoCoDAD.forward`` (models/mocodad.py:129-184) and the
modules under it.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import McdConfig, check

N_JOINTS = 17
N_COORDS = 2


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class ScoringEngine:
    """One reverse-diffusion scorer on one CUDA device.

    Args mirror the fields ``MoCoDAD.__init__`` reads from the YAML namespace (mocodad.py:46-81):
    ``seg_len``, number of conditioning frames and whether they come first, ``noise_steps``,
    ``loss_fn``, ``embedding_dim`` (== ``latent_dim``), ``h_dim``, ``channels``.
    """

    def __init__(self, *, seg_len: int, n_frames_cond: int, cond_first: bool = True, noise_steps: int,
                 loss_fn: str = "smooth_l1", embedding_dim: int = 16, h_dim: int = 32,
                 channels: Sequence[int] = (32, 16, 32), device="cuda:0", n_joints: int = N_JOINTS,
                 num_coords: int = N_COORDS, latent_embedding_dim: int = 0, hidden_sizes: Sequence[int] = ()):
        """``latent_embedding_dim`` > 0 (with ``hidden_sizes``) builds the LATENT variant's engine (models/mocodad_latent.py:
        down half of the denoiser + MLP denoiser); it serves ``cond_encode`` and the ``latent_*`` methods only."""
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"ScoringEngine needs a CUDA device (got {self.device}); there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if loss_fn not in _lib.LOSS_FN:
            raise ValueError(f"unknown loss_fn {loss_fn!r}")
        ch = list(channels) + [0, 0, 0]
        hidden = [int(h) for h in hidden_sizes]
        if len(hidden) > 8:
            raise _lib.McdError(-2, f"hidden_sizes of length {len(hidden)} (at most 8)")
        self.latent_dim = int(latent_embedding_dim)
        self.cfg = McdConfig(n_coords=num_coords, n_joints=n_joints, n_frames=seg_len, n_frames_cond=n_frames_cond,
                             cond_first=int(bool(cond_first)), embedding_dim=embedding_dim, cond_h_dim=h_dim,
                             cond_channels=(C.c_int32 * 3)(*ch[:3]), noise_steps=noise_steps,
                             loss_fn=_lib.LOSS_FN[loss_fn], device=self.device.index,
                             latent_dim=self.latent_dim, n_hidden=len(hidden) if self.latent_dim else 0,
                             hidden=(C.c_int32 * 8)(*(hidden + [0] * 8)[:8]))
        if n_frames_cond > 0 and len(channels) != 3:
            raise _lib.McdError(-2, f"conditioning encoder with {len(channels)} hidden layers (kernels are built for 3)")
        handle = C.c_void_p()
        check(self.lib.mcd_model_create(C.byref(self.cfg), C.byref(handle)))
        self._h = handle
        self.seg_len, self.n_cond, self.T = seg_len, n_frames_cond, seg_len - n_frames_cond
        self.N, self.E = noise_steps, embedding_dim
        self.finalized = False
        self._ws: Optional[torch.Tensor] = None
        props = torch.cuda.get_device_properties(self.device)
        self.num_sms = props.multi_processor_count
        self.tile_unit = self.num_sms * max(1, 408 // (self.T * N_JOINTS))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and h.value:
            try:
                self.lib.mcd_model_destroy(h)
            except Exception:
                pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Mapping[str, torch.Tensor]) -> None:
        """Stream a reference ``state_dict`` (any superset of the needed entries) into the handle,
        fold BatchNorm / re-lay weights, upload.  Replaces build_model + load_state_dict."""
        keep = []
        for name, t in sd.items():
            if not torch.is_tensor(t) or not t.dtype.is_floating_point:
                continue  # num_batches_tracked
            h = t.detach().to("cpu", torch.float32).contiguous()
            keep.append(h)
            check(self.lib.mcd_model_set_tensor(self._h, name.encode(), h.data_ptr(), h.numel()))
        with torch.cuda.device(self.device):
            check(self.lib.mcd_model_finalize(self._h))
        self.finalized = True

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _chk(self, t: torch.Tensor, shape: Tuple[int, ...], name: str) -> torch.Tensor:
        if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError(f"{name}: need a contiguous float32 tensor on {self.device} "
                             f"(got {t.dtype}, {t.device}, contiguous={t.is_contiguous()})")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def workspace_bytes(self, n: int) -> int:
        return int(self.lib.mcd_workspace_bytes(self._h, int(n)))

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def _new(self, *shape) -> torch.Tensor:
        return torch.empty(*shape, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ single ops (parity surface)
    def cond_encode(self, data: torch.Tensor) -> torch.Tensor:
        """MoCoDAD._encode_condition (mocodad.py:546-560) on whole windows [B,2,seg_len,V] -> [B,E]."""
        B = data.shape[0]
        self._chk(data, (B, N_COORDS, self.seg_len, N_JOINTS), "data")
        out = self._new(B, self.E)
        ws = self._workspace(self.workspace_bytes(max(B, 1)))
        with torch.cuda.device(self.device):
            check(self.lib.mcd_cond_encode(self._h, data.data_ptr(), B, out.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()))
        return out

    def unet_forward(self, x: torch.Tensor, t: int, cond_emb: Optional[torch.Tensor]) -> torch.Tensor:
        """STSAE_Unet.forward (stsae_unet.py:406-438): predicted noise for x [n,2,T,V] at step t."""
        n = x.shape[0]
        self._chk(x, (n, N_COORDS, self.T, N_JOINTS), "x")
        cb = 0
        if cond_emb is not None:
            cb = cond_emb.shape[0]
            self._chk(cond_emb, (cb, self.E), "cond_emb")
        eps = torch.empty_like(x)
        ws = self._workspace(self.workspace_bytes(max(n, 1)))
        with torch.cuda.device(self.device):
            check(self.lib.mcd_unet_forward(self._h, x.data_ptr(), n, int(t), _ptr(cond_emb), cb, eps.data_ptr(),
                                            ws.data_ptr(), ws.numel(), self._stream()))
        return eps

    def unet_tap(self, x: torch.Tensor, t: int, cond_emb: Optional[torch.Tensor], layer: str,
                 channels: int, joints: int) -> torch.Tensor:
        n = x.shape[0]
        self._chk(x, (n, N_COORDS, self.T, N_JOINTS), "x")
        cb = 0 if cond_emb is None else cond_emb.shape[0]
        out = self._new(n, channels, self.T, joints)
        ws = self._workspace(self.workspace_bytes(max(n, 1)))
        with torch.cuda.device(self.device):
            check(self.lib.mcd_unet_tap(self._h, x.data_ptr(), n, int(t), _ptr(cond_emb), cb, layer.encode(),
                                        out.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()))
        return out

    def ddpm_step(self, x: torch.Tensor, eps: torch.Tensor, t: int, noise: Optional[torch.Tensor] = None, *,
                  seed: int = 0, first_window: int = 0, sample: int = 0, noise_slot: int = 0) -> torch.Tensor:
        """In-place DDPM update of x (mocodad.py:172-178).  Returns x."""
        n = x.shape[0]
        shape = (n, N_COORDS, self.T, N_JOINTS)
        self._chk(x, shape, "x"); self._chk(eps, shape, "eps")
        if noise is not None:
            self._chk(noise, shape, "noise")
        with torch.cuda.device(self.device):
            check(self.lib.mcd_ddpm_step(self._h, x.data_ptr(), eps.data_ptr(), _ptr(noise), n, int(t), int(seed),
                                         int(first_window), int(sample), int(noise_slot), self._stream()))
        return x

    def randn_windows(self, n: int, *, seed: int, first_window: int = 0, sample: int = 0, noise_slot: int = 0) -> torch.Tensor:
        x = self._new(n, N_COORDS, self.T, N_JOINTS)
        with torch.cuda.device(self.device):
            check(self.lib.mcd_randn_windows(self._h, x.data_ptr(), n, int(seed), int(first_window), int(sample),
                                             int(noise_slot), self._stream()))
        return x

    def window_loss(self, x0: torch.Tensor, data: torch.Tensor, G: int = 1) -> Dict[str, torch.Tensor]:
        """Per-window loss of G samples against the corrupt frames of ``data`` + best/worst over G."""
        B = data.shape[0]
        self._chk(data, (B, N_COORDS, self.seg_len, N_JOINTS), "data")
        self._chk(x0.reshape(G * B, N_COORDS, self.T, N_JOINTS), (G * B, N_COORDS, self.T, N_JOINTS), "x0")
        losses, best, worst = self._new(G, B), self._new(B), self._new(B)
        with torch.cuda.device(self.device):
            check(self.lib.mcd_window_loss(self._h, x0.data_ptr(), data.data_ptr(), B, G, losses.data_ptr(),
                                           best.data_ptr(), worst.data_ptr(), self._stream()))
        return {"losses": losses, "best": best, "worst": worst}

    def frame_scores(self, loss: torch.Tensor, frames: torch.Tensor, row: torch.Tensor, row_len: torch.Tensor,
                     stride: int) -> torch.Tensor:
        """Score assembly, first stage (mocodad.py:386-391 over utils/eval_utils.py:27-34): ``out[row[n], frames[n, k] - 1]`` =
        max over the windows n that cover the frame of ``loss[n]`` (0 where no window does; frame number 0 wraps to the row's
        last frame like the reference's numpy index).  ``row`` < 0 skips a window.  Returns float32 [rows, stride]."""
        N, rows = loss.shape[0], row_len.shape[0]
        for name, t, dt in (("loss", loss, torch.float32), ("frames", frames, torch.int64), ("row", row, torch.int64),
                            ("row_len", row_len, torch.int32)):
            if t.device != self.device or t.dtype != dt or not t.is_contiguous():
                raise ValueError(f"{name}: need a contiguous {dt} tensor on the engine's device")
        if frames.dim() != 2 or frames.shape[0] != N or row.shape != (N,) or loss.dim() != 1 or row_len.dim() != 1:
            raise ValueError("frame_scores: loss [N], frames [N, seg_len], row [N], row_len [rows]")
        out = self._new(rows, int(stride))
        with torch.cuda.device(self.device):
            check(self.lib.mcd_frame_scores(self._h, loss.data_ptr(), frames.data_ptr(), row.data_ptr(), row_len.data_ptr(), N,
                                            int(frames.shape[1]), rows, int(stride), out.data_ptr(), self._stream()))
        return out

    def frame_scores_host(self, loss: np.ndarray, frames: np.ndarray, row: np.ndarray, row_len: np.ndarray, stride: int) -> np.ndarray:
        """``frame_scores`` on host arrays (the form ``postproc.dataset_auc`` hands over at the end of a test epoch)."""
        dev = self.device
        f = np.ascontiguousarray(frames, dtype=np.int64).reshape(len(loss), -1)
        out = self.frame_scores(torch.from_numpy(np.ascontiguousarray(loss, dtype=np.float32)).to(dev),
                                torch.from_numpy(f).to(dev), torch.from_numpy(np.ascontiguousarray(row, dtype=np.int64)).to(dev),
                                torch.from_numpy(np.ascontiguousarray(row_len, dtype=np.int32)).to(dev), stride)
        return out.cpu().numpy()

    # ------------------------------------------------------------------ the hot loop
    def reverse_diffusion(self, data: torch.Tensor, n_generated_samples: int, *, noise: Optional[torch.Tensor] = None,
                          seed: int = 0, first_window: int = 0, want_losses: bool = False, want_worst: bool = False,
                          want_samples: bool = False, tile_windows: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """The body of MoCoDAD.forward (mocodad.py:129-184) for a batch ``data`` [B,2,seg_len,V].

        ``noise``: None -> in-kernel Philox keyed by (seed, first_window + b, g, slot, element); or the
        pre-drawn tensor [G, max(N-1,1), B, 2, T, V] in the reference's draw order (slot 0 = x_T).
        Returns a dict with 'best' [B] and, on request, 'losses' [G,B], 'worst' [B], 'x0' [G,B,2,T,V].
        """
        B, G = data.shape[0], int(n_generated_samples)
        self._chk(data, (B, N_COORDS, self.seg_len, N_JOINTS), "data")
        if noise is not None:
            self._chk(noise, (G, max(self.N - 1, 1), B, N_COORDS, self.T, N_JOINTS), "noise")
        out = {"best": self._new(B)}
        if want_losses:
            out["losses"] = self._new(G, B)
        if want_worst:
            out["worst"] = self._new(B)
        if want_samples:
            out["x0"] = self._new(G, B, N_COORDS, self.T, N_JOINTS)
        if B == 0:
            return out
        nv = G * B
        tile = min(nv, int(tile_windows) if tile_windows else 128 * self.tile_unit)
        ws = self._workspace(self.workspace_bytes(tile) + 4 * (B * self.E + nv) + 1024)
        with torch.cuda.device(self.device):
            check(self.lib.mcd_reverse_diffusion(
                self._h, data.data_ptr(), B, G, _ptr(noise), int(seed), int(first_window), _ptr(out.get("losses")),
                out["best"].data_ptr(), _ptr(out.get("worst")), _ptr(out.get("x0")), ws.data_ptr(), ws.numel(),
                self._stream()))
        return out

    def score_windows_host(self, data: torch.Tensor, n_generated_samples: int, *, seed: int = 0,
                           first_window: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """HOST in, HOST out: ``data`` is a CPU float32 tensor [B,2,seg_len,V] (pinned for speed); the
        call copies it to the device, runs the loop and copies the [B] 'best' scores back."""
        if data.device.type != "cpu" or data.dtype != torch.float32 or not data.is_contiguous():
            raise ValueError("score_windows_host: need a contiguous float32 CPU tensor")
        B = data.shape[0]
        if tuple(data.shape) != (B, N_COORDS, self.seg_len, N_JOINTS):
            raise ValueError(f"score_windows_host: bad shape {tuple(data.shape)}")
        if out is None:
            out = torch.empty(B, dtype=torch.float32)
        check(self.lib.mcd_score_windows_host(self._h, data.data_ptr(), B, int(n_generated_samples), int(seed),
                                              int(first_window), out.data_ptr()))
        return out

    # ------------------------------------------------------------------ f4: the latent variant (models/mocodad_latent.py)
    def latent_encode(self, data: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(conditioning embedding [B,E], latent code [B,latent]) of whole windows [B,2,seg_len,V]: ``_encode_condition`` and
        ``STSE_Unet.forward`` at the constant step -1 on the corrupt frames (mocodad_latent.py:91-100)."""
        B = data.shape[0]
        self._chk(data, (B, N_COORDS, self.seg_len, N_JOINTS), "data")
        emb, code = self._new(B, self.E), self._new(B, self.latent_dim)
        ws = self._workspace(self.workspace_bytes(max(B, 1)))
        with torch.cuda.device(self.device):
            check(self.lib.mcd_latent_encode(self._h, data.data_ptr(), B, emb.data_ptr(), code.data_ptr(), ws.data_ptr(), ws.numel(),
                                             self._stream()))
        return emb, code

    def latent_denoise(self, x: torch.Tensor, t: int, cond_emb: Optional[torch.Tensor]) -> torch.Tensor:
        """One ``Denoiser.forward`` call (components.py:264-291): predicted noise for latent vectors x [n,latent] at step t."""
        n = x.shape[0]
        self._chk(x, (n, self.latent_dim), "x")
        cb = 0
        if cond_emb is not None:
            cb = cond_emb.shape[0]
            self._chk(cond_emb, (cb, self.E), "cond_emb")
        eps = torch.empty_like(x)
        with torch.cuda.device(self.device):
            check(self.lib.mcd_latent_denoise(self._h, x.data_ptr(), n, int(t), _ptr(cond_emb), cb, eps.data_ptr(), self._stream()))
        return eps

    def latent_reverse_diffusion(self, data: torch.Tensor, n_generated_samples: int, *, noise: Optional[torch.Tensor] = None,
                                 seed: int = 0, first_window: int = 0, want_losses: bool = False, want_worst: bool = False,
                                 want_samples: bool = False) -> Dict[str, torch.Tensor]:
        """``MoCoDADlatent.forward`` at stage 'diffusion' (mocodad_latent.py:69-132) for a batch ``data`` [B,2,seg_len,V].
        ``noise``: None -> Philox, or [G, max(N-1,1), B, latent] in the reference's draw order (slot 0 = x_T).
        Returns 'best' [B], 'code' [B,latent] and, on request, 'losses' [G,B], 'worst' [B], 'x0' [G,B,latent]."""
        B, G, L = data.shape[0], int(n_generated_samples), self.latent_dim
        self._chk(data, (B, N_COORDS, self.seg_len, N_JOINTS), "data")
        if noise is not None:
            self._chk(noise, (G, max(self.N - 1, 1), B, L), "noise")
        out = {"best": self._new(B), "code": self._new(B, L)}
        if want_losses:
            out["losses"] = self._new(G, B)
        if want_worst:
            out["worst"] = self._new(B)
        if want_samples:
            out["x0"] = self._new(G, B, L)
        if B == 0:
            return out
        ws = self._workspace(self.workspace_bytes(B) + 4 * (B * self.E + G * B) + 1024)
        with torch.cuda.device(self.device):
            check(self.lib.mcd_latent_reverse_diffusion(
                self._h, data.data_ptr(), B, G, _ptr(noise), int(seed), int(first_window), _ptr(out.get("losses")),
                out["best"].data_ptr(), _ptr(out.get("worst")), _ptr(out.get("x0")), out["code"].data_ptr(), ws.data_ptr(),
                ws.numel(), self._stream()))
        return out

    # ------------------------------------------------------------------ f1: dataset items on the device
    def expand_transforms(self, base: torch.Tensor, mats: np.ndarray, first_item: int, n_items: int) -> torch.Tensor:
        """Dataset items ``first_item .. first_item + n_items - 1`` of the reference's ``PoseDataset`` built on the device:
        item idx = transform ``idx // N`` (rows of ``mats`` [K,6]) of base window ``idx % N`` (utils/dataset.py:67-76)."""
        N = base.shape[0]
        self._chk(base, (N, N_COORDS, self.seg_len, N_JOINTS), "base")
        mats = np.ascontiguousarray(mats, dtype=np.float32).reshape(-1, 6)
        out = self._new(int(n_items), N_COORDS, self.seg_len, N_JOINTS)
        with torch.cuda.device(self.device):
            check(self.lib.mcd_expand_transforms(self._h, base.data_ptr(), N, mats.ctypes.data_as(_lib.c_float_p), mats.shape[0],
                                                 int(first_item), int(n_items), out.data_ptr(), self._stream()))
        return out

    def score_dataset_host(self, base: torch.Tensor, n_generated_samples: int, *, num_transform: int = 5, batch: int = 1024,
                           seed: int = 0) -> torch.Tensor:
        """HOST in, HOST out over the whole ``num_transform``-fold dataset: the base windows [N,2,seg_len,V] cross PCIe
        once, every batch of dataset items is expanded on the device (``expand_transforms``) and scored; Philox noise is
        keyed by the dataset index, so the result equals scoring the materialised dataset batch by batch."""
        if base.device.type != "cpu" or base.dtype != torch.float32 or not base.is_contiguous():
            raise ValueError("score_dataset_host: need a contiguous float32 CPU tensor")
        N = base.shape[0]
        mats = pose_transform_matrices(num_transform)
        d_base = base.to(self.device, non_blocking=True)
        total = num_transform * N
        scores = torch.empty(total, dtype=torch.float32, device=self.device)
        for i0 in range(0, total, batch):
            n = min(batch, total - i0)
            items = self.expand_transforms(d_base, mats, i0, n)
            scores[i0:i0 + n] = self.reverse_diffusion(items, n_generated_samples, seed=seed, first_window=i0)["best"]
        return scores.cpu()

    # ------------------------------------------------------------------ f1: trajectory rows -> dataset items on the device
    @staticmethod
    def _scaler_ptrs(center, scale):
        if center is None and scale is None:
            return None, None, None
        if center is None or scale is None:
            raise ValueError("center and scale must be given together")
        center = np.ascontiguousarray(center, dtype=np.float64).reshape(2 * N_JOINTS)
        scale = np.ascontiguousarray(scale, dtype=np.float64).reshape(2 * N_JOINTS)
        return center.ctypes.data_as(_lib.c_double_p), scale.ctypes.data_as(_lib.c_double_p), (center, scale)

    def normalize_frames(self, rows: torch.Tensor, vid_res: Sequence[float], out: Optional[torch.Tensor] = None, *,
                         center: Optional[np.ndarray] = None, scale: Optional[np.ndarray] = None) -> torch.Tensor:
        """Bounding-box-centre coordinates of every frame row [F,34] (utils/data.py:165-187, 11-44); ``out`` may be ``rows``.
        With ``center`` / ``scale`` (the fitted RobustScaler's ``center_`` / ``scale_``) the rows are robust-scaled as well
        (utils/data.py:345-354) -- once per row instead of once per window that contains it."""
        F = rows.shape[0]
        self._chk(rows, (F, 2 * N_JOINTS), "rows")
        out = torch.empty_like(rows) if out is None else self._chk(out, (F, 2 * N_JOINTS), "out")
        cp, sp, _keep = self._scaler_ptrs(center, scale)
        with torch.cuda.device(self.device):
            check(self.lib.mcd_normalize_frames(self._h, rows.data_ptr(), F, float(vid_res[0]), float(vid_res[1]), cp, sp,
                                                out.data_ptr(), self._stream()))
        return out

    def build_items(self, rows: torch.Tensor, win_start: torch.Tensor, center: Optional[np.ndarray] = None,
                    scale: Optional[np.ndarray] = None, *, mats: Optional[np.ndarray] = None, first_item: int = 0,
                    n_items: Optional[int] = None, row_step: int = 1) -> torch.Tensor:
        """Dataset items ``first_item .. first_item + n_items - 1`` [n,2,seg_len,17] straight from the normalised frame rows:
        window ``idx % N`` (rows ``win_start[w] + k * row_step``), robust-scaled here (``center`` / ``scale`` given) or already by
        ``normalize_frames``, transform ``idx // N`` (``mats`` [K,6]; None = the untransformed base windows).
        utils/preprocessing.py:55-86, utils/data.py:345-354, utils/dataset.py:67-76, 241-256."""
        F, N = rows.shape[0], win_start.shape[0]
        self._chk(rows, (F, 2 * N_JOINTS), "rows")
        if win_start.device != self.device or win_start.dtype != torch.int64 or not win_start.is_contiguous() or win_start.dim() != 1:
            raise ValueError("win_start: need a contiguous 1-D int64 tensor on the engine's device")
        cp, sp, _keep = self._scaler_ptrs(center, scale)
        K = 1
        mats_p = None
        if mats is not None:
            mats = np.ascontiguousarray(mats, dtype=np.float32).reshape(-1, 6)
            K, mats_p = mats.shape[0], mats.ctypes.data_as(_lib.c_float_p)
        n_items = K * N - first_item if n_items is None else int(n_items)
        out = self._new(max(n_items, 0), N_COORDS, self.seg_len, N_JOINTS)
        with torch.cuda.device(self.device):
            check(self.lib.mcd_build_items(self._h, rows.data_ptr(), F, win_start.data_ptr(), N, int(row_step), cp, sp, mats_p, K,
                                           int(first_item), n_items, out.data_ptr(), self._stream()))
        return out

    def fit_scaler_host(self, coords: np.ndarray, lengths: np.ndarray, vid_res: Sequence[float], *, seg_stride: int = 1,
                        exp_dir: Optional[str] = None):
        """Train-split scaler fit (utils/get_robust_data.py:115-119): the frame rows are normalised on the device
        (``normalize_frames`` without a scaler), the quantiles are sklearn's own ``RobustScaler.fit`` on the host
        (``mocodad_b200.ingest.fit_robust_scaler``); with ``exp_dir`` the estimator is pickled where the test split looks for it."""
        from . import ingest
        coords = np.ascontiguousarray(coords, dtype=np.float32)
        if coords.ndim != 2 or coords.shape[1] != 2 * N_JOINTS:
            raise ValueError(f"coords: expected [F,{2 * N_JOINTS}], got {coords.shape}")
        d_rows = torch.from_numpy(coords).to(self.device)
        norm = self.normalize_frames(d_rows, vid_res, out=d_rows).cpu().numpy()
        return ingest.fit_robust_scaler(norm, lengths, self.seg_len, seg_stride, exp_dir=exp_dir)

    def score_trajectories_host(self, coords: np.ndarray, win_start: np.ndarray, center: np.ndarray, scale: np.ndarray,
                                vid_res: Sequence[float], n_generated_samples: int, *, num_transform: int = 5, batch: int = 1024,
                                seed: int = 0, row_step: int = 1, item_range: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        """HOST in, HOST out from the parsed trajectory files: frame rows ``coords`` [F,34] (image coordinates) and the window
        table ``win_start`` [N] (``mocodad_b200.ingest``) cross PCIe once; normalisation, windowing, scaling and the
        ``num_transform`` transforms happen in HBM, batch by batch, in front of the reverse-diffusion loop.  Returns the 'best'
        score of dataset items ``item_range`` (default: all ``num_transform * N``; a rank passes its shard) -- Philox noise is
        keyed by the dataset index, so the scores do not depend on ``batch`` or on how the range is split across ranks."""
        coords = np.ascontiguousarray(coords, dtype=np.float32)
        win_start = np.ascontiguousarray(win_start, dtype=np.int64)
        if coords.ndim != 2 or coords.shape[1] != 2 * N_JOINTS:
            raise ValueError(f"coords: expected [F,{2 * N_JOINTS}], got {coords.shape}")
        N, F = win_start.shape[0], coords.shape[0]
        total = num_transform * N
        lo, hi = (0, total) if item_range is None else (int(item_range[0]), int(item_range[1]))
        if not 0 <= lo <= hi <= total:
            raise ValueError(f"item_range {item_range} outside the dataset of {total} items")
        if hi == lo:
            return torch.empty(0, dtype=torch.float32)
        if N and (win_start.min() < 0 or win_start.max() + (self.seg_len - 1) * row_step >= F):
            raise ValueError("win_start: a window reaches outside the frame rows")
        mats = pose_transform_matrices(num_transform)
        d_rows = torch.from_numpy(coords).to(self.device, non_blocking=True)
        d_start = torch.from_numpy(win_start).to(self.device, non_blocking=True)
        self.normalize_frames(d_rows, vid_res, out=d_rows, center=center, scale=scale)   # normalised + scaled, in place
        scores = torch.empty(hi - lo, dtype=torch.float32, device=self.device)
        for i0 in range(lo, hi, batch):
            n = min(batch, hi - i0)
            items = self.build_items(d_rows, d_start, mats=mats, first_item=i0, n_items=n, row_step=row_step)
            scores[i0 - lo:i0 - lo + n] = self.reverse_diffusion(items, n_generated_samples, seed=seed, first_window=i0)["best"]
        return scores.cpu()

    # ------------------------------------------------------------------ host-side tables
    def schedule(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        return schedule(self.N)

    # ------------------------------------------------------------------ measurement
    def launch_count(self) -> int:
        return int(self.lib.mcd_launch_count(self._h))

    def profile_enable(self, on: bool = True) -> None:
        with torch.cuda.device(self.device):
            check(self.lib.mcd_profile_enable(self._h, int(on)))

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        n = self.lib.mcd_profile_slots()
        ms, cnt, win = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_int64 * n)()
        with torch.cuda.device(self.device):
            check(self.lib.mcd_profile_read(self._h, ms, cnt, win))
        res = {}
        for i in range(n):
            b, f = C.c_double(), C.c_double()
            check(self.lib.mcd_profile_slot_cost(self._h, i, C.byref(b), C.byref(f)))
            res[self.lib.mcd_profile_slot_name(i).decode()] = {
                "ms": ms[i], "launches": int(cnt[i]), "windows": int(win[i]),
                "bytes_per_window": b.value, "flops_per_window": f.value}
        return res


def pose_transform_matrices(num_transform: int = 5) -> np.ndarray:
    """Rows 0-1 of the reference's ``ae_trans_list[:num_transform]`` (utils/dataset_utils.py:308-314) as [K,6] float32."""
    lib = _lib.load()
    out = np.empty((num_transform, 6), dtype=np.float32)
    for k in range(num_transform):
        check(lib.mcd_pose_transform_matrix(k, out[k].ctypes.data_as(_lib.c_float_p)))
    return out


def schedule(noise_steps: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(beta, alpha, alpha_hat) fp32 CPU tensors -- models/mocodad.py:799-808 via the library's host function."""
    lib = _lib.load()
    arrs = [np.empty(noise_steps, dtype=np.float32) for _ in range(3)]
    check(lib.mcd_schedule(noise_steps, *[a.ctypes.data_as(_lib.c_float_p) for a in arrs]))
    return tuple(torch.from_numpy(a) for a in arrs)


def pos_encoding(t: int, channels: int) -> torch.Tensor:
    lib = _lib.load()
    a = np.empty(channels, dtype=np.float32)
    check(lib.mcd_pos_encoding(int(t), channels, a.ctypes.data_as(_lib.c_float_p)))
    return torch.from_numpy(a)


def ddpm_coefficients(noise_steps: int, t: int) -> Tuple[float, float, float]:
    lib = _lib.load()
    c = [C.c_float() for _ in range(3)]
    check(lib.mcd_ddpm_coefficients(noise_steps, int(t), *[C.byref(x) for x in c]))
    return tuple(x.value for x in c)


def probe_fp32_tflops(device: int = 0) -> float:
    lib = _lib.load()
    v = C.c_double()
    check(lib.mcd_probe_fp32_tflops(int(device), C.byref(v)))
    return v.value


def probe_fp32_detail(device: int = 0) -> Tuple[float, float]:
    """(scalar FFMA, packed FFMA2) TFLOP/s of the device."""
    lib = _lib.load()
    a, b = C.c_double(), C.c_double()
    check(lib.mcd_probe_fp32_detail(int(device), C.byref(a), C.byref(b)))
    return a.value, b.value
