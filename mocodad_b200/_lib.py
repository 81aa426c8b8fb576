"""ctypes binding of the C ABI in ``include/mocodad_b200.h`` -- the stub a maintainer of the
reference would add beneath ``MoCoDAD.forward`` (models/mocodad.py:129-184); see INTEGRATION.md.

There is exactly one implementation of the path: the CUDA library.  If it is missing or cannot
be loaded this module raises; nothing here computes anything.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import _build

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)


class McdConfig(C.Structure):
    """``struct mcd_config`` (include/mocodad_b200.h)."""
    _fields_ = [
        ("n_coords", C.c_int32), ("n_joints", C.c_int32), ("n_frames", C.c_int32),
        ("n_frames_cond", C.c_int32), ("cond_first", C.c_int32), ("embedding_dim", C.c_int32),
        ("cond_h_dim", C.c_int32), ("cond_channels", C.c_int32 * 3), ("noise_steps", C.c_int32),
        ("loss_fn", C.c_int32), ("device", C.c_int32),
        ("latent_dim", C.c_int32), ("n_hidden", C.c_int32), ("hidden", C.c_int32 * 8),
    ]


MCD_OK = 0
ABI_VERSION = 5   # MCD_ABI_VERSION in include/mocodad_b200.h
STATUS_NAMES = {0: "MCD_OK", -1: "MCD_ERR_INVALID_ARG", -2: "MCD_ERR_UNSUPPORTED",
                -3: "MCD_ERR_NOT_FINALIZED", -4: "MCD_ERR_MISSING_TENSOR", -5: "MCD_ERR_CUDA",
                -6: "MCD_ERR_WORKSPACE"}
LOSS_FN = {"smooth_l1": 0, "l1": 1, "mse": 2}

# name -> (restype, argtypes); the single source of truth for tests/test_abi.py as well
SIGNATURES = {
    "mcd_abi_version": (C.c_int, []),
    "mcd_last_error": (C.c_char_p, []),
    "mcd_shape_supported": (C.c_int, [C.c_int32, C.c_int32]),
    "mcd_schedule": (C.c_int, [C.c_int32, c_float_p, c_float_p, c_float_p]),
    "mcd_pos_encoding": (C.c_int, [C.c_int32, C.c_int32, c_float_p]),
    "mcd_ddpm_coefficients": (C.c_int, [C.c_int32, C.c_int32, c_float_p, c_float_p, c_float_p]),
    "mcd_model_create": (C.c_int, [C.POINTER(McdConfig), C.POINTER(C.c_void_p)]),
    "mcd_model_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "mcd_model_finalize": (C.c_int, [C.c_void_p]),
    "mcd_model_destroy": (None, [C.c_void_p]),
    "mcd_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "mcd_cond_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mcd_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "mcd_unet_tap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_char_p,
                               C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mcd_ddpm_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_uint64,
                                C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "mcd_randn_windows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_int32, C.c_int32,
                                    C.c_void_p]),
    "mcd_pose_transform_matrix": (C.c_int, [C.c_int32, c_float_p]),
    "mcd_expand_transforms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, c_float_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p,
                                        C.c_void_p]),
    "mcd_normalize_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, c_double_p, c_double_p, C.c_void_p,
                                       C.c_void_p]),
    "mcd_build_items": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, c_double_p, c_double_p,
                                  c_float_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "mcd_frame_scores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64,
                                   C.c_int64, C.c_void_p, C.c_void_p]),
    "mcd_window_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "mcd_reverse_diffusion": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_uint64,
                                        C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_size_t, C.c_void_p]),
    "mcd_latent_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mcd_latent_denoise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "mcd_latent_reverse_diffusion": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_uint64, C.c_int64,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                               C.c_void_p]),
    "mcd_score_windows_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_int64,
                                         C.c_void_p]),
    "mcd_launch_count": (C.c_int64, [C.c_void_p]),
    "mcd_debug_trace_next": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "mcd_profile_slots": (C.c_int, []),
    "mcd_profile_slot_name": (C.c_char_p, [C.c_int]),
    "mcd_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "mcd_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mcd_profile_slot_cost": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mcd_probe_fp32_tflops": (C.c_int, [C.c_int32, C.POINTER(C.c_double)]),
    "mcd_probe_fp32_detail": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}


class McdError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


_LIB: Optional[C.CDLL] = None


def lib_path() -> str:
    return os.environ.get("MOCODAD_B200_LIB", _build.LIB_PATH)


def load() -> C.CDLL:
    """dlopen the CUDA library (never builds implicitly, never falls back)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m mocodad_b200._build` (needs nvcc). "
            "mocodad_b200 has no CPU or PyTorch fallback for the scoring path.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mcd_abi_version() != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.mcd_abi_version()} != {ABI_VERSION} (stale build?)")
    _LIB = lib
    return lib


def check(status: int) -> None:
    if status != MCD_OK:
        raise McdError(status, load().mcd_last_error().decode("utf-8", "replace"))
