"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/mocodad_b200.h
declares, its HOST functions match the reference-generated fixtures, and argument errors are
reported through the status/last-error convention.  No GPU compute is attempted here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from mocodad_b200 import _lib, engine
from oracle import ref_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mocodad_b200.h")).read()
    return sorted(set(re.findall(r"MCD_API[^;(]*?\b(mcd_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding and header disagree"
    assert lib.mcd_abi_version() == 5


@pytest.mark.parametrize("N", [2, 10, 50, 1000])
def test_schedule_is_bit_identical_to_reference(lib, golden, N):
    g = golden("schedule")
    beta, alpha, alpha_hat = engine.schedule(N)
    assert np.array_equal(beta.numpy(), g[f"beta_{N}"])
    assert np.array_equal(alpha_hat.numpy(), g[f"alpha_hat_{N}"])
    assert np.array_equal(alpha.numpy(), (1.0 - torch.from_numpy(g[f"beta_{N}"])).numpy())


@pytest.mark.parametrize("N", [2, 10, 1000])
def test_ddpm_coefficients_match_eager_fp32(lib, N):
    beta, alpha, alpha_hat = ref_port.schedule(N)
    for t in sorted({1, 2, N // 2, N - 1} - {0}):
        if t >= N:
            continue
        c1, c2, c3 = engine.ddpm_coefficients(N, t)
        assert c1 == float(1 / torch.sqrt(alpha[t]))
        assert c2 == float((1 - alpha[t]) / torch.sqrt(1 - alpha_hat[t]))
        assert c3 == float(torch.sqrt(beta[t]))


@pytest.mark.parametrize("channels", [16, 8, 64])
def test_pos_encoding_matches_oracle(lib, channels):
    for t in (0, 1, 5, 9, 999):
        want = ref_port.pos_encoding(torch.tensor([[float(t)]]), channels)[0]
        got = engine.pos_encoding(t, channels)
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=2e-6)


def test_shape_support_table(lib):
    assert lib.mcd_shape_supported(3, 3) == 1
    assert lib.mcd_shape_supported(24, 3) == 1
    assert lib.mcd_shape_supported(3, 0) == 1
    assert lib.mcd_shape_supported(6, 3) == 1 and lib.mcd_shape_supported(12, 3) == 1
    assert lib.mcd_shape_supported(6, 6) == 1 and lib.mcd_shape_supported(12, 12) == 1 and lib.mcd_shape_supported(3, 4) == 0
    assert lib.mcd_shape_supported(5, 3) == 0 and lib.mcd_shape_supported(9, 3) == 0


def _cfg(**kw):
    base = dict(n_coords=2, n_joints=17, n_frames=6, n_frames_cond=3, cond_first=1, embedding_dim=16, cond_h_dim=32,
                cond_channels=(C.c_int32 * 3)(32, 16, 32), noise_steps=10, loss_fn=0, device=0)
    base.update(kw)
    return _lib.McdConfig(**base)


def test_argument_errors_are_reported(lib):
    h = C.c_void_p()
    assert lib.mcd_model_create(None, C.byref(h)) == -1
    cfg = _cfg(n_joints=25)
    assert lib.mcd_model_create(C.byref(cfg), C.byref(h)) == -2  # MCD_ERR_UNSUPPORTED, like the reference's einsum error
    assert b"17/12/10" in lib.mcd_last_error()
    cfg = _cfg(n_frames=8)  # T = 5: no kernels compiled
    assert lib.mcd_model_create(C.byref(cfg), C.byref(h)) == -2
    cfg = _cfg(noise_steps=0)
    assert lib.mcd_model_create(C.byref(cfg), C.byref(h)) == -1
    assert lib.mcd_schedule(0, None, None, None) == -1
    with pytest.raises(_lib.McdError):
        _lib.check(lib.mcd_pos_encoding(3, 7, None))


def test_lifecycle_without_weights(lib):
    cfg = _cfg()
    h = C.c_void_p()
    assert lib.mcd_model_create(C.byref(cfg), C.byref(h)) == 0
    try:
        assert lib.mcd_workspace_bytes(h, 4) > 0
        # compute before finalize is refused
        assert lib.mcd_unet_forward(h, None, 0, 1, None, 0, None, None, 0, None) == -3
        arr = np.zeros(4, np.float32)
        assert lib.mcd_model_set_tensor(h, b"model.st_gcnnsp1a.0.prelu.weight", arr.ctypes.data, 1) == 0
        # finalize with an incomplete state_dict names the first missing entry
        assert lib.mcd_model_finalize(h) == -4
        assert b"model.st_gcnnsp1a.0" in lib.mcd_last_error()
        assert lib.mcd_launch_count(h) == 0
    finally:
        lib.mcd_model_destroy(h)


def test_profile_slot_table(lib):
    n = lib.mcd_profile_slots()
    names = [lib.mcd_profile_slot_name(i).decode() for i in range(n)]
    assert names[:11] == list(ref_port.UNET_BLOCKS)
    assert {"down1", "down2", "up3", "up2", "ddpm_step", "window_loss"} <= set(names)


def test_pose_transform_matrices_match_reference(lib):
    """ae_trans_list of the reference (utils/dataset_utils.py:255-270, 308-314), pinned by tests/golden/transforms.npz."""
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "transforms.npz"))["mats"]  # [5,3,3]
    for k in range(5):
        m = np.zeros(6, dtype=np.float32)
        assert lib.mcd_pose_transform_matrix(k, m.ctypes.data_as(_lib.c_float_p)) == 0
        assert np.array_equal(m.reshape(2, 3), gold[k, :2]), k
    assert lib.mcd_pose_transform_matrix(5, m.ctypes.data_as(_lib.c_float_p)) != 0
