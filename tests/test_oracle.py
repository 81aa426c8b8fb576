"""The CPU oracle (oracle/ref_port.py) against the fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  make_golden.py asserted bit-equality on the machine that generated them;
here a small fp32 tolerance allows for a different CPU / BLAS build."""
import numpy as np
import pytest
import torch

from oracle import ref_port, synth

CASES = {"avenue_T3": (6, 10, 3, 6), "plumb_N2": (6, 2, 2, 5), "stress_T24": (27, 10, 2, 3),
         "mid_T6": (9, 10, 2, 3), "mid_T12": (15, 10, 2, 3)}
TOL = dict(rtol=0, atol=5e-6)


def _inputs(name):
    seg_len, N, G, B = CASES[name]
    T = seg_len - 3
    sd = synth.synth_state_dict(synth.state_dict_spec(T=T, T_cond=3), seed=0)
    batch = synth.synth_batch(B, seg_len, seed=1)
    noise = synth.synth_noise(G, N, B, T, seed=2)
    return sd, batch, noise, (seg_len, N, G, B, T)


@pytest.mark.parametrize("name", list(CASES))
def test_fixture_meta(golden, name):
    assert tuple(golden(name)["meta"]) == CASES[name]


@pytest.mark.parametrize("N", [2, 10, 50, 1000])
def test_schedule_matches_reference(golden, N):
    g = golden("schedule")
    beta, alpha, alpha_hat = ref_port.schedule(N)
    assert np.array_equal(beta.numpy(), g[f"beta_{N}"])
    assert np.array_equal(alpha_hat.numpy(), g[f"alpha_hat_{N}"])


@pytest.mark.parametrize("name", list(CASES))
def test_cond_encoder_and_first_denoiser_call(golden, name):
    g = golden(name)
    sd, batch, noise, (seg_len, N, G, B, T) = _inputs(name)
    with torch.no_grad():
        cond, _ = ref_port.select_frames(batch[0], (0, 1, 2))
        emb = ref_port.cond_encode(sd, cond)
        taps = {}
        eps = ref_port.unet_forward(sd, noise[0, 0], torch.full((B,), N - 1, dtype=torch.long), emb, taps=taps)
    np.testing.assert_allclose(emb.numpy(), g["cond_emb"], **TOL)
    np.testing.assert_allclose(eps.numpy(), g["eps_first"], **TOL)
    for k in g.files:
        if k.startswith("tap_"):
            np.testing.assert_allclose(taps[k[4:]].numpy(), g[k], **TOL)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("strategy", ["best", "worst", "mean", "median", "mean_pose", "median_pose", "quantile:0.25", "all"])
def test_reverse_diffusion_losses(golden, name, strategy):
    g = golden(name)
    sd, batch, noise, (seg_len, N, G, B, T) = _inputs(name)
    with torch.no_grad():
        loss, sel = ref_port.reverse_diffusion(sd, batch[0], noise_steps=N, n_generated_samples=G, noise=noise,
                                               strategy=strategy)
    key = "loss_" + strategy.replace(":", "_").replace(".", "p")
    np.testing.assert_allclose(loss.numpy(), g[key], rtol=0, atol=2e-5)
    if strategy == "best":
        np.testing.assert_allclose(sel.numpy(), g["x_sel"], rtol=0, atol=1e-4)


def test_synthetic_checkpoint_is_stable():
    """PCG64 streams keyed by (seed, crc32(name)) regenerate the same weights everywhere."""
    sd = synth.synth_state_dict(synth.state_dict_spec(T=3, T_cond=3), seed=0)
    assert len(sd) == 335
    assert sum(v.numel() for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k) == 142294
    a = sd["model.st_gcnnsd1.0.gcn.A"]
    assert a.shape == (3, 17, 17)
    assert abs(float(a.abs().max())) <= 1 / np.sqrt(17) + 1e-6
