"""GPU parity of the LATENT variant (models/mocodad_latent.py, stage 'diffusion') through the C ABI:
`mcd_latent_encode` / `mcd_latent_denoise` / `mcd_latent_reverse_diffusion` against tests/golden/latent_T3.npz -- tensors the
UNMODIFIED reference MoCoDADlatent produced (oracle/make_latent_golden.py) -- and against oracle/latent_port.py at sizes
where every persistent CTA of the MLP kernel carries several tiles.  Tolerances: single calls 2e-5, losses after the
N-1-step chain 1e-4 (north_star), as for the pose-space path."""
import argparse
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import latent_port, ref_port, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(golden):
    g = golden("latent_T3")
    seg_len, N, G, B, latent = (int(v) for v in g["meta"][:5])
    hidden = [int(v) for v in g["meta"][5:]]
    spec = OrderedDict((str(n), tuple(int(d) for d in s.split(",")) if s else ()) for n, s in zip(g["spec_names"], g["spec_shapes"]))
    sd = synth.synth_state_dict(spec, seed=0)
    return g, sd, (seg_len, N, G, B, latent, hidden)


def _engine(sd, seg_len, N, latent, hidden, loss_fn="smooth_l1"):
    from mocodad_b200 import ScoringEngine
    eng = ScoringEngine(seg_len=seg_len, n_frames_cond=3, noise_steps=N, loss_fn=loss_fn, device=DEV, latent_embedding_dim=latent,
                        hidden_sizes=hidden)
    eng.load_state_dict(sd)
    return eng


def _np(t):
    return t.detach().cpu().numpy()


def test_latent_code_tap_and_denoiser_call_vs_reference(golden):
    g, sd, (seg_len, N, G, B, latent, hidden) = _setup(golden)
    eng = _engine(sd, seg_len, N, latent, hidden)
    batch = synth.synth_batch(B, seg_len, seed=1)
    emb, code = eng.latent_encode(batch[0].to(DEV))
    with torch.no_grad():
        cond, corrupt = ref_port.select_frames(batch[0], (0, 1, 2))
        want_emb = ref_port.cond_encode(sd, cond)
    np.testing.assert_allclose(_np(emb), want_emb.numpy(), rtol=0, atol=2e-5)
    np.testing.assert_allclose(_np(code), g["latent_code"], rtol=0, atol=2e-5)
    # the last block of the down half at the encoder's constant step t = -1
    tap = eng.unet_tap(corrupt.contiguous().to(DEV), -1, emb, "st_gcnnsd3.1", 64, 10)
    np.testing.assert_allclose(_np(tap), g["tap_sd3_1"], rtol=0, atol=2e-5)
    noise = torch.from_numpy(g["noise"])
    eps = eng.latent_denoise(noise[0, 0].to(DEV).contiguous(), 7, emb)
    np.testing.assert_allclose(_np(eps), g["eps_t7"], rtol=0, atol=1e-5)


def test_latent_reverse_diffusion_vs_reference(golden):
    g, sd, (seg_len, N, G, B, latent, hidden) = _setup(golden)
    eng = _engine(sd, seg_len, N, latent, hidden)
    batch = synth.synth_batch(B, seg_len, seed=1)
    noise = torch.from_numpy(g["noise"]).to(DEV)
    res = eng.latent_reverse_diffusion(batch[0].to(DEV), G, noise=noise, want_losses=True, want_worst=True, want_samples=True)
    np.testing.assert_allclose(_np(res["best"]), g["loss_best"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(_np(res["losses"].mean(0)), g["loss_mean"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(_np(res["losses"].median(0)[0]), g["loss_median"], rtol=0, atol=1e-4)
    assert torch.equal(res["best"], res["losses"].min(0)[0]) and torch.equal(res["worst"], res["losses"].max(0)[0])
    idx = res["losses"].argmin(0)
    # the latent vectors themselves are O(100): 1e-4 absolute on the loss, the same relative accuracy on the vectors
    np.testing.assert_allclose(_np(res["x0"][idx, torch.arange(B, device=DEV)]), g["latent_sel"], rtol=2e-6, atol=1e-4)
    np.testing.assert_allclose(_np(res["code"]), g["latent_code"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("strategy", ["best", "mean", "median"])
def test_module_surface_vs_reference(golden, strategy):
    """MoCoDADlatent(args).forward(batch, aggr_strategy=..., return_=...) -- the call eval_MoCoDAD.py's Trainer makes."""
    from mocodad_b200 import MoCoDADlatent
    from test_latent_oracle import _latent_args
    g, sd, (seg_len, N, G, B, latent, hidden) = _setup(golden)
    m = MoCoDADlatent(_latent_args(seg_len=seg_len, noise_steps=N, n_generated_samples=G, b200_rng="torch",
                                   latent_embedding_dim=latent, hidden_sizes=hidden))
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    noise = torch.from_numpy(g["noise"])
    feed = iter([noise[gi, k] for gi in range(G) for k in range(N - 1)])
    real = torch.randn
    torch.randn = lambda *shape, **kw: next(feed).to(kw.get("device", "cpu"))
    try:
        out = m.forward(synth.synth_batch(B, seg_len, seed=1), aggr_strategy=strategy, return_="all" if strategy == "best" else "loss")
    finally:
        torch.randn = real
    np.testing.assert_allclose(_np(out[0]), g["loss_" + strategy], rtol=0, atol=1e-4)
    if strategy == "best":
        np.testing.assert_allclose(_np(out[1]), g["latent_sel"], rtol=2e-6, atol=1e-4)
        assert len(out) == 6
    assert out[-4].shape == (B, 2, seg_len, 17)


@pytest.mark.parametrize("B,G,N", [(301, 20, 4), (33, 3, 10), (1, 1, 2)])
def test_many_tiles_per_cta_vs_oracle(golden, B, G, N):
    """B*G up to 6 020 latent vectors = 189 tiles of 32 on <= 148 persistent CTAs (ragged last tile), vs the CPU oracle."""
    g, sd, (seg_len, _, _, _, latent, hidden) = _setup(golden)
    eng = _engine(sd, seg_len, N, latent, hidden)
    batch = synth.synth_batch(B, seg_len, seed=40 + B)
    gen = torch.Generator().manual_seed(B)
    noise = torch.randn(G, max(N - 1, 1), B, latent, generator=gen)
    with torch.no_grad():
        want, _, code = latent_port.latent_reverse_diffusion(sd, batch[0], noise_steps=N, n_generated_samples=G, n_layers=len(hidden),
                                                             noise=noise)
    res = eng.latent_reverse_diffusion(batch[0].to(DEV), G, noise=noise.to(DEV))
    np.testing.assert_allclose(_np(res["code"]), code.numpy(), rtol=0, atol=2e-5)
    np.testing.assert_allclose(_np(res["best"]), want.numpy(), rtol=0, atol=1e-4)


def test_philox_mode_is_batch_invariant_and_deterministic(golden):
    g, sd, (seg_len, N, G, B, latent, hidden) = _setup(golden)
    eng = _engine(sd, seg_len, N, latent, hidden)
    d = synth.synth_batch(70, seg_len, seed=3)[0].to(DEV)
    whole = eng.latent_reverse_diffusion(d, 4, seed=999, first_window=10)["best"]
    lo = eng.latent_reverse_diffusion(d[:33].contiguous(), 4, seed=999, first_window=10)["best"]
    hi = eng.latent_reverse_diffusion(d[33:].contiguous(), 4, seed=999, first_window=43)["best"]
    assert torch.equal(torch.cat([lo, hi]), whole)
    assert torch.equal(eng.latent_reverse_diffusion(d, 4, seed=999, first_window=10)["best"], whole)
    assert not torch.equal(eng.latent_reverse_diffusion(d, 4, seed=1000, first_window=10)["best"], whole)
    assert torch.isfinite(whole).all()


def test_pose_space_calls_are_refused_on_a_latent_handle(golden):
    from mocodad_b200 import _lib
    g, sd, (seg_len, N, G, B, latent, hidden) = _setup(golden)
    eng = _engine(sd, seg_len, N, latent, hidden)
    d = synth.synth_batch(4, seg_len, seed=3)[0].to(DEV)
    with pytest.raises(_lib.McdError):
        eng.reverse_diffusion(d, 2)
    with pytest.raises(_lib.McdError):
        eng.unet_forward(torch.zeros(4, 2, 3, 17, device=DEV), 1, None)
    with pytest.raises(_lib.McdError):
        eng.unet_tap(torch.zeros(4, 2, 3, 17, device=DEV), 1, None, "st_gcnnsu4.0", 64, 12)   # up half: not on this handle
