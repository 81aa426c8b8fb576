"""Parity of the CUDA path (through the C ABI) with the reference.

Two anchors: (1) tests/golden/*.npz -- tensors produced by the UNMODIFIED reference
(oracle/make_golden.py) on seeded synthetic weights / windows / noise that oracle/synth.py
regenerates bit-identically here; (2) oracle/ref_port.py run on the host for shapes the fixtures do
not cover.  Tolerances (north_star: 1e-4 fp32 on identical inputs and injected noise):
  single denoiser call / encoder / taps   2e-5 abs   (fp32 re-association + folded BatchNorm)
  x_0 after the N-1 step chain            1e-4 abs   (chain amplifies by prod 1/sqrt(alpha) ~ 201)
  per-window loss                         1e-4 abs
  DDPM update with injected noise         bit-exact
"""
import numpy as np
import pytest
import torch

from oracle import ref_port, synth

pytestmark = pytest.mark.gpu

CASES = {"avenue_T3": (6, 10, 3, 6), "plumb_N2": (6, 2, 2, 5), "stress_T24": (27, 10, 2, 3),
         "mid_T6": (9, 10, 2, 3), "mid_T12": (15, 10, 2, 3)}
DEV = "cuda:0"


def _engine(seg_len, N, loss_fn="smooth_l1", n_cond=3, cond_first=True, sd=None, seed=0):
    from mocodad_b200 import ScoringEngine
    eng = ScoringEngine(seg_len=seg_len, n_frames_cond=n_cond, cond_first=cond_first, noise_steps=N, loss_fn=loss_fn,
                        device=DEV)
    if sd is None:
        sd = synth.synth_state_dict(synth.state_dict_spec(T=seg_len - n_cond, T_cond=n_cond if n_cond else 3,
                                                          conditioning_architecture="AE" if n_cond else None), seed=seed)
    eng.load_state_dict(sd)
    return eng, sd


@pytest.fixture(scope="module")
def setups():
    cache = {}

    def get(name):
        if name not in cache:
            seg_len, N, G, B = CASES[name]
            T = seg_len - 3
            eng, sd = _engine(seg_len, N)
            batch = synth.synth_batch(B, seg_len, seed=1)
            noise = synth.synth_noise(G, N, B, T, seed=2)
            cache[name] = (eng, sd, batch, noise, (seg_len, N, G, B, T))
        return cache[name]
    return get


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name", list(CASES))
def test_cond_encoder_vs_reference(setups, golden, name):
    eng, sd, batch, noise, dims = setups(name)
    emb = eng.cond_encode(batch[0].to(DEV))
    np.testing.assert_allclose(_np(emb), golden(name)["cond_emb"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("name", list(CASES))
def test_first_denoiser_call_vs_reference(setups, golden, name):
    eng, sd, batch, noise, (seg_len, N, G, B, T) = setups(name)
    g = golden(name)
    emb = torch.from_numpy(g["cond_emb"]).to(DEV)
    eps = eng.unet_forward(noise[0, 0].to(DEV).contiguous(), N - 1, emb)
    np.testing.assert_allclose(_np(eps), g["eps_first"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("name", ["avenue_T3", "plumb_N2"])
def test_every_layer_tap_vs_reference(setups, golden, name):
    eng, sd, batch, noise, (seg_len, N, G, B, T) = setups(name)
    g = golden(name)
    emb = torch.from_numpy(g["cond_emb"]).to(DEV)
    x = noise[0, 0].to(DEV).contiguous()
    taps = [k for k in g.files if k.startswith("tap_")]
    assert len(taps) == 15
    for k in taps:
        want = g[k]
        got = eng.unet_tap(x, N - 1, emb, k[4:], want.shape[1], want.shape[3])
        np.testing.assert_allclose(_np(got), want, rtol=0, atol=2e-5, err_msg=k)


def test_every_layer_tap_T24_vs_oracle(setups):
    eng, sd, batch, noise, (seg_len, N, G, B, T) = setups("stress_T24")
    with torch.no_grad():
        cond, _ = ref_port.select_frames(batch[0], (0, 1, 2))
        emb = ref_port.cond_encode(sd, cond)
        taps = {}
        ref_port.unet_forward(sd, noise[0, 0], torch.full((B,), 4, dtype=torch.long), emb, taps=taps)
    x = noise[0, 0].to(DEV).contiguous()
    for k, want in taps.items():
        got = eng.unet_tap(x, 4, emb.to(DEV), k, want.shape[1], want.shape[3])
        np.testing.assert_allclose(_np(got), want.numpy(), rtol=0, atol=2e-5, err_msg=k)


@pytest.mark.parametrize("name", list(CASES))
def test_ddpm_update_is_bit_exact(setups, name):
    eng, sd, batch, noise, (seg_len, N, G, B, T) = setups(name)
    beta, alpha, alpha_hat = ref_port.schedule(N)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(B, 2, T, 17, generator=gen)
    eps = torch.randn(B, 2, T, 17, generator=gen)
    z = torch.randn(B, 2, T, 17, generator=gen)
    for t in range(N - 1, 0, -1):
        tt = torch.full((B,), t, dtype=torch.long)
        zz = z if t > 1 else torch.zeros_like(z)
        want = ref_port.ddpm_update(x, eps, zz, alpha[tt][:, None, None, None], alpha_hat[tt][:, None, None, None],
                                    beta[tt][:, None, None, None])
        got = eng.ddpm_step(x.to(DEV).clone(), eps.to(DEV), t, z.to(DEV))
        assert torch.equal(got.cpu(), want), f"t={t}"


@pytest.mark.parametrize("name", list(CASES))
def test_reverse_diffusion_vs_reference(setups, golden, name):
    eng, sd, batch, noise, (seg_len, N, G, B, T) = setups(name)
    g = golden(name)
    res = eng.reverse_diffusion(batch[0].to(DEV), G, noise=noise.to(DEV), want_losses=True, want_worst=True,
                                want_samples=True)
    np.testing.assert_allclose(_np(res["best"]), g["loss_best"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(_np(res["worst"]), g["loss_worst"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(_np(res["losses"]).T, g["loss_all"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(_np(res["losses"].mean(0)), g["loss_mean"], rtol=0, atol=1e-4)
    # the selected ('best') sample itself
    idx = res["losses"].argmin(0)
    sel = res["x0"][idx, torch.arange(B, device=DEV)]
    np.testing.assert_allclose(_np(sel), g["x_sel"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("strategy", ["best", "worst", "mean", "median", "mean_pose", "median_pose", "quantile:0.25", "all"])
def test_module_forward_all_strategies_vs_reference(golden, strategy):
    """Through the reference-facing surface: MoCoDAD(args).forward(batch, aggr_strategy, return_)."""
    import argparse
    from mocodad_b200 import MoCoDAD
    from test_module import BASE
    name = "avenue_T3"
    seg_len, N, G, B = CASES[name]
    g = golden(name)
    cfg = dict(BASE, seg_len=seg_len, noise_steps=N, n_generated_samples=G, b200_rng="torch")
    m = MoCoDAD(argparse.Namespace(**cfg))
    m.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=3, T_cond=3), seed=0))
    m = m.to(DEV).eval()
    noise = synth.synth_noise(G, N, B, 3, seed=2)
    feed = iter([noise[gi, k] for gi in range(G) for k in range(N - 1)])
    real = torch.randn
    torch.randn = lambda *shape, **kw: next(feed).to(kw.get("device", "cpu"))
    try:
        ret = "loss" if strategy in ("mean", "median", "quantile:0.25") else "all"
        out = m.forward(synth.synth_batch(B, seg_len, seed=1), aggr_strategy=strategy, return_=ret)
    finally:
        torch.randn = real
    key = "loss_" + strategy.replace(":", "_").replace(".", "p")
    np.testing.assert_allclose(_np(out[0]), g[key], rtol=0, atol=1e-4)
    if strategy == "best":
        np.testing.assert_allclose(_np(out[1]), g["x_sel"], rtol=0, atol=1e-4)
    assert len(out) == (5 if ret == "loss" else 6)
    assert out[-4].shape == (B, 2, seg_len, 17)


@pytest.mark.parametrize("loss_fn", ["l1", "mse", "smooth_l1"])
def test_loss_functions_vs_oracle(loss_fn):
    eng, sd = _engine(6, 10, loss_fn=loss_fn)
    batch = synth.synth_batch(9, 6, seed=4)
    gen = torch.Generator().manual_seed(11)
    x0 = 2.0 * torch.randn(2, 9, 2, 3, 17, generator=gen)  # |d| straddles the SmoothL1 knee
    res = eng.window_loss(x0.to(DEV), batch[0].to(DEV), G=2)
    _, corrupt = ref_port.select_frames(batch[0], (0, 1, 2))
    want = torch.stack([ref_port.window_loss(x0[g], corrupt, loss_fn) for g in range(2)])
    np.testing.assert_allclose(_np(res["losses"]), want.numpy(), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(_np(res["best"]), want.min(0)[0].numpy(), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(_np(res["worst"]), want.max(0)[0].numpy(), rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("seg_len,n_cond,cond_first", [(6, 3, False), (3, 0, True), (24, 0, True), (12, 6, True), (24, 12, True),
                                                        (12, 6, False)])
def test_other_conditioning_layouts_vs_oracle(seg_len, n_cond, cond_first):
    """Conditioning on the LAST frames, and 'no_condition' (whole window denoised, no encoder)."""
    N, G, B = 4, 2, 5
    T = seg_len - n_cond
    eng, sd = _engine(seg_len, N, n_cond=n_cond, cond_first=cond_first)
    batch = synth.synth_batch(B, seg_len, seed=7)
    noise = synth.synth_noise(G, N, B, T, seed=8)
    idx = () if n_cond == 0 else (tuple(range(n_cond)) if cond_first else tuple(range(seg_len - n_cond, seg_len)))
    with torch.no_grad():
        want, _ = ref_port.reverse_diffusion(sd, batch[0], noise_steps=N, n_generated_samples=G, noise=noise,
                                             conditioning_indices=idx, inject=n_cond > 0)
    res = eng.reverse_diffusion(batch[0].to(DEV), G, noise=noise.to(DEV))
    np.testing.assert_allclose(_np(res["best"]), want.numpy(), rtol=0, atol=1e-4)


@pytest.mark.parametrize("B", [1, 7, 8, 9, 149])
def test_ragged_batch_sizes_vs_oracle(B):
    """Batches that do not fill a CTA tile (8 windows at T=3), incl. a single window."""
    N, G = 3, 1
    eng, sd = _engine(6, N)
    batch = synth.synth_batch(B, 6, seed=B)
    noise = synth.synth_noise(G, N, B, 3, seed=B + 1)
    with torch.no_grad():
        want, _ = ref_port.reverse_diffusion(sd, batch[0], noise_steps=N, n_generated_samples=G, noise=noise)
    res = eng.reverse_diffusion(batch[0].to(DEV), G, noise=noise.to(DEV))
    np.testing.assert_allclose(_np(res["best"]), want.numpy(), rtol=0, atol=1e-4)


def test_empty_batch_and_bad_arguments():
    from mocodad_b200 import _lib
    eng, sd = _engine(6, 10)
    res = eng.reverse_diffusion(torch.empty(0, 2, 6, 17, device=DEV), 3)
    assert res["best"].shape == (0,)
    with pytest.raises(ValueError):
        eng.reverse_diffusion(torch.zeros(4, 2, 6, 17), 3)  # CPU tensor
    with pytest.raises(ValueError):
        eng.reverse_diffusion(torch.zeros(4, 2, 7, 17, device=DEV), 3)  # wrong seg_len
    with pytest.raises(_lib.McdError):
        eng.unet_forward(torch.zeros(4, 2, 3, 17, device=DEV), 10, None)  # t outside the schedule
    with pytest.raises(_lib.McdError):
        eng.unet_tap(torch.zeros(4, 2, 3, 17, device=DEV), 1, None, "nope", 2, 17)


def test_tiling_of_the_virtual_batch_is_invisible(setups):
    """Processing the G*B virtual batch in small workspace tiles gives bit-identical scores."""
    eng, sd, batch, noise, (seg_len, N, G, B, T) = setups("avenue_T3")
    d, nz = batch[0].to(DEV), noise.to(DEV)
    whole = eng.reverse_diffusion(d, G, noise=nz, want_losses=True)
    for tile in (1, 5, 8):
        part = eng.reverse_diffusion(d, G, noise=nz, want_losses=True, tile_windows=tile)
        assert torch.equal(part["losses"], whole["losses"]), tile
        assert torch.equal(part["best"], whole["best"]), tile


def test_philox_noise_is_standard_normal_and_keyed():
    eng, sd = _engine(27, 10)
    x = eng.randn_windows(4096, seed=999, first_window=0, sample=0, noise_slot=0)
    v = x.flatten().double()
    assert abs(float(v.mean())) < 3e-3
    assert abs(float(v.var()) - 1.0) < 5e-3
    assert abs(float((v ** 4).mean()) - 3.0) < 5e-2          # kurtosis of N(0,1)
    assert abs(float((v.abs() > 1.959964).double().mean()) - 0.05) < 2e-3
    # keyed by (seed, window, sample, slot): a shifted window range reproduces the overlapping windows
    y = eng.randn_windows(4096, seed=999, first_window=1000, sample=0, noise_slot=0)
    assert torch.equal(y[:3096], x[1000:])
    for kw in (dict(seed=1000), dict(sample=1), dict(noise_slot=1)):
        args = dict(seed=999, first_window=0, sample=0, noise_slot=0)
        args.update(kw)
        z = eng.randn_windows(4096, **args)
        assert abs(float((z.flatten() * x.flatten()).mean())) < 5e-3  # independent streams
    # adjacent windows / elements are uncorrelated
    assert abs(float((x[:-1] * x[1:]).mean())) < 5e-3


def test_philox_scores_do_not_depend_on_batching_or_rank_count():
    """Scoring windows [0,B) in one call or as two 'ranks' with first_window offsets is bit-identical."""
    eng, sd = _engine(6, 10)
    B, G = 37, 4
    d = synth.synth_batch(B, 6, seed=3)[0].to(DEV)
    whole = eng.reverse_diffusion(d, G, seed=999, first_window=100)["best"]
    lo = eng.reverse_diffusion(d[:20].contiguous(), G, seed=999, first_window=100)["best"]
    hi = eng.reverse_diffusion(d[20:].contiguous(), G, seed=999, first_window=120)["best"]
    assert torch.equal(torch.cat([lo, hi]), whole)
    again = eng.reverse_diffusion(d, G, seed=999, first_window=100)["best"]
    assert torch.equal(again, whole)  # deterministic
    other = eng.reverse_diffusion(d, G, seed=1000, first_window=100)["best"]
    assert not torch.equal(other, whole)


def test_host_entry_matches_device_entry():
    eng, sd = _engine(6, 10)
    B, G = 33, 3
    data = synth.synth_batch(B, 6, seed=12)[0]
    dev = eng.reverse_diffusion(data.to(DEV), G, seed=7, first_window=5)["best"]
    host = eng.score_windows_host(data.pin_memory(), G, seed=7, first_window=5)
    assert host.device.type == "cpu"
    assert torch.equal(host, dev.cpu())


def test_full_size_batch_properties():
    """BASELINE.json sizes (B=1024 windows of [2,24,17], N=10): properties that need no oracle run.
    best-of-G is the minimum of the per-sample losses; permuting the windows permutes the scores;
    anomalous (shuffled-joint) windows score higher than their source windows on average."""
    eng, sd = _engine(27, 10)
    B, G = 1024, 2
    data = synth.synth_batch(B, 27, seed=21)[0].to(DEV)
    res = eng.reverse_diffusion(data, G, seed=1, want_losses=True, want_worst=True)
    assert torch.isfinite(res["losses"]).all()
    assert torch.equal(res["best"], res["losses"].min(0)[0])
    assert torch.equal(res["worst"], res["losses"].max(0)[0])
    # injected-noise run is permutation-equivariant
    noise = torch.randn(1, 9, 64, 2, 24, 17, device=DEV)
    perm = torch.randperm(64, device=DEV)
    a = eng.reverse_diffusion(data[:64].contiguous(), 1, noise=noise)["best"]
    b = eng.reverse_diffusion(data[:64][perm].contiguous(), 1, noise=noise[:, :, perm].contiguous())["best"]
    assert torch.equal(a[perm], b)


def test_launch_counter_and_profile_slots():
    eng, sd = _engine(6, 4)
    d = synth.synth_batch(16, 6, seed=2)[0].to(DEV)
    n0 = eng.launch_count()
    eng.profile_enable(True)
    eng.reverse_diffusion(d, 2, seed=3)
    prof = eng.profile_read()
    eng.profile_enable(False)
    launched = eng.launch_count() - n0
    # encoder 4+1, per tile: randn + 3 steps x (time embedding + 11 blocks + down1 + down2; up3 / up2 are fused into the blocks
    # before them and the DDPM update into the last block) + loss, + best
    assert launched == 5 + 1 + 3 * 14 + 1 + 1
    assert sum(v["launches"] for v in prof.values()) == launched
    assert prof["st_gcnnsd3.0"]["launches"] == 3 and prof["st_gcnnsd3.0"]["windows"] == 3 * 32
    assert all(v["ms"] > 0 for v in prof.values() if v["launches"])


def test_auc_parity_through_post_processing():
    """End of the path: per-window scores -> score assembly -> frame-level AUC (mocodad.py:337-430, restated in
    mocodad_b200/postproc.py and pinned to the reference by tests/test_postproc.py).  north_star: AUC within +-0.1
    (percentage points) of the reference path on identical inputs and noise."""
    from mocodad_b200 import postproc, synthetic
    clips = {(1, 1): 140, (1, 2): 120}
    _, trans, meta, frames, gt = synthetic.synth_scored_dataset(clips, num_transform=2, persons_per_clip=2, seed=9)
    n = len(trans)
    N, G = 10, 3
    eng, sd = _engine(6, N)
    data = synth.synth_batch(n, 6, seed=31)[0]
    # windows inside anomalous frames get jittered joints, so the score carries some signal
    lab = np.array([gt[(int(m[0]), int(m[1]))][int(f[0]) - 1:int(f[-1])].mean() for m, f in zip(meta, frames)])
    data = data + torch.from_numpy((lab[:, None, None, None] * np.random.default_rng(3).normal(0, 1.5, data.shape)).astype(np.float32))
    noise = synth.synth_noise(G, N, n, 3, seed=32)
    with torch.no_grad():
        want, _ = ref_port.reverse_diffusion(sd, data, noise_steps=N, n_generated_samples=G, noise=noise)
    got = eng.reverse_diffusion(data.to(DEV), G, noise=noise.to(DEV))["best"].cpu()
    kw = dict(num_transform=2, pad_size=-1, frames_shift=2, filter_kernel_size=3)
    auc_ref = postproc.dataset_auc(want.numpy(), trans, meta, frames, gt, **kw)
    auc_got = postproc.dataset_auc(got.numpy(), trans, meta, frames, gt, **kw)
    assert abs(auc_got - auc_ref) <= 1e-3, (auc_got, auc_ref)          # 0.1 percentage points
    # in-kernel Philox noise: a different but equally distributed stream -> statistically the same detector
    phil = eng.reverse_diffusion(data.to(DEV), G, seed=999)["best"].cpu()
    auc_phil = postproc.dataset_auc(phil.numpy(), trans, meta, frames, gt, **kw)
    assert abs(auc_phil - auc_ref) <= 0.03, (auc_phil, auc_ref)


EXTRA_T = [6, 12]   # further frame counts the library carries kernels for (MCD_FOR_EACH_T); no reference fixture: oracle on the fly


@pytest.mark.parametrize("T", EXTRA_T)
def test_every_layer_tap_other_frame_counts_vs_oracle(T):
    seg_len, N, B = T + 3, 10, 5
    eng, sd = _engine(seg_len, N)
    batch = synth.synth_batch(B, seg_len, seed=21)
    x0 = synth.synth_noise(1, N, B, T, seed=22)[0, 0]
    with torch.no_grad():
        cond, _ = ref_port.select_frames(batch[0], (0, 1, 2))
        emb = ref_port.cond_encode(sd, cond)
        taps = {}
        eps = ref_port.unet_forward(sd, x0, torch.full((B,), 7, dtype=torch.long), emb, taps=taps)
    np.testing.assert_allclose(_np(eng.cond_encode(batch[0].to(DEV))), emb.numpy(), rtol=0, atol=2e-5)
    x = x0.to(DEV).contiguous()
    for k, want in taps.items():
        got = eng.unet_tap(x, 7, emb.to(DEV), k, want.shape[1], want.shape[3])
        np.testing.assert_allclose(_np(got), want.numpy(), rtol=0, atol=2e-5, err_msg=f"T={T} {k}")
    got = eng.unet_forward(x, 7, emb.to(DEV))
    np.testing.assert_allclose(_np(got), eps.numpy(), rtol=0, atol=2e-5)


@pytest.mark.parametrize("T", EXTRA_T)
def test_reverse_diffusion_other_frame_counts_vs_oracle(T):
    seg_len, N, G, B = T + 3, 10, 2, 37
    eng, sd = _engine(seg_len, N)
    batch = synth.synth_batch(B, seg_len, seed=31)
    noise = synth.synth_noise(G, N, B, T, seed=32)
    with torch.no_grad():
        want, _ = ref_port.reverse_diffusion(sd, batch[0], noise_steps=N, n_generated_samples=G, noise=noise)
    res = eng.reverse_diffusion(batch[0].to(DEV), G, noise=noise.to(DEV))
    np.testing.assert_allclose(_np(res["best"]), want.numpy(), rtol=0, atol=1e-4)
    # Philox mode: keyed by the global window index, invariant to how the batch is cut
    whole = eng.reverse_diffusion(batch[0].to(DEV), G, seed=5)["best"]
    parts = torch.cat([eng.reverse_diffusion(batch[0][a:b].to(DEV).contiguous(), G, seed=5, first_window=a)["best"]
                       for a, b in ((0, 11), (11, 37))])
    assert torch.equal(whole, parts)


# ---- the regime bench.py times: persistent CTAs that run MANY tiles each (ring slots it % NXB, mbarrier parity bits
# (it / NXB) & 1, TMEM accumulator-set ping-pong, residual look-ahead), compared per window with the CPU oracle.
# T=24: one window per CTA tile, 148 CTAs -> G*B = 1 200 windows are 8.1 tiles per CTA; T=3: 8 windows per tile ->
# G*B = 10 000 windows are 8.4 tiles per CTA.  `tile` additionally cuts the virtual batch into >= 3 passes.
@pytest.mark.parametrize("seg_len,B,G,N,tile", [(27, 600, 2, 3, None), (27, 600, 2, 3, 444), (6, 2000, 5, 3, None),
                                                  (6, 2000, 5, 3, 3552), (15, 700, 2, 3, None)])
def test_many_tiles_per_cta_vs_oracle(seg_len, B, G, N, tile):
    T = seg_len - 3
    eng, sd = _engine(seg_len, N)
    batch = synth.synth_batch(B, seg_len, seed=51)
    noise = synth.synth_noise(G, N, B, T, seed=52)
    with torch.no_grad():
        want, _, gen = ref_port.reverse_diffusion(sd, batch[0], noise_steps=N, n_generated_samples=G, noise=noise, return_samples=True)
        _, corrupt = ref_port.select_frames(batch[0], (0, 1, 2))
        want_all = torch.stack([ref_port.window_loss(x, corrupt) for x in gen])
    res = eng.reverse_diffusion(batch[0].to(DEV), G, noise=noise.to(DEV), want_losses=True, want_samples=True, tile_windows=tile)
    np.testing.assert_allclose(_np(res["x0"]), torch.stack(gen).numpy(), rtol=0, atol=1e-4)      # every element of every window
    np.testing.assert_allclose(_np(res["losses"]), want_all.numpy(), rtol=0, atol=1e-4)
    np.testing.assert_allclose(_np(res["best"]), want.numpy(), rtol=0, atol=1e-4)
    # and bit-identical to the same windows scored in one-tile-per-CTA launches (the regime the small fixtures cover)
    unit = 148 * max(1, 408 // (T * 17))
    small = eng.reverse_diffusion(batch[0].to(DEV), G, noise=noise.to(DEV), want_losses=True, tile_windows=unit)
    assert torch.equal(small["losses"], res["losses"])


@pytest.mark.parametrize("T,n", [(24, 1200), (3, 9600)])
def test_layer_taps_with_many_tiles_per_cta_vs_oracle(T, n):
    """Block-level taps at that size, one of every tensor-core kernel flavour: residual convolution + N-merged MMAs (sd3.1,
    su4.1), residual convolution with single Xlo / Y2 buffers (sd1.0 at V=17), identity residual (sd2.1), wide output (sd3.0),
    plus the joint resamples and the first / last (edge) blocks through the final eps."""
    seg_len, N = T + 3, 10
    eng, sd = _engine(seg_len, N)
    batch = synth.synth_batch(n, seg_len, seed=61)
    x0 = synth.synth_noise(1, N, n, T, seed=62)[0, 0]
    with torch.no_grad():
        cond, _ = ref_port.select_frames(batch[0], (0, 1, 2))
        emb = ref_port.cond_encode(sd, cond)
        taps = {}
        eps = ref_port.unet_forward(sd, x0, torch.full((n,), 6, dtype=torch.long), emb, taps=taps)
    x, demb = x0.to(DEV).contiguous(), emb.to(DEV)
    for k in ("st_gcnnsd1.0", "st_gcnnsd2.1", "st_gcnnsd3.0", "st_gcnnsd3.1", "st_gcnnsu4.1", "st_gcnnsu3.0", "down2", "up3"):
        want = taps[k]
        got = eng.unet_tap(x, 6, demb, k, want.shape[1], want.shape[3])
        np.testing.assert_allclose(_np(got), want.numpy(), rtol=0, atol=2e-5, err_msg=f"T={T} {k}")
    got_eps = eng.unet_forward(x, 6, demb)
    np.testing.assert_allclose(_np(got_eps), eps.numpy(), rtol=0, atol=2e-5)
    # the production call fuses the up-path CNN_layers (up3, up2) into blocks sd3.1 / su4.1 and adds their output onto the skip
    # tensors in place (vector reductions in L2); the tap call runs stand-alone kernels: same operation order and a single
    # (commutative) rounding for the skip add, so the two paths must agree bit for bit
    own = eng.unet_tap(x, 6, demb, "st_gcnnsu3.1", 2, 17)
    assert torch.equal(got_eps, own + x)


def test_frame_scores_kernel_vs_host_and_auc_bit_identity(golden):
    """SURVEY.md 8 row f2, first stage on the device (mcd_frame_scores): per-row frame maxima over windows == the pinned host
    restatement bit for bit (wrapping frame number 0, skipped rows, overlapping windows, empty rows), and the AUC through
    ``dataset_auc`` with the kernel as accelerator == the all-host AUC exactly == the reference's."""
    from mocodad_b200 import postproc, synthetic
    eng, _ = _engine(6, 4)
    rng = np.random.default_rng(5)
    N, L, rows = 20000, 6, 37
    row_len = rng.integers(8, 300, size=rows).astype(np.int32)
    row_len[5] = 6
    row = rng.integers(-1, rows, size=N).astype(np.int64)
    row[row == 11] = 12                                   # row 11 stays empty
    start = rng.integers(0, 10_000, size=N)
    frames = np.empty((N, L), dtype=np.int64)
    for n in range(N):
        ln = int(row_len[max(row[n], 0)])
        s0 = int(start[n] % max(ln - L + 1, 1))
        frames[n] = np.arange(s0, s0 + L) + (1 if n % 50 else 0)   # every 50th window starts at frame number 0 (numpy index -1)
    loss = rng.random(N, dtype=np.float32) * 3
    loss[::97] = 0.0
    got = eng.frame_scores_host(loss, frames, row, row_len, int(row_len.max()))
    for r in range(rows):
        sel = row == r
        want = postproc.person_frame_scores(loss[sel], frames[sel], int(row_len[r])).astype(np.float32)
        assert np.array_equal(got[r, :row_len[r]], want), r
        assert not got[r, row_len[r]:].any()
    assert not got[11].any()
    # through the whole tail, on the fixture epochs the reference's AUCs are pinned on
    ref = golden("postproc")
    from test_postproc import CASES as PCASES
    for name, (clips, dataset, pad, shift, ksize, ntr) in PCASES.items():
        out, trans, meta, frames2, gt = synthetic.synth_scored_dataset(clips, num_transform=ntr)
        kw = dict(num_transform=ntr, pad_size=pad, frames_shift=shift, filter_kernel_size=ksize,
                  avenue_masks=postproc.avenue_hr_mask() if dataset == "HR-Avenue" else None)
        host = postproc.dataset_auc(out, trans, meta, frames2, gt, **kw)
        dev = postproc.dataset_auc(out, trans, meta, frames2, gt, frame_scores=eng.frame_scores_host, **kw)
        assert host == dev and abs(dev - float(ref[name])) < 1e-9, name


def test_seeded_torch_rng_on_cuda_reproduces_the_eager_reference_draws():
    """SURVEY.md 8 row a10, "identical inputs and seeds": with `b200_rng: torch` the module draws its noise with torch.randn on
    the CUDA generator in the reference's call order (mocodad.py:162,176), so under one torch.manual_seed it scores with exactly
    the noise the reference's randn_like calls would see on this GPU.  Reference arm: the oracle port's ATen operators as CUDA
    eager (TF32 off -- cuDNN would otherwise round the 1x1 convolutions to 10 mantissa bits), randn_like = torch.randn_like."""
    import argparse
    from mocodad_b200 import MoCoDAD
    from test_module import BASE
    seg_len, N, G, B = 6, 10, 4, 300
    sd = synth.synth_state_dict(synth.state_dict_spec(T=3, T_cond=3), seed=0)
    m = MoCoDAD(argparse.Namespace(**dict(BASE, seg_len=seg_len, noise_steps=N, n_generated_samples=G, b200_rng="torch")))
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    batch = synth.synth_batch(B, seg_len, seed=71)
    torch.manual_seed(20240607)
    got = m.forward(batch)[0]
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(20240607)
        with torch.no_grad(), torch.device(DEV):   # the port creates its schedule / index tensors on the default device
            want, _ = ref_port.reverse_diffusion({k: v.to(DEV) for k, v in sd.items()}, batch[0].to(DEV), noise_steps=N,
                                                 n_generated_samples=G, randn_like=torch.randn_like)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    np.testing.assert_allclose(_np(got), _np(want), rtol=0, atol=1e-4)
    torch.manual_seed(1)
    assert not torch.allclose(m.forward(batch)[0], got, atol=1e-3)     # another seed, another noise


@pytest.mark.parametrize("seg_len,B", [(6, 16), (27, 4)])
def test_thousand_step_chain_vs_oracle(seg_len, B):
    """BASELINE.json configs[4] runs noise_steps = 1000 (999 denoiser calls per sample; prod 1/sqrt(alpha) ~ 2e4).  Measured on
    the reference itself (CPU, fp32 vs fp64 arithmetic, same inputs): max |dx0| 7e-6, max |dloss| 7e-7 -- the denoiser is
    contractive, the worst-case amplification does not materialise -- so the 1e-4 budget is kept at N = 1000 too."""
    T, N = seg_len - 3, 1000
    eng, sd = _engine(seg_len, N)
    batch = synth.synth_batch(B, seg_len, seed=81)
    noise = synth.synth_noise(1, N, B, T, seed=82)
    with torch.no_grad():
        want, _, gen = ref_port.reverse_diffusion(sd, batch[0], noise_steps=N, n_generated_samples=1, noise=noise, return_samples=True)
    res = eng.reverse_diffusion(batch[0].to(DEV), 1, noise=noise.to(DEV), want_samples=True)
    dx = float((res["x0"][0].cpu() - gen[0]).abs().max())
    dl = float((res["best"].cpu() - want).abs().max())
    print(f"N=1000 T={T}: max |x0 - oracle| = {dx:.3e}, max |loss - oracle| = {dl:.3e}")
    assert dx <= 1e-4 and dl <= 1e-4, (dx, dl)
