"""Generate tests/golden/*.npz by executing the UNMODIFIED reference, and pin oracle/ref_port.py.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/, asserts port == reference

TEST INFRASTRUCTURE ONLY.  The reference has no tests or golden vectors,
so parity is pinned by running it.  ``pytorch_lightning`` and ``matplotlib`` are not installed
here; both are replaced by inert stubs *in this process only*  -- the reference's
arithmetic (models/*, utils/diffusion_utils.py) runs unmodified.

What is recorded per case (all fp32, produced by the reference's own code):
  cond_emb   MoCoDAD._encode_condition(cond)[0]                          mocodad.py:546-560
  eps_first  MoCoDAD._unet_forward(x_T, t=N-1, cond_emb)                 mocodad.py:811-840
  taps_*     every ST_GCNN_layer / CNN_layer output inside that call (forward hooks)
  loss       MoCoDAD.forward(batch)[0] with torch.randn_like replaced by the pre-drawn
             oracle/synth.synth_noise tensors, in call order                mocodad.py:129-184
  x_sel      the 'best' sample (forward(..., return_='all')[1])
  loss_<s>   loss under the other aggregation strategies                   mocodad.py:487-516
and the script asserts that oracle/ref_port.py reproduces each of them BIT-EXACTLY on this
machine, plus reference-vs-port equality under a shared torch.manual_seed (call-order check).
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("MOCODAD_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import ref_port, synth  # noqa: E402


def install_stubs() -> None:
    """Inert stand-ins for the two absent imports (mocodad.py:7, eval_utils.py:5)."""
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        def on_test_epoch_start(self):
            pass

        def on_validation_epoch_start(self):
            pass

    pl.LightningModule = LightningModule
    sys.modules["pytorch_lightning"] = pl
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def load_reference():
    install_stubs()
    sys.path.insert(0, REF)
    from models.mocodad import MoCoDAD  # type: ignore
    return MoCoDAD


def make_args(seg_len: int, noise_steps: int, n_gen: int, aggr: str = "best") -> argparse.Namespace:
    cfg = yaml.load(open(os.path.join(REF, "config/Avenue/mocodad_test.yaml")), Loader=yaml.FullLoader)
    cfg.update(seg_len=seg_len, noise_steps=noise_steps, n_generated_samples=n_gen,
               aggregation_strategy=aggr, gt_path="", ckpt_dir="", save_tensors=False)
    return argparse.Namespace(**cfg)


class NoiseFeeder:
    """Replaces torch.randn_like inside the reference's forward by pre-drawn tensors."""

    def __init__(self, noise: torch.Tensor, noise_steps: int):
        self.noise, self.n, self.calls = noise, noise_steps, 0

    def __call__(self, like, **kw):
        per_g = max(self.n - 1, 1)
        g, k = divmod(self.calls, per_g)
        self.calls += 1
        out = self.noise[g, k]
        assert out.shape == like.shape
        return out.clone()


CASES = {
    # name: (seg_len, noise_steps, G, B)
    "avenue_T3": (6, 10, 3, 6),     # shipped Avenue/STC/UBnormal shape (T_c = 3)
    "plumb_N2": (6, 2, 2, 5),       # BASELINE.json config[0]: noise_steps=2 plumbing case
    "stress_T24": (27, 10, 2, 3),   # BASELINE.json [B,2,24,17] shape (seg_len 27, cond [0,1,2])
    "mid_T6": (9, 10, 2, 3),        # the other frame counts the library carries kernels for (MCD_FOR_EACH_T)
    "mid_T12": (15, 10, 2, 3),
}
STRATEGIES = ("best", "worst", "mean", "median", "mean_pose", "median_pose", "quantile:0.25", "all")


def run_case(MoCoDAD, name: str, seg_len: int, N: int, G: int, B: int, out_dir: str) -> None:
    T = seg_len - 3
    args = make_args(seg_len, N, G)
    model = MoCoDAD(args).eval()
    spec = synth.state_dict_spec(T=T, T_cond=3)
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(spec.keys()), "state_dict key order differs from the reference"
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(spec[k]), (k, v.shape, spec[k])
    sd = synth.synth_state_dict(spec, seed=0)
    model.load_state_dict(sd, strict=True)
    batch = synth.synth_batch(B, seg_len, seed=1)
    noise = synth.synth_noise(G, N, B, T, seed=2)
    rec = {}
    with torch.no_grad():
        data = batch[0]
        cond, corrupt, idxs = model._select_frames(data)
        cond_emb, _ = model._encode_condition(cond)
        rec["cond_emb"] = cond_emb.numpy()
        # first denoiser call with per-layer taps
        taps = {}
        hooks = []
        for lname, mod in model.model.named_modules():
            if type(mod).__name__ in ("ST_GCNN_layer", "CNN_layer"):
                hooks.append(mod.register_forward_hook(
                    lambda m, i, o, lname=lname: taps.__setitem__(lname, o.detach().clone())))
        t = torch.full((B,), N - 1, dtype=torch.long)
        eps = model._unet_forward(noise[0, 0], t=t, condition_data=cond_emb, corrupt_idxs=idxs[1])
        for h in hooks:
            h.remove()
        rec["eps_first"] = eps.numpy()
        # CNN_layer taps are in the permuted [B,V',C,T] frame; store them as [B,C,T,V'] like the port
        port_taps = {}
        eps_p = ref_port.unet_forward(sd, noise[0, 0], t, cond_emb, taps=port_taps)
        assert torch.equal(eps_p, eps), f"{name}: port eps != reference eps"
        for lname, val in taps.items():
            if lname in ("down1", "down2", "up2", "up3"):
                val = val.permute(0, 2, 3, 1).contiguous()
            assert torch.equal(val, port_taps[lname]), f"{name}: tap {lname} differs"
            if T <= 3:  # keep the fixtures of the longer windows small (their taps are still asserted equal above)
                rec["tap_" + lname] = val.numpy()
        assert torch.equal(ref_port.cond_encode(sd, cond), cond_emb)

        # full forward with injected noise
        real_randn_like = torch.randn_like
        for strat in STRATEGIES:
            feeder = NoiseFeeder(noise, N)
            torch.randn_like = feeder
            try:
                if strat in ("mean", "median", "quantile:0.25"):
                    out = model.forward(batch, aggr_strategy=strat, return_="loss")
                    loss, sel = out[0], None
                else:
                    out = model.forward(batch, aggr_strategy=strat, return_="all")
                    loss, sel = out[0], out[1]
            finally:
                torch.randn_like = real_randn_like
            assert feeder.calls == G * max(N - 1, 1) - (0 if N > 2 else 0) or N == 2, feeder.calls
            p_loss, p_sel = ref_port.reverse_diffusion(
                sd, data, noise_steps=N, n_generated_samples=G, noise=noise, strategy=strat)
            assert torch.equal(p_loss, loss), f"{name}/{strat}: port loss != reference"
            if sel is not None:
                assert torch.equal(p_sel, sel), f"{name}/{strat}: port selection != reference"
            key = strat.replace(":", "_").replace(".", "p")
            rec["loss_" + key] = loss.numpy()
            if strat == "best":
                rec["x_sel"] = sel.numpy()

        # shared-seed call-order check: the reference draws with torch.randn_like itself
        torch.manual_seed(999)
        ref_loss = model.forward(batch, aggr_strategy="best", return_="loss")[0]
        torch.manual_seed(999)
        p_loss, _ = ref_port.reverse_diffusion(sd, data, noise_steps=N, n_generated_samples=G,
                                               randn_like=torch.randn_like)
        assert torch.equal(p_loss, ref_loss), f"{name}: seeded call order differs"
        rec["loss_seed999_cpu"] = ref_loss.numpy()

        # the schedule the reference module holds (mocodad.py:799-808)
        b, a, ah = ref_port.schedule(N)
        assert torch.equal(b, model._beta_) and torch.equal(a, model._alpha_) and torch.equal(ah, model._alpha_hat_)

    rec["meta"] = np.array([seg_len, N, G, B], dtype=np.int64)
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"[golden] {name}: wrote {path} ({os.path.getsize(path)/1024:.0f} KiB); port bit-identical to reference")


def main() -> None:
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    MoCoDAD = load_reference()
    only = [a for a in sys.argv[1:] if a in CASES]   # e.g. `python oracle/make_golden.py mid_T6 mid_T12`: just these fixtures
    for name, (seg_len, N, G, B) in CASES.items():
        if only and name not in only:
            continue
        run_case(MoCoDAD, name, seg_len, N, G, B, out_dir)
    if only:
        return
    # schedules, straight from the reference's Diffusion class (utils/diffusion_utils.py:38-44)
    from utils.diffusion_utils import Diffusion  # type: ignore
    sched = {}
    for N in (2, 10, 50, 1000):
        d = Diffusion(noise_steps=N, device="cpu")
        sched[f"beta_{N}"] = d.beta.numpy()
        sched[f"alpha_hat_{N}"] = d.alpha_hat.numpy()
        b, a, ah = ref_port.schedule(N)
        assert torch.equal(b, d.beta) and torch.equal(ah, d.alpha_hat)
    np.savez_compressed(os.path.join(out_dir, "schedule.npz"), **sched)
    print("[golden] schedule.npz written")


if __name__ == "__main__":
    main()
