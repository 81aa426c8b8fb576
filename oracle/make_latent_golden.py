"""Pins oracle/latent_port.py against the UNMODIFIED reference (models/mocodad_latent.py, stage 'diffusion') and writes
tests/golden/latent_T3.npz.  Run in the build container (needs /root/reference): python oracle/make_latent_golden.py

Groundwork for SURVEY.md 8 row f4: there is no CUDA path for the latent variant yet; this fixes the oracle it will be
checked against.  The reference module is built from config/UBnormal/mocodad-latent_test.yaml (seg_len 6, latent 64,
hidden_sizes [64,128,128,64]); its weights are a seeded synthetic checkpoint (names / shapes taken from the module itself and
stored in the fixture), BatchNorm statistics randomised; torch.randn / torch.randn_like are replaced by pre-drawn tensors
inside forward (mocodad_latent.py:104,117)."""
import argparse
import os
import sys
import tempfile
from collections import OrderedDict

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402  (stubs for pytorch_lightning / matplotlib, reference path)
from oracle import latent_port, ref_port, synth  # noqa: E402

N, G, B, SEG_LEN, LATENT = 10, 3, 6, 6, 64


def main() -> None:
    mg.load_reference()
    from models.mocodad_latent import MoCoDADlatent  # type: ignore  (the reference)
    cfg = yaml.load(open(os.path.join(mg.REF, "config/UBnormal/mocodad-latent_test.yaml")), Loader=yaml.FullLoader)
    with tempfile.TemporaryDirectory() as tmp:
        ckpt = os.path.join(tmp, "pretrain.ckpt")
        torch.save({"state_dict": {}}, ckpt)   # _freeze_main_net_and_load_ckpt loads it with strict=False; weights come below
        cfg.update(seg_len=SEG_LEN, noise_steps=N, n_generated_samples=G, gt_path="", ckpt_dir="", save_tensors=False,
                   pretrained_model_ckpt_path=ckpt, latent_embedding_dim=LATENT)
        model = MoCoDADlatent(argparse.Namespace(**cfg)).eval()
    hidden = list(cfg["hidden_sizes"])
    spec = OrderedDict((k, tuple(v.shape)) for k, v in model.state_dict().items())
    sd = synth.synth_state_dict(spec, seed=0)
    model.load_state_dict(sd, strict=True)
    T = SEG_LEN - 3
    batch = synth.synth_batch(B, SEG_LEN, seed=1)
    g = torch.Generator().manual_seed(12)
    noise = torch.randn(G, N - 1, B, LATENT, generator=g)

    calls = {"n": 0}

    def feed(*a, **k):
        gi, ki = divmod(calls["n"], N - 1)
        calls["n"] += 1
        return noise[gi, ki].clone()
    rec = {}
    real_randn, real_randn_like = torch.randn, torch.randn_like
    for strat in ("best", "mean", "median"):
        calls["n"] = 0
        torch.randn, torch.randn_like = feed, feed
        try:
            with torch.no_grad():
                out = model.forward(batch, aggr_strategy=strat, return_="all" if strat == "best" else "loss")
        finally:
            torch.randn, torch.randn_like = real_randn, real_randn_like
        assert calls["n"] == G * (N - 1), calls
        with torch.no_grad():
            loss, sel, code = latent_port.latent_reverse_diffusion(sd, batch[0], noise_steps=N, n_generated_samples=G,
                                                                   n_layers=len(hidden), noise=noise, strategy=strat)
        assert torch.equal(loss, out[0]), f"{strat}: port loss != reference"
        rec["loss_" + strat] = out[0].numpy()
        if strat == "best":
            assert torch.equal(sel, out[1]), "port selection != reference"
            rec["latent_sel"] = out[1].numpy()
    with torch.no_grad():
        cond, corrupt, idxs = model._select_frames(batch[0])
        cond_emb, _ = model._encode_condition(cond)
        t = torch.full((B,), -1, dtype=torch.long)
        code_ref = model._unet_forward(corrupt, t=t, condition_data=cond_emb, corrupt_idxs=idxs[1])
        taps = {}
        code_port = latent_port.latent_encode(sd, corrupt, ref_port.cond_encode(sd, cond), taps=taps)
        assert torch.equal(code_ref, code_port), "latent code differs"
        tt = torch.full((B,), 7, dtype=torch.long)
        eps_ref = model.denoiser(noise[0, 0], tt, cond_emb)
        eps_port = latent_port.denoiser_forward(sd, noise[0, 0], tt, cond_emb, len(hidden))
        assert torch.equal(eps_ref, eps_port), "denoiser output differs"
    rec.update(latent_code=code_ref.numpy(), eps_t7=eps_ref.numpy(), noise=noise.numpy(), tap_sd3_1=taps["st_gcnnsd3.1"].numpy(),
               meta=np.array([SEG_LEN, N, G, B, LATENT] + hidden, dtype=np.int64),
               spec_names=np.array(list(spec.keys())), spec_shapes=np.array([",".join(map(str, s)) for s in spec.values()]))
    path = os.path.join(ROOT, "tests", "golden", "latent_T3.npz")
    np.savez_compressed(path, **rec)
    print(f"[golden] latent_T3: port bit-identical to MoCoDADlatent (stage 'diffusion', {len(spec)} state_dict entries); wrote {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
