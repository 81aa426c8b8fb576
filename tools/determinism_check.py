"""Run-to-run bit reproducibility of the denoiser (same input, same launch, repeated) and its distance to the oracle.
usage: python tools/determinism_check.py [T] [n] [repeats]      (on the GPU box)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from mocodad_b200 import ScoringEngine, synthetic as synth  # noqa: E402
from oracle import ref_port  # noqa: E402  (test infrastructure: the checker)

T = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
seg_len, N = T + 3, 10
eng = ScoringEngine(seg_len=seg_len, n_frames_cond=3, noise_steps=N, device="cuda:0")
sd = synth.synth_state_dict(synth.state_dict_spec(T=T, T_cond=3), seed=0)
eng.load_state_dict(sd)
batch = synth.synth_batch(n, seg_len, seed=61)
x0 = synth.synth_noise(1, N, n, T, seed=62)[0, 0]
with torch.no_grad():
    cond, _ = ref_port.select_frames(batch[0], (0, 1, 2))
    emb = ref_port.cond_encode(sd, cond)
    taps = {}
    eps = ref_port.unet_forward(sd, x0, torch.full((n,), 6, dtype=torch.long), emb, taps=taps)
x, demb = x0.cuda().contiguous(), emb.cuda()
outs = [eng.unet_forward(x, 6, demb).cpu().numpy() for _ in range(reps)]
for i, o in enumerate(outs):
    d = np.abs(o - eps.numpy())
    print(f"rep {i}: max|gpu-oracle| {d.max():.3e}  n>2e-5 {int((d > 2e-5).sum())}  bit-identical to rep 0: {bool((o == outs[0]).all())}"
          f"  differing elements {int((o != outs[0]).sum())}")
for k in ("st_gcnnsd3.0", "st_gcnnsd3.1", "st_gcnnsu4.0", "st_gcnnsu4.1", "st_gcnnsu3.0"):
    want = taps[k]
    got = [eng.unet_tap(x, 6, demb, k, want.shape[1], want.shape[3]).cpu().numpy() for _ in range(3)]
    print(k, "max|gpu-oracle|", f"{np.abs(got[0] - want.numpy()).max():.3e}", "reps identical:", all(bool((g == got[0]).all()) for g in got),
          "max|want|", f"{np.abs(want.numpy()).max():.2f}")
