"""Which block of the denoiser is not bit-reproducible?  Runs the layer taps of the U-Net's middle (the blocks the tensor-core
kernels serve) `reps` times on the same input and attributes every run-to-run difference to the FIRST tap of the chain that
differs from repeat 0, with the shape of the difference (windows, channels, rows).
usage: python tools/determinism_taps.py [T] [n] [reps]      (on the GPU box; MOCODAD_B200_LIB selects a library variant)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from mocodad_b200 import ScoringEngine, synthetic as synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
seg_len, N = T + 3, 10
CHAIN = (("st_gcnnsd2.1", 64, 12), ("st_gcnnsd3.0", 128, 10), ("st_gcnnsd3.1", 64, 10), ("st_gcnnsu4.0", 64, 12), ("st_gcnnsu4.1", 32, 12),
         ("st_gcnnsu3.0", 32, 17))
eng = ScoringEngine(seg_len=seg_len, n_frames_cond=3, noise_steps=N, device="cuda:0")
eng.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=T, T_cond=3), seed=0))
batch = synth.synth_batch(n, seg_len, seed=61)
x = synth.synth_noise(1, N, n, T, seed=62)[0, 0].cuda().contiguous()
emb = eng.cond_encode(batch[0].cuda())


def taps():
    return [eng.unet_tap(x, 6, emb, k, c, v).clone() for k, c, v in CHAIN]


ref = taps()
torch.cuda.synchronize()
bad = {}
for r in range(1, reps):
    cur = taps()
    for (k, c, v), a, b in zip(CHAIN, ref, cur):
        if not torch.equal(a, b):
            d = (a != b).cpu().numpy()
            w, ch, t_, v_ = np.nonzero(d)
            rows = t_ * v + v_
            mx = float((a - b).abs().max())
            bad.setdefault(k, []).append(r)
            print(f"rep {r}: first differing tap {k}: {d.sum()} elements, max|diff| {mx:.3e}, windows {sorted(set(w.tolist()))[:8]} "
                  f"(tile mod 148: {sorted(set((w % 148).tolist()))[:8]}, tile round {sorted(set((w // 148).tolist()))[:8]}), "
                  f"channels {min(ch)}..{max(ch)} ({len(set(ch.tolist()))} distinct; chunks {sorted(set((ch // 16).tolist()))}), "
                  f"rows {min(rows)}..{max(rows)} ({len(set(rows.tolist()))} distinct)")
            break
print(f"T={T} n={n} reps={reps} lib={os.environ.get('MOCODAD_B200_LIB', 'default')}: "
      + (", ".join(f"{k}: {len(v)} differing repeats" for k, v in bad.items()) if bad else "all repeats bit-identical"))
