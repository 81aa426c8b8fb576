"""The workload compute-sanitizer is pointed at (tools/gpu_sanitize.sh): every kernel of the scoring path with several
tiles per persistent CTA, small enough for the instrumented run.  usage: python tools/sanitize_case.py [T24|T3|latent|all]

  T24     one denoiser call + 2-step reverse diffusion on 148*3 = 444 windows of [2,24,17] (3 tiles per CTA of the
          warp-specialised tensor-core block kernel: ring slots, mbarrier phases and the TMEM set ping-pong all wrap)
  T3      reverse diffusion on 3 700 windows of [2,3,17] (8 windows per tile -> 3.1 tiles per CTA), N=3, G=1
  latent  the latent variant's encode + MLP-denoiser loop on 148*32*2 + 5 vectors

Prints one line per case with a checksum; compare it with an uninstrumented run to confirm the tool saw the real path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mocodad_b200 import ScoringEngine, synthetic as synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"


def engine(seg_len, N, **kw):
    T = seg_len - 3
    eng = ScoringEngine(seg_len=seg_len, n_frames_cond=3, noise_steps=N, device="cuda:0", **kw)
    eng.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=T, T_cond=3, latent_embedding_dim=kw.get("latent_embedding_dim", 0),
                                                                      hidden_sizes=kw.get("hidden_sizes", ())), seed=0))
    return eng


if which in ("T24", "all"):
    eng = engine(27, 3)
    n = 444
    data = synth.synth_batch(n, 27, seed=5)[0].cuda()
    emb = eng.cond_encode(data)
    eps = eng.unet_forward(synth.synth_noise(1, 3, n, 24, seed=6)[0, 0].cuda().contiguous(), 2, emb)
    best = eng.reverse_diffusion(data, 1, seed=7)["best"]
    torch.cuda.synchronize()
    print(f"T24: {n} windows, eps checksum {float(eps.double().sum()):.6f}, best checksum {float(best.double().sum()):.6f}")
if which in ("T3", "all"):
    eng = engine(6, 3)
    n = 3700
    data = synth.synth_batch(n, 6, seed=8)[0].cuda()
    best = eng.reverse_diffusion(data, 1, seed=9)["best"]
    torch.cuda.synchronize()
    print(f"T3: {n} windows, best checksum {float(best.double().sum()):.6f}")
if which in ("latent", "all"):
    eng = engine(6, 4, latent_embedding_dim=64, hidden_sizes=[64, 128, 128, 64])
    B, G = 3159, 3
    data = synth.synth_batch(B, 6, seed=10)[0].cuda()
    best = eng.latent_reverse_diffusion(data, G, seed=11)["best"]
    torch.cuda.synchronize()
    print(f"latent: {B * G} vectors, best checksum {float(best.double().sum()):.6f}")
