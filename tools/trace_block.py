"""Device-side timeline of one tensor-core block launch (debug aid): prints per-role event times for CTA 0.
Needs the instrumented build: `python -m mocodad_b200._build --trace` (here, before gpurun; the .so travels).
usage (on the GPU box): python tools/trace_block.py [slot=2] [n_windows=2368]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MOCODAD_B200_LIB"] = os.path.join(ROOT, "mocodad_b200", "libmocodad_b200_trace.so")
import torch
from mocodad_b200 import ScoringEngine, synthetic as synth
from mocodad_b200._lib import check

slot = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
eng = ScoringEngine(seg_len=27, n_frames_cond=3, noise_steps=10, device="cuda:0")
eng.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=24, T_cond=3), seed=0))
x = torch.randn(n, 2, 24, 17, device="cuda")
cond = torch.randn(n, 16, device="cuda")
eng.unet_forward(x, 5, cond)  # warm
cap = 4096
rec = torch.zeros(cap, 4, dtype=torch.int64, device="cuda")
check(eng.lib.mcd_debug_trace_next(eng._h, slot, rec.data_ptr(), cap))
eng.unet_forward(x, 5, cond)
torch.cuda.synchronize()
r = rec.cpu()
r = r[r[:, 3] > 0]
t0 = int(r[:, 3].min())
roles = {0: "T", 1: "A", 2: "MMA", 3: "LOAD", 4: "EPI"}
evn = {0: {0: "top", 1: "x_full", 2: "y1_empty", 3: "done"}, 1: {0: "top", 1: "y1_full", 2: "y2_free", 3: "mix done", 4: "ops ready"},
       2: {0: "top", 1: "ops_full", 2: "w_full", 3: "issued"}, 3: {0: "top", 1: "x_empty", 2: "x issued"},
       4: {0: "top", 1: "acc_full", 2: "done"}}
rows = sorted(r.tolist(), key=lambda q: q[3])
print(f"{len(rows)} records; showing pairs 4..9")
for role, it, ev, clk in rows:
    if 4 <= it <= 9 or (role == 4 and 2 <= it <= 4):
        print(f"{clk - t0:9d}  {roles[role]:5s} it={it:3d}  {evn[role].get(ev, ev)}")
