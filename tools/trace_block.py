"""Device-side timeline of one tensor-core block launch (debug aid): prints per-role event times for CTA 0.
Needs the instrumented build: `python -m mocodad_b200._build --trace` (here, before gpurun; the .so travels).
usage (on the GPU box): python tools/trace_block.py [slot=2] [n_windows=2368] [events|waits] [seg_len=27]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MOCODAD_B200_LIB", os.path.join(ROOT, "mocodad_b200", "libmocodad_b200_trace.so"))
import torch
from mocodad_b200 import ScoringEngine, synthetic as synth
from mocodad_b200._lib import check

slot = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
events = not (len(sys.argv) > 3 and sys.argv[3] == "waits")  # "waits": wait accounting only (no per-event perturbation)
seg_len = int(sys.argv[4]) if len(sys.argv) > 4 else 27
T = seg_len - 3
eng = ScoringEngine(seg_len=seg_len, n_frames_cond=3, noise_steps=10, device="cuda:0")
eng.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=T, T_cond=3), seed=0))
x = torch.randn(n, 2, T, 17, device="cuda")
cond = torch.randn(n, 16, device="cuda")
eng.unet_forward(x, 5, cond)  # warm
cap = 4096
rec = torch.zeros(cap, 4, dtype=torch.int64, device="cuda")
check(eng.lib.mcd_debug_trace_next(eng._h, slot, rec.data_ptr(), cap if events else -cap))
eng.unet_forward(x, 5, cond)
torch.cuda.synchronize()
r = rec.cpu()
roles = {0: "T", 1: "A", 2: "MMA", 3: "LOAD", 4: "EPI", 5: "WLOAD", 6: "XLO"}
waits = {0: ["x_full", "y1_empty"], 1: ["y1_full", "y2_free"], 2: ["w_full", "acc_empty", "xlo_full", "ops_full"], 3: ["x_empty"],
         4: ["acc_full", "emb phase", "tmem ld+wait", "math+store"], 5: ["w_free"], 6: ["x_full", "xlo_free"]}
acc = r[(r[:, 1] >= 99) & (r[:, 1] < 104)]
print(f"slot {slot} ({eng.lib.mcd_profile_slot_name(slot).decode()}), {n} windows: cycles CTA 0 spent waiting, by role")
for role in sorted(roles):
    rows = acc[acc[:, 0] == role]
    if len(rows) == 0:
        continue
    total = int(rows[rows[:, 1] == 99][0, 3])
    parts = []
    for k, name in enumerate(waits[role]):
        w = int(rows[rows[:, 1] == 100 + k][0, 3])
        parts.append(f"{name} {w} ({100.0 * w / max(total, 1):.1f}%)")
    print(f"  {roles[role]:5s} total {total:9d}  " + "  ".join(parts))
if events:
    r = r[(r[:, 3] > 0) & (r[:, 1] < 99)]
    t0 = int(r[:, 3].min())
    evn = {0: {0: "top", 1: "x_full", 2: "y1_empty", 3: "done"}, 1: {0: "top", 1: "y1_full", 2: "y2_free", 3: "mix done", 4: "ops ready"},
           2: {0: "top", 1: "res issued", 2: "ops_full", 3: "issued"}, 3: {0: "top", 1: "x_empty", 2: "x issued"},
           4: {0: "top", 1: "acc_full", 2: "done"}}
    rows = sorted(r.tolist(), key=lambda q: q[3])
    print(f"{len(rows)} records; showing pairs 4..9")
    for role, it, ev, clk in rows:
        if 4 <= it <= 9 or (role == 4 and 2 <= it <= 4):
            print(f"{clk - t0:9d}  {roles[role]:5s} it={it:3d}  {evn[role].get(ev, ev)}")
