"""Runs the window-ingest kernels on synthetic frame rows at bench size, for ncu:
  ncu --set full --clock-control none --import-source on -k regex:'normalize_frames|build_items' -c 4 -f -o gpurun_out/ingest \
      python tools/ingest_profile.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocodad_b200 import ScoringEngine, synthetic as synth  # noqa: E402
from mocodad_b200.engine import pose_transform_matrices  # noqa: E402

dev = torch.device("cuda:0")
eng = ScoringEngine(seg_len=27, n_frames_cond=3, noise_steps=10, device=dev)
eng.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=24, T_cond=3), seed=0))
n_traj, traj_len = 1024, 1024
gen = torch.Generator(device=dev).manual_seed(4242)
rows = (torch.rand(n_traj * traj_len, 34, device=dev, generator=gen) * 320.0 + 8.0).contiguous()
rows[torch.rand(rows.shape, device=dev, generator=gen) < 0.08] = 0.0
win_start = (torch.arange(n_traj, device=dev)[:, None] * traj_len + torch.arange(traj_len - 26, device=dev)[None, :]).reshape(-1).contiguous()
mats = pose_transform_matrices(5)
center, scale = np.zeros(34), np.full(34, 0.25)
for _ in range(2):
    norm = eng.normalize_frames(rows, (640.0, 360.0), center=center, scale=scale)
    items = eng.build_items(norm, win_start, mats=mats, first_item=win_start.numel() - 131072, n_items=262144)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
if "--time" in sys.argv:   # plain timing (never under ncu): CUDA events on the launching stream, 20 launches each
    ev[0].record()
    for _ in range(20):
        eng.normalize_frames(rows, (640.0, 360.0), out=norm, center=center, scale=scale)
    ev[1].record()
    for _ in range(20):
        items = eng.build_items(norm, win_start, mats=mats, first_item=win_start.numel() - 131072, n_items=262144)
    ev[2].record()
    torch.cuda.synchronize()
    a, b = ev[0].elapsed_time(ev[1]) / 20 * 1e-3, ev[1].elapsed_time(ev[2]) / 20 * 1e-3
    print(f"normalize_frames {a * 1e6:.1f} us/launch = {2 * rows.numel() * 4 / a / 1e9:.0f} GB/s; "
          f"build_items {b * 1e6:.1f} us/launch = {(items.numel() + rows.numel() * 0) * 4 / b / 1e9:.0f} GB/s (output bytes only)")
print("rows", rows.shape[0], "items", items.shape[0], "finite", bool(torch.isfinite(items).all()))
