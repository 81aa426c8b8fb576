#!/usr/bin/env python
"""bench.py -- throughput of the reverse-diffusion scoring path (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this implementation (CUDA)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on host cores

One "step" = one pass of the hot path (the body of MoCoDAD.forward, models/mocodad.py:129-184) over
one batch of synthetic windows: B windows/GPU of [2, seg_len=27, 17] (3 conditioning + T=24 denoised
frames, BASELINE.json's [B,2,24,17] shape), noise_steps=10 (9 denoiser calls), n_generated_samples=50,
'best' aggregation, SmoothL1 -- i.e. 450 window-steps per window.  Weights: random "trained-looking"
checkpoint (no checkpoints ship with the reference).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "skeleton_windows_per_sec_full_reverse_diffusion"
UNIT = "windows/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--batch", type=int, default=1024, help="windows per GPU per step")
    p.add_argument("--seg-len", type=int, default=27)
    p.add_argument("--noise-steps", type=int, default=10)
    p.add_argument("--gen", type=int, default=50, help="n_generated_samples")
    p.add_argument("--cpu-sample", default="128x5", help="cpu_baseline sample: windows x samples")
    p.add_argument("--eager-samples", type=int, default=5, help="generated samples timed by the PyTorch CUDA-eager baseline arm")
    p.add_argument("--ref-sample", default="32x5", help="--impl reference: windows x samples per step")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-shipped", action="store_true", help="skip the shipped-shape (T=3, B=1024 / 2048) sub-record")
    p.add_argument("--stress", action="store_true",
                   help="BASELINE.json configs[4]-style sweep (e.g. --batch 125000 --noise-steps 1000 --gen 1 on 8 GPUs): device-timed "
                        "value only, warm-up on 1/32 of the batch, no e2e / baselines / per-kernel profile.  One step = one full sweep "
                        "(~66 s at that size): pass --steps 1 unless an average over several sweeps is wanted")
    return p.parse_args()


def workload_config(a):
    T = a.seg_len - 3
    return {"workload": f"synthetic windows [B,2,{T},17] (seg_len {a.seg_len}, 3 conditioning frames, 'inject'), "
                        f"noise_steps={a.noise_steps}, n_generated_samples={a.gen}, best-of-n SmoothL1 "
                        "(BASELINE.json configs[1] shape)",
            "windows_per_gpu_per_step": a.batch, "seg_len": a.seg_len, "T": T, "V": 17,
            "noise_steps": a.noise_steps, "n_generated_samples": a.gen,
            "window_steps_per_window": a.gen * (a.noise_steps - 1),
            "cache": "per-step working set (activations of one pass of the virtual batch, GBs) exceeds the 126 MB L2; "
                     "fresh Philox noise every step"}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_reference_rate(a, n_windows: int, n_samples: int, steps: int, warmup: int):
    """Times oracle/ref_port.py (the reference's PyTorch operators, CPU) on a bounded sample of the
    workload; returns (equivalent windows/s at the workload's n_generated_samples, seconds/step)."""
    from mocodad_b200 import synthetic as synth
    from oracle import ref_port
    T = a.seg_len - 3
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.synth_state_dict(synth.state_dict_spec(T=T, T_cond=3), seed=0)
    data = synth.synth_batch(n_windows, a.seg_len, seed=1)[0]
    gen = torch.Generator().manual_seed(999)

    def one():
        with torch.no_grad():
            ref_port.reverse_diffusion(sd, data, noise_steps=a.noise_steps, n_generated_samples=n_samples,
                                       randn_like=lambda x: torch.randn(x.shape, generator=gen))
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    window_steps_per_s = n_windows * n_samples * (a.noise_steps - 1) / dt
    return window_steps_per_s / (a.gen * (a.noise_steps - 1)), dt


def cuda_eager_reference_rate(a, dev, n_samples: int, steps: int = 2, warmup: int = 1):
    """The baseline leg's second arm: oracle/ref_port.py -- the reference's own ATen operators (einsum / conv2d / batch_norm /
    prelu / linear, with its permutes and copies) in the reference's order -- run as PyTorch CUDA eager on this GPU, at the
    workload's batch size, on ``n_samples`` of the G generated samples (the samples are sequential in the reference, so the time
    scales linearly).  This is the denominator of north_star's ">= 10x the reference PyTorch-CUDA path"; /root/reference does
    not exist on the GPU box, and the port omits the reference's per-step host->device schedule uploads, so the
    figure favours the reference.  Returns (equivalent windows/s at G samples, seconds per step, launches unknown)."""
    from mocodad_b200 import synthetic as synth
    from oracle import ref_port
    T = a.seg_len - 3
    sd = {k: v.to(dev) for k, v in synth.synth_state_dict(synth.state_dict_spec(T=T, T_cond=3), seed=0).items()}
    data = synth.synth_batch(a.batch, a.seg_len, seed=1)[0].to(dev)

    def one():
        with torch.no_grad(), torch.device(dev):
            loss, _ = ref_port.reverse_diffusion(sd, data, noise_steps=a.noise_steps, n_generated_samples=n_samples,
                                                 randn_like=torch.randn_like)
        return loss
    for _ in range(warmup):
        one()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        loss = one()
    e1.record()
    torch.cuda.synchronize(dev)
    wall = (time.perf_counter() - t0) / steps
    assert loss.is_cuda and bool(torch.isfinite(loss).all())
    dt = max(e0.elapsed_time(e1) * 1e-3 / steps, wall)
    return a.batch * n_samples / dt / a.gen, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nw, ns = (int(x) for x in a.ref_sample.split("x"))
    value, dt = cpu_reference_rate(a, nw, ns, a.steps, a.warmup)
    cores = os.cpu_count() or 1
    sample = (f"{nw} windows x {ns} samples x {a.noise_steps - 1} denoiser calls per step on the host CPU; rate scaled "
              f"to the workload's {a.gen} samples/window (window-steps are identical work)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- CUDA arm
def _stdout_to_stderr():
    """NCCL prints its INFO lines (the communicator / rank lines the driver checks) on the process's stdout.  Route fd 1 to
    stderr so they stay visible, and hand back a private copy of the original stdout for the one JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _emit(fd, line: dict):
    text = json.dumps(line) + "\n"
    if fd is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(fd, text.encode())


def run_b200(a):
    import torch.distributed as dist
    from mocodad_b200 import ScoringEngine
    from mocodad_b200.engine import probe_fp32_detail
    from mocodad_b200 import synthetic as synth
    from mocodad_b200.sharding import gather_scores

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the scoring path has no CPU implementation (use --impl reference "
                         "for the host-core baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    out_fd = None
    if world > 1:
        # NCCL's communicator lines must reach the driver (it checks the rank count); they go to stderr, the JSON line alone to stdout
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE") and not os.environ.get("MCD_QUIET_NCCL"):
            os.environ["NCCL_DEBUG"] = "INFO"          # boxes export VERSION by default: the communicator lines need INFO
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        out_fd = _stdout_to_stderr()
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    G, N = a.gen, a.noise_steps

    def measure(seg_len: int, B: int, steps: int, warmup: int, sample_clocks: bool):
        """value (inputs resident in HBM, device-timed) and e2e (host buffers through mcd_score_windows_host) for one shape."""
        T = seg_len - 3
        eng = ScoringEngine(seg_len=seg_len, n_frames_cond=3, noise_steps=N, device=dev)
        eng.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=T, T_cond=3), seed=0))
        host = synth.synth_batch(B, seg_len, seed=1 + rank)[0].pin_memory()
        data = host.to(dev)
        first = rank * B  # this rank's windows in the global index space (weak scaling: B windows per rank)
        step_no = [0]

        def step_device():
            res = eng.reverse_diffusion(data, G, seed=999, first_window=first + step_no[0] * world * B)
            step_no[0] += 1
            # N > 1: the path's one exchange -- all-gather of the per-window scores
            return gather_scores(res["best"], world * B) if world > 1 else res["best"]

        def step_host():
            out = eng.score_windows_host(host, G, seed=999, first_window=first + step_no[0] * world * B)
            step_no[0] += 1
            if world > 1:   # the same exchange, end to end: scores back onto the device, one all-gather, all scores to the host
                out = gather_scores(out.to(dev, non_blocking=True), world * B).cpu()
            return out

        for _ in range(warmup):
            step_device()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = eng.launch_count()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.__enter__()
        barrier()
        ev0.record()
        for _ in range(steps):
            best = step_device()
        ev1.record()
        barrier()
        if sampler:
            sampler.__exit__()
        launches = eng.launch_count() - launches0
        ms = max_over_ranks(ev0.elapsed_time(ev1)) / steps
        assert bool(torch.isfinite(best).all())
        for _ in range(min(warmup, 2)):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            scores = step_host()
        torch.cuda.synchronize(dev)
        e2e_s = max_over_ranks(time.perf_counter() - t0) / steps
        barrier()
        assert bool(torch.isfinite(scores).all())
        return {"eng": eng, "host": host, "data": data, "ms": ms, "value": world * B / (ms * 1e-3), "e2e_s": e2e_s,
                "e2e_value": world * B / e2e_s, "launches": launches, "clocks": sampler.summary() if sampler else None,
                "step_device": step_device}

    T = a.seg_len - 3
    B = a.batch
    if a.stress:
        eng = ScoringEngine(seg_len=a.seg_len, n_frames_cond=3, noise_steps=N, device=dev)
        eng.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=T, T_cond=3), seed=0))
        data = synth.synth_batch(B, a.seg_len, seed=1 + rank)[0].to(dev)
        small = data[: max(1, B // 32)].contiguous()
        for _ in range(max(a.warmup, 1)):
            eng.reverse_diffusion(small, G, seed=999, first_window=rank * B)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        with ClockSampler(local) as clocks:
            barrier()
            ev0.record()
            for k in range(a.steps):
                best = eng.reverse_diffusion(data, G, seed=999, first_window=(k * world + rank) * B)["best"]
                if world > 1:
                    best = gather_scores(best, world * B)
            ev1.record()
            barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1)) / a.steps
        assert bool(torch.isfinite(best).all())
        if rank == 0:
            value = world * B / (ms * 1e-3)
            bytes_ws = {3: 230952.0, 24: 1847616.0}.get(T)
            peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
            hbm_peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
            line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 1),
                    "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": dict(workload_config(a), stress=True, total_windows=world * B,
                                   warmup_note="warm-up steps score 1/32 of the batch (same kernels)"),
                    "window_steps_per_sec": value * G * (N - 1), "gpu_launches": eng.launch_count() - l0, "clocks": clocks.summary(),
                    "whole_step_hbm": None if bytes_ws is None else {
                        "achieved_per_gpu": round(bytes_ws * value * G * (N - 1) / world / 1e9, 1), "peak": hbm_peak, "unit": "GB/s",
                        "frac": round(bytes_ws * value * G * (N - 1) / world / 1e9 / hbm_peak, 4)}}
            _emit(out_fd, line)
        if world > 1:
            dist.destroy_process_group()
        return
    main = measure(a.seg_len, B, a.steps, a.warmup, True)
    eng, host, data, ms, value, e2e_s, e2e_value = (main[k] for k in ("eng", "host", "data", "ms", "value", "e2e_s", "e2e_value"))
    launches, clocks_summary, step_device = main["launches"], main["clocks"], main["step_device"]

    # ---- row f1: dataset items (5 affine transforms per base window) built on the device ----------
    from mocodad_b200.engine import pose_transform_matrices
    mats = pose_transform_matrices(5)
    n_base = (B + 4) // 5
    base = data[:n_base].contiguous()
    for _ in range(3):
        eng.expand_transforms(base, mats, 0, min(B, 5 * n_base))
    xe0, xe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xe0.record()
    for _ in range(20):
        items = eng.expand_transforms(base, mats, 0, min(B, 5 * n_base))
    xe1.record()
    torch.cuda.synchronize(dev)
    xf_s = xe0.elapsed_time(xe1) * 1e-3 / 20
    ingest = {"kernel": "expand_transforms", "items_per_s": items.shape[0] / xf_s,
              "GBps": round(2 * items.numel() * 4 / xf_s / 1e9, 1),
              "note": "dataset item = affine transform idx//N of base window idx%N (utils/dataset.py:67-76), built on the "
                      "device from base windows uploaded once; launch-latency-bound at this batch size"}

    # ---- row f1, second slice: trajectory frame rows -> dataset items (normalise, window, scale, transform) in HBM -------
    n_traj, traj_len = 1024, 1024
    gen = torch.Generator(device=dev).manual_seed(4242)
    rows = (torch.rand(n_traj * traj_len, 34, device=dev, generator=gen) * 320.0 + 8.0).contiguous()
    rows[torch.rand(rows.shape, device=dev, generator=gen) < 0.08] = 0.0
    span = eng.seg_len
    win_start = (torch.arange(n_traj, device=dev)[:, None] * traj_len + torch.arange(traj_len - span + 1, device=dev)[None, :]).reshape(-1).contiguous()
    n_win = win_start.numel()
    n_it = min(5 * n_win, 262144)
    center, scale = np.zeros(34), np.full(34, 0.25)
    norm = eng.normalize_frames(rows, (640.0, 360.0), center=center, scale=scale)
    for _ in range(3):
        eng.normalize_frames(rows, (640.0, 360.0), out=norm, center=center, scale=scale)
        tr_items = eng.build_items(norm, win_start, mats=mats, first_item=n_win - n_it // 2, n_items=n_it)
    te = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    te[0].record()
    for _ in range(10):
        eng.normalize_frames(rows, (640.0, 360.0), out=norm, center=center, scale=scale)
    te[1].record()
    for _ in range(10):
        tr_items = eng.build_items(norm, win_start, mats=mats, first_item=n_win - n_it // 2, n_items=n_it)
    te[2].record()
    torch.cuda.synchronize(dev)
    nf_s, bi_s = te[0].elapsed_time(te[1]) * 1e-4, te[1].elapsed_time(te[2]) * 1e-4
    assert bool(torch.isfinite(tr_items).all())
    ingest["trajectories"] = {
        "frame_rows": rows.shape[0], "windows": n_win, "items_built": n_it,
        "normalize_frames": {"rows_per_s": rows.shape[0] / nf_s, "GBps": round(2 * rows.numel() * 4 / nf_s / 1e9, 1)},
        "build_items": {"items_per_s": n_it / bi_s, "GBps": round((tr_items.numel() + rows.numel()) * 4 / bi_s / 1e9, 1),
                        "includes": "output tensor allocation by the caller (torch caching allocator)"},
        "note": "reference on-disk trajectories (frame, 17 x,y) -> bounding-box-centre coordinates (utils/data.py:165-187) -> "
                "RobustScaler (utils/data.py:345-354, fused into the same pass) -> sliding windows (utils/preprocessing.py:55-86) -> 5 transforms "
                "(utils/dataset.py:67-76); frame rows cross PCIe once, no window tensor is materialised on the host"}
    del rows, norm, tr_items, win_start

    # ---- per-kernel device times (CUDA events around every launch, one extra step) -----------
    eng.profile_enable(True)
    step_device()
    prof = eng.profile_read()
    eng.profile_enable(False)
    total_ms = sum(v["ms"] for v in prof.values())
    kernels = []
    for name, v in prof.items():
        if v["launches"] == 0:
            continue
        sec = v["ms"] * 1e-3
        kernels.append({"kernel": name, "launches": v["launches"], "ms": round(v["ms"], 3),
                        "share": round(v["ms"] / total_ms, 4),
                        "GBps": round(v["bytes_per_window"] * v["windows"] / sec / 1e9, 1),
                        "TFLOPs": round(v["flops_per_window"] * v["windows"] / sec / 1e12, 2)})
    kernels.sort(key=lambda k: -k["ms"])
    top = kernels[0]
    pv = prof[top["kernel"]]

    # ---- the shipped shape (every config under config/ has seg_len 6 -> T = 3; batch 1024 Avenue / UBnormal, 2048 STC) ----
    shipped = None
    if not a.no_shipped and a.seg_len != 6:
        shipped = []
        for sb in (1024, 2048):
            r = measure(6, sb, max(3, a.steps), 3, False)
            algo = 230952.0 * G * (N - 1)   # algorithmic bytes per window (per-block-fused contract, T=3)
            shipped.append({"r": r, "B": sb, "algo": algo})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    if "hbm_gbs" in peaks:
        hbm_peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ffma_peak, ffma2_peak = probe_fp32_detail(local)
    fp32_peak = max(ffma_peak, ffma2_peak)
    top_sec_per_launch = pv["ms"] * 1e-3 / pv["launches"]
    windows_per_launch = pv["windows"] / pv["launches"]
    top_bytes_per_launch = pv["bytes_per_window"] * windows_per_launch
    top_flops_per_launch = pv["flops_per_window"] * windows_per_launch
    achieved = top_bytes_per_launch / top_sec_per_launch / 1e9
    per_call = [k for k in prof if k.startswith("st_gcnn") or k in ("down1", "down2", "up3", "up2", "ddpm_step")]
    unet_flops = sum(prof[k]["flops_per_window"] for k in per_call)  # one denoiser call + DDPM update, per window
    unet_bytes = {3: 230952.0, 24: 1847616.0}.get(T)                 # algorithmic bytes per window-step
    # measured DRAM traffic of that kernel: ncu --set full capture (profiles/ncu_traffic.json; per window at the launch size named there)
    traffic, traffic_src = None, None
    for tname in ("ncu_traffic.json", "r01_ncu_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            per_window = tj["dram_bytes_per_window"].get(top["kernel"]) if tj.get("T", 24) == T else None
            if per_window is not None:
                traffic = per_window * windows_per_launch
                traffic_src = f"profiles/{tname} ({tj.get('windows_per_launch', 2368)} windows per captured launch)"
                break
    # tensor-pipe view of the same kernel: 3xTF32 executes 3 tf32 MMAs per fp32 product of the 1x1 convolution(s)
    tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
    conv = {"st_gcnnsd1.0": (16, 32, 17), "st_gcnnsd1.1": (32, 32, 17), "st_gcnnsd2.0": (32, 64, 12), "st_gcnnsd2.1": (64, 64, 12),
            "st_gcnnsd3.0": (64, 128, 10), "st_gcnnsd3.1": (128, 64, 10), "st_gcnnsu4.0": (64, 64, 12), "st_gcnnsu4.1": (64, 32, 12),
            "st_gcnnsu3.0": (32, 32, 17)}
    tensor = None
    if top["kernel"] in conv:
        ci, co, v = conv[top["kernel"]]
        mma_flops = 3 * 2.0 * ci * co * T * v * (2 if ci != co else 1) * windows_per_launch
        tensor = {"achieved_tflops": round(mma_flops / top_sec_per_launch / 1e12, 1), "peak_tflops": round(tf32_peak, 1),
                  "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (kind::tf32 runs at half the bf16 rate)",
                  "frac": round(mma_flops / top_sec_per_launch / 1e12 / tf32_peak, 4),
                  "note": "tcgen05.mma kind::tf32, three MMAs per fp32 product (hi*hi + hi*lo + lo*hi)"}
    # FMA-pipe view: only the position mixes run there in the tensor-core blocks (the 1x1 convolutions are on the tensor pipe)
    fma_flops_per_launch = top_flops_per_launch
    fma_counts = "the T-mix and A-mix FMAs of the kernel (2*Cin*P*(T+V) per window); its convolution flops are under 'tensor'"
    conv_first = {"st_gcnnsd3.1": 12, "st_gcnnsu4.1": 17}   # conv-first blocks: joints of the fused up-path CNN_layer's output
    if top["kernel"] in conv:
        ci, co, v = conv[top["kernel"]]
        fma_flops_per_launch = 2.0 * ci * T * v * (T + v) * windows_per_launch
        if top["kernel"] in conv_first:   # mixes run on the Cout side there, plus the fused joint resample (V -> Vup) on the FMA pipe
            vup = conv_first[top["kernel"]]
            fma_flops_per_launch = (2.0 * co * T * v * (T + v) + 2.0 * co * T * v * vup) * windows_per_launch
            fma_counts = ("conv-first block: T-mix and A-mix on the Cout side (2*Cout*P*(T+V) per window) + the fused joint resample "
                          "(2*Cout*T*V*Vup); its convolution flops are under 'tensor'")
    fp32 = {"achieved_tflops": round(fma_flops_per_launch / top_sec_per_launch / 1e12, 2),
            "peak_tflops": round(fp32_peak, 2), "peak_source": "mcd_probe_fp32_detail: max of FFMA / FFMA2 register loops on this GPU",
            "probe_ffma_tflops": round(ffma_peak, 2), "probe_ffma2_tflops": round(ffma2_peak, 2),
            "frac": round(fma_flops_per_launch / top_sec_per_launch / 1e12 / fp32_peak, 4),
            "counts": fma_counts,
            "algorithmic_tflops_all_pipes": round(top_flops_per_launch / top_sec_per_launch / 1e12, 2),
            "whole_step_tflops": round(B * G * (N - 1) * unet_flops / (ms * 1e-3) / 1e12, 2)}
    hbm = {"achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4)}
    # which roofline the dominant kernel sits closest to (largest fraction of its peak)
    fracs = {"hbm": hbm["frac"], "fp32": fp32["frac"], "tensor": tensor["frac"] if tensor else 0.0}
    bound = max(fracs, key=fracs.get)
    head = {"hbm": (hbm["achieved"], hbm_peak, "GB/s"), "fp32": (fp32["achieved_tflops"], fp32["peak_tflops"], "TFLOP/s"),
            "tensor": ((tensor or {}).get("achieved_tflops"), (tensor or {}).get("peak_tflops"), "TFLOP/s")}[bound]
    roofline = {"bound": bound, "kernel": top["kernel"], "achieved": head[0], "peak": head[1], "unit": head[2],
                "frac": fracs[bound], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "us_per_launch": round(top_sec_per_launch * 1e6, 1), "windows_per_launch": windows_per_launch,
                "algorithmic_bytes_per_launch": top_bytes_per_launch, "hbm": hbm, "fp32": fp32, "tensor": tensor,
                "note": "'bound' = the roofline of the dominant kernel with the largest achieved fraction (hbm: algorithmic activation bytes "
                        "in + out; fp32: position mixes on the FMA pipe + convolution flops; tensor: 3xTF32 MMAs of the 1x1 convolutions)",
                "whole_step_hbm": None if unet_bytes is None else {
                    "achieved": round(B * G * (N - 1) * unet_bytes / (ms * 1e-3) / 1e9, 1), "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(B * G * (N - 1) * unet_bytes / (ms * 1e-3) / 1e9 / hbm_peak, 4),
                    "note": "all kernels of a step: algorithmic bytes per window-step (SURVEY.md 8d) x window-steps / step time"}}

    cpu_baseline = None
    if world == 1 and not a.no_cpu_baseline:
        nw, ns = (int(x) for x in a.cpu_sample.split("x"))
        cv, cdt = cpu_reference_rate(a, nw, ns, steps=1, warmup=1)
        cpu_baseline = {"value": cv, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"oracle/ref_port.py (the reference's torch CPU operators): {nw} windows x {ns} samples x "
                                  f"{N - 1} steps in {cdt:.1f} s, scaled to {G} samples/window"}
        try:
            ns_e = max(1, min(G, a.eager_samples))
            ev, edt = cuda_eager_reference_rate(a, dev, ns_e)
            cpu_baseline["reference_cuda_eager"] = {
                "value": ev, "unit": UNIT, "device": torch.cuda.get_device_name(dev),
                "sample": f"oracle/ref_port.py (the reference's ATen operators in its order) as PyTorch CUDA eager on this GPU: "
                          f"{B} windows x {ns_e} samples x {N - 1} steps in {edt:.2f} s, scaled to {G} samples/window",
                "speedup_e2e": round(e2e_value / ev, 1)}
            cpu_baseline["sample"] += (f" || PyTorch CUDA-eager arm (same operators on this {torch.cuda.get_device_name(dev)}, B={B}, {ns_e} of {G} "
                                       f"samples timed): {ev:.1f} windows/s -> this path is {e2e_value / ev:.1f}x end to end")
        except Exception as exc:  # a reported baseline must never take the bench line down
            cpu_baseline["reference_cuda_eager"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    shipped_T3 = None
    if shipped:
        shipped_T3 = []
        for rec in shipped:
            r, sb = rec["r"], rec["B"]
            entry = {"workload": f"synthetic windows [B,2,3,17] (seg_len 6, every shipped config), B={sb}/GPU, noise_steps={N}, "
                                 f"n_generated_samples={G}", "windows_per_gpu_per_step": sb, "value": r["value"], "unit": UNIT,
                     "ms_per_step": r["ms"], "window_steps_per_sec": r["value"] * G * (N - 1),
                     "e2e": {"value": r["e2e_value"], "unit": UNIT, "ms_per_step": r["e2e_s"] * 1e3,
                             "h2d_bytes_per_step": r["host"].numel() * 4, "d2h_bytes_per_step": sb * 4},
                     "whole_step_hbm": {"achieved": round(rec["algo"] * r["value"] / world / 1e9, 1), "peak": hbm_peak, "unit": "GB/s",
                                        "frac": round(rec["algo"] * r["value"] / world / 1e9 / hbm_peak, 4)},
                     "gpu_launches": r["launches"]}
            if world == 1 and not a.no_cpu_baseline:
                try:
                    a3 = argparse.Namespace(**dict(vars(a), seg_len=6, batch=sb))
                    ev3, edt3 = cuda_eager_reference_rate(a3, dev, max(1, min(G, a.eager_samples)))
                    entry["reference_cuda_eager"] = {"value": ev3, "unit": UNIT, "speedup_e2e": round(r["e2e_value"] / ev3, 1),
                                                     "sample": f"oracle port as PyTorch CUDA eager, B={sb}, {max(1, min(G, a.eager_samples))} of {G} samples "
                                                               f"timed ({edt3:.2f} s), scaled"}
                except Exception as exc:
                    entry["reference_cuda_eager"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            shipped_T3.append(entry)
        if cpu_baseline is not None:
            cpu_baseline["sample"] += " || shipped shape T=3: " + "; ".join(
                f"B={e['windows_per_gpu_per_step']}: {e['e2e']['value']:.0f} windows/s e2e vs CUDA-eager "
                f"{e.get('reference_cuda_eager', {}).get('value', float('nan')):.0f} ({e.get('reference_cuda_eager', {}).get('speedup_e2e', 'n/a')}x)"
                for e in shipped_T3)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(a),
            "window_steps_per_sec": value * G * (N - 1),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": host.numel() * 4, "d2h_bytes_per_step": B * 4,
                    "ms_per_step": e2e_s * 1e3,
                    "api": "mcd_score_windows_host (pinned host windows in, host scores out)" +
                           (" + the all-gather of scores (device) and its copy back to the host" if world > 1 else "")},
            "gpu_launches": launches, "clocks": clocks_summary, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "shipped_T3": shipped_T3, "ingest": ingest, "kernels": kernels}
    _emit(out_fd, line)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
