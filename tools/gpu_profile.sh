#!/bin/bash
# Profiling pass for profiles/: (1) launch list of one bench step (kernel shares), (2) ncu --set full of one denoiser call.
# usage (on the GPU box, from the repo root): bash tools/gpu_profile.sh <tag>
tag=${1:-prof}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 296 --gen 8 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_launches.log 2>&1; echo "launch list rc=$?"
# one denoiser call = time_embedding + 11 blocks + 4 resamples + ddpm = 17 launches; skip the encoder (5), randn (1) and the first call
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 23 -c 17 -f -o gpurun_out/${tag}_full $CMD > gpurun_out/${tag}_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | grep ${tag}
# (3) the window-ingest kernels: plain timing first (CUDA events, never under ncu), then ncu --set full
timeout 300 python tools/ingest_profile.py --time > gpurun_out/${tag}_ingest_time.log 2>&1; tail -2 gpurun_out/${tag}_ingest_time.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'normalize_frames|build_items' -c 4 -f -o gpurun_out/${tag}_ingest \
  python tools/ingest_profile.py > gpurun_out/${tag}_ingest.log 2>&1; echo "ingest rc=$?"
