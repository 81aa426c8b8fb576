#!/bin/bash
# Profiling pass for profiles/ (round 2): (1) launch list of one bench step (kernel shares), (2) ncu --set full of one
# denoiser call AT THE BENCH'S LAUNCH SIZE (17 168 windows per launch at [B,2,24,17], B=1024, G=50), summarised on the box
# (the .ncu-rep files stay there: gpurun_out/ is limited to 64 MiB).
# usage (on the GPU box, from the repo root): bash tools/gpu_profile.sh <tag>
tag=${1:-prof}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-shipped"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_launches.log 2>&1; echo "launch list rc=$?"
python profiles/summarize_ncu.py launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt; head -20 gpurun_out/${tag}_launches_summary.txt
# one denoiser call = time_embedding + 11 blocks + 4 resamples = 16 launches (the DDPM update is fused into the last block);
# skip the encoder (5), randn (1) and the first call (16)
timeout 900 ncu --set full --clock-control none --launch-skip 22 -c 16 -f -o /tmp/${tag}_full $CMD > gpurun_out/${tag}_full.log 2>&1; echo "full rc=$?"
python profiles/summarize_ncu.py full /tmp/${tag}_full.ncu-rep 17168 gpurun_out/${tag}_ncu_traffic.json > gpurun_out/${tag}_ncu_summary.txt; echo "summary rc=$?"
ls -la gpurun_out | grep ${tag}
