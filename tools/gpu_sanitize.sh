#!/bin/bash
# compute-sanitizer over the scoring path with several tiles per persistent CTA (tools/sanitize_case.py):
# memcheck, racecheck (shared-memory hazards between the warp roles) and synccheck (barrier misuse).
# usage (on the GPU box, from the repo root): bash tools/gpu_sanitize.sh <tag> [cases...]     logs -> gpurun_out/<tag>_*.log
tag=${1:-san}; shift
cases=${@:-T24 T3 latent}
mkdir -p gpurun_out
for c in $cases; do
  timeout 300 python tools/sanitize_case.py $c > gpurun_out/${tag}_plain_$c.log 2>&1; echo "plain $c rc=$?"; tail -1 gpurun_out/${tag}_plain_$c.log
  for tool in memcheck racecheck synccheck; do
    timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py $c > gpurun_out/${tag}_${tool}_$c.log 2>&1
    echo "$tool $c rc=$?"; grep -E "checksum|ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/${tag}_${tool}_$c.log | tail -4
  done
done
