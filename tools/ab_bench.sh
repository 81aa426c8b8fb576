#!/bin/bash
# bench-only A/B of library variants (no pytest).  usage: bash tools/ab_bench.sh <run-tag> "<bench args>" base v1 v2 ...
run=$1; args=$2; shift; shift
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = base ]; then unset MOCODAD_B200_LIB; else export MOCODAD_B200_LIB=$PWD/mocodad_b200/libmocodad_b200_${tag}.so; fi
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-shipped $args > gpurun_out/${run}_${tag}.json 2> gpurun_out/${run}_${tag}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${run}_${tag}.json").read().strip().splitlines()[-1])
    print("$tag value", round(d["value"], 1), "clk", d["clocks"]["sm_mhz"], {k["kernel"]: round(k["ms"], 2) for k in d["kernels"][:9]})
except Exception as e: print("$tag bench parse failed", e)
PY
done
