#!/bin/bash
# per-kernel bench line at the shipped shape (T=3, B=1024) for library variants.  usage: bash tools/ab_t3.sh <run-tag> base nocf ...
run=$1; shift
for tag in "$@"; do
  if [ "$tag" = base ]; then unset MOCODAD_B200_LIB; else export MOCODAD_B200_LIB=$PWD/mocodad_b200/libmocodad_b200_${tag}.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-shipped --seg-len 6 > gpurun_out/${run}_${tag}_T3.json 2> gpurun_out/${run}_${tag}_T3.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${run}_${tag}_T3.json").read().strip().splitlines()[-1])
    print("$tag T3 value", round(d["value"], 1), "clk", d["clocks"]["sm_mhz"], {k["kernel"]: round(k["ms"], 2) for k in d["kernels"][:16]})
except Exception as e: print("$tag T3 bench parse failed", e)
PY
done
