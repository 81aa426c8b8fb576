#!/bin/bash
# A/B of experiment builds on ONE GPU box: for each library variant (mocodad_b200/libmocodad_b200_<tag>.so, "base" = the default
# library) run the GPU parity suite and the T=24 bench line.  usage: bash tools/ab_variants.sh <run-tag> base tq5 tq6 ...
run=$1; shift
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = base ]; then unset MOCODAD_B200_LIB; else export MOCODAD_B200_LIB=$PWD/mocodad_b200/libmocodad_b200_${tag}.so; fi
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${run}_${tag}_pytest.log 2>&1; echo "$tag pytest rc=$? $(tail -1 gpurun_out/${run}_${tag}_pytest.log)"
  for rep in 1 2; do
    timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${AB_BENCH_ARGS:---no-shipped} > gpurun_out/${run}_${tag}_bench${rep}.json 2> gpurun_out/${run}_${tag}_bench${rep}.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${run}_${tag}_bench${rep}.json").read().strip().splitlines()[-1])
    ks = {k["kernel"]: round(k["ms"], 2) for k in d["kernels"][:16]}
    print("$tag rep$rep value", round(d["value"], 1), "clk", d["clocks"]["sm_mhz"], ks, [(r["windows_per_gpu_per_step"], round(r["value"])) for r in (d.get("shipped_T3") or [])])
except Exception as e: print("$tag bench parse failed", e)
PY
  done
done
