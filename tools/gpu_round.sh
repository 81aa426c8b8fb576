#!/bin/bash
# One GPU-box round trip: parity tests, bench line, ncu launch list, ncu source-level capture of the TC block kernels.
# usage (from the repo root, on the GPU box): bash tools/gpu_round.sh <tag> [tests|notests]
tag=${1:-run}; mode=${2:-tests}
mkdir -p gpurun_out
if [ "$mode" = tests ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
  tail -3 gpurun_out/${tag}_pytest.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_T24.json 2> gpurun_out/${tag}_bench_T24.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_T24.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
    for k in d["kernels"]: print(k)
except Exception as e: print("bench parse failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stgcn_block_tc --launch-skip 8 -c 8 -f -o gpurun_out/${tag}_tc \
  python bench.py --steps 1 --warmup 1 --batch 296 --gen 8 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
