#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_dist.py -x -q 2>&1 | tail -4
for mini in 1 0; do
  B200IPM_LDLT_MINI=$mini timeout 200 python bench.py --no-cpu-baseline > gpurun_out/s7_bench_m$mini.json 2> gpurun_out/s7_bench_m$mini.err
  python - <<PY
import json
d=json.load(open("gpurun_out/s7_bench_m$mini.json"))
print("mini=$mini", round(d["value"],2), "steps/s  e2e", round(d["e2e"]["value"],2), {k: round(v,3) for k,v in d["phase_ms"].items()}, "ldlt single", round(d["kernels"]["ldlt_factor"]["ms"],3))
PY
done
