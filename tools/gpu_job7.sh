#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_dist.py -x -q 2>&1 | tail -4
for f in 6 70; do
  timeout 200 python bench.py --no-cpu-baseline --flags $f > gpurun_out/s10_bench_f$f.json 2> gpurun_out/s10_bench_f$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/s10_bench_f$f.json"))
print("flags=$f", round(d["value"],2), "steps/s  e2e", round(d["e2e"]["value"],2), {k: round(v,3) for k,v in d["phase_ms"].items()})
PY
done
python tools/first_attempt.py 6 10 2>&1 | tail -10 | cut -c1-260
