#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -x -q -k "tcgen05 or ldlt or speculative" 2>&1 | tail -5
for f in 0 6; do
  timeout 200 python bench.py --no-cpu-baseline --flags $f > gpurun_out/s4_bench_f$f.json 2> gpurun_out/s4_bench_f$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/s4_bench_f$f.json"))
print("flags=$f", round(d["value"],2), "steps/s  e2e", round(d["e2e"]["value"],2), {k: round(v,3) for k,v in d["phase_ms"].items()})
print("   ", {k: round(v["ms"],3) for k,v in d["kernels"].items()})
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:oz_syrk -c 1 -f -o gpurun_out/prof_oz_v1 python tools/oz_probe.py big 1 > gpurun_out/ncu_oz.log 2>&1
tail -3 gpurun_out/ncu_oz.log
