#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/oz_probe4.log
: > $L
for cfg in "big 129" "big 113" "big 97" "big 0" "big 2"; do
  timeout 120 python tools/oz_probe.py $cfg >> $L 2>&1 || echo "FAILED($?): $cfg" >> $L
done
cat $L
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_dist.py -x -q 2>&1 | tail -4
for f in 6 134; do
  timeout 200 python bench.py --no-cpu-baseline --flags $f > gpurun_out/s12_bench_f$f.json 2> gpurun_out/s12_bench_f$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/s12_bench_f$f.json"))
print("flags=$f", round(d["value"],2), "steps/s  e2e", round(d["e2e"]["value"],2), {k: round(v,3) for k,v in d["phase_ms"].items()}, "resid", d["kkt_residual_inf"])
print("   ", {k: round(v["ms"],3) for k,v in d["kernels"].items() if "tcgen05" in k}, d["roofline"]["frac"])
PY
done
