#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/oz_probe5.log
: > $L
for cfg in "mid 129" "big 129" "big 97"; do
  timeout 120 python tools/oz_probe.py $cfg >> $L 2>&1 || echo "FAILED($?): $cfg" >> $L
done
cat $L
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -x -q -k "tcgen05" 2>&1 | tail -3
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/s13_bench.json"))
print(round(d["value"],2), "steps/s  e2e", round(d["e2e"]["value"],2), {k: round(v,3) for k,v in d["phase_ms"].items()}, "resid", d["kkt_residual_inf"])
print("   ", {k: round(v["ms"],3) for k,v in d["kernels"].items() if "tcgen05" in k}, d["roofline"]["frac"])
PY
