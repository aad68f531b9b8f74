#!/bin/bash
# end-of-session evidence run: full GPU test suite, bench (both arms), ncu launch list + full capture of the top kernel
mkdir -p gpurun_out
TAG=${1:-s2}
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print(round(d["value"],2), "steps/s  e2e", round(d["e2e"]["value"],2), {k: round(v,3) for k,v in d["phase_ms"].items()})
print("roofline", {k: d["roofline"][k] for k in ("achieved","peak","frac")}, "hbm", d["roofline_hbm"]["frac"], "cpu", d.get("cpu_baseline",{}).get("value"))
print("clocks", d["clocks"], "launches", d["gpu_launches"])
r=json.load(open("gpurun_out/${TAG}_bench_ref.json")); print("ref arm", r["value"], r["cpu_baseline"]["sample"][:120])
PY
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_step.py > gpurun_out/${TAG}_prof_step.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:oz_syrk -c 1 -f -o gpurun_out/${TAG}_prof_oz python tools/prof_step.py > /dev/null 2>&1
ls -la gpurun_out | tail -8
