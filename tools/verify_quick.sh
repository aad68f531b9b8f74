#!/bin/bash
# quick GPU check after an engine change: engine / slot / L-BFGS parity tests and a short bench line with the trajectory leg
cd ${GRAFT_REPO_ROOT:-.}
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_slots.py tests/test_gpu_lbfgs.py -q -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --c4-n 0 > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err; python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/verify_bench.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms'])
t=d['trajectory']; print('traj warm', t['ms_mean'], t['ms_median'], t['ms_per_step_list'], 'cold', t['cold_first_trip']['ms_mean'])
print('traffic', d['roofline']['traffic'])
PY
