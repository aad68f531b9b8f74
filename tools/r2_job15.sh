set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -x > gpurun_out/r2_t15.log 2>&1; tail -3 gpurun_out/r2_t15.log
timeout 300 python tools/trace_factor.py 3 2>&1 | tail -10
B200IPM_LDLT_SPLITA=0 timeout 300 python tools/prof_solve.py 2>&1 | tail -1
timeout 300 python tools/prof_solve.py 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --c4-n 0 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2_bench_c.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms'], d['gpu_launches'])
t=d['trajectory']; print(t['ms_per_step_list'])
print(d['roofline']['frac'], d['roofline']['ms'])
PY
