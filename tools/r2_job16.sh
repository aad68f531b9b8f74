set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -x > gpurun_out/r2_t16.log 2>&1; tail -3 gpurun_out/r2_t16.log
timeout 300 python tools/trace_factor.py 3 2>&1 | tail -10
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --c4-n 0 --traj-steps 0 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2_bench_d.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms'], d['gpu_launches'])
print(d['roofline']['frac'], d['roofline']['ms'])
PY
