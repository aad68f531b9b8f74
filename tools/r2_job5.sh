set -x
cd $GRAFT_REPO_ROOT
python tools/prof_solve.py 2>&1 | tail -3
B200IPM_SOLVE_GROUPED=0 python tools/prof_solve.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dist.py tests/test_gpu_engine.py -q -x > gpurun_out/r2_t5.log 2>&1; tail -5 gpurun_out/r2_t5.log
B200IPM_SOLVE_GROUPED=0 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -x > gpurun_out/r2_t5b.log 2>&1; tail -3 gpurun_out/r2_t5b.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_b.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms'], d['trajectory']['ms_per_step_list'], d['gpu_launches'])
PY
