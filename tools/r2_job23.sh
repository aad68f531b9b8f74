cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "ldlt" > gpurun_out/r2_t23.log 2>&1; tail -5 gpurun_out/r2_t23.log
echo "--- solve256=1"; timeout 200 python tools/prof_solve.py 2>&1 | tail -2
echo "--- solve256=0"; B200IPM_SOLVE256=0 timeout 200 python tools/prof_solve.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_headline.py -q -x > gpurun_out/r2_t23b.log 2>&1; tail -5 gpurun_out/r2_t23b.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --c4-n 0 --traj-steps 0 > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2_bench_g.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms'], d['gpu_launches'])
print(d['roofline']['frac'], d['roofline']['ms'], d['roofline_hbm']['triangular_solve'])
PY
