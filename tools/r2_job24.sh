cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "ldlt" > gpurun_out/r2_t24.log 2>&1; tail -3 gpurun_out/r2_t24.log
for mode in 0 1; do
echo "--- binv_mode=$mode"; B200IPM_BINV_MODE=$mode timeout 200 python tools/trace_factor.py 3 2>&1 | grep -E "factor ms|periods"
done
echo "--- solve256=0"; B200IPM_SOLVE256=0 timeout 200 python tools/trace_factor.py 3 2>&1 | grep -E "factor ms|periods"
for mode in 0 1; do
B200IPM_BINV_MODE=$mode python bench.py --steps 20 --warmup 3 --no-cpu-baseline --c4-n 0 --traj-steps 0 > gpurun_out/r2_bench_h$mode.json 2> gpurun_out/r2_bench_h$mode.err; python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2_bench_h$mode.json') if l.startswith('{')][0]
print($mode, d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms'], d['gpu_launches'])
print(d['roofline']['frac'], d['roofline']['ms'], d['roofline_hbm']['triangular_solve']['ms'])
PY
done
