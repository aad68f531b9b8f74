set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_slots.py tests/test_gpu_lbfgs.py -q -x > gpurun_out/r2_t6.log 2>&1; tail -3 gpurun_out/r2_t6.log
for f in 6 262; do
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --flags $f > gpurun_out/r2_bench_f$f.json 2> gpurun_out/r2_bench_f$f.err; python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_f$f.json'))
print($f, d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms'], d['gpu_launches'])
t=d['trajectory']; print(t['ms_per_step_list']); print(t['soc_tried'], t['cert_used'], t['factorisations_physical']); print(t['phase_ms'])
PY
done
