cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_headline.py -q -x -k "syrk or ldlt or mu_sweep" 2>&1 | tail -3
timeout 200 python tools/trace_factor.py 3 2>&1 | grep -E "factor ms|periods|oz "
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --c4-n 0 --traj-steps 0 > gpurun_out/r2_bench_j.json 2> gpurun_out/r2_bench_j.err; python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2_bench_j.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms'], d['gpu_launches'])
print(d['roofline']['frac'], d['roofline']['ms'], d['kernels']['hess_syrk_tcgen05'], d['kernels']['condense_syrk_tcgen05'], d['kernels'].get('hess_syrk_tcgen05_kernel_only'))
PY
