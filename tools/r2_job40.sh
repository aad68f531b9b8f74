cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -x -k "ldlt or adversarial or teacher or golden or step or solve" 2>&1 | tail -3
timeout 200 python tools/trace_factor.py 3 2>&1 | grep -E "factor ms|tile  |mini "
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --c4-n 0 --traj-steps 0 > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err; python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2_bench_k.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms'], d['gpu_launches'])
PY
