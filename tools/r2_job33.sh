cd $GRAFT_REPO_ROOT
timeout 200 python tools/trace_factor.py 3 2>&1 | grep -E "factor ms|periods|mini"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --c4-n 0 --traj-steps 0 > gpurun_out/r2_bench_i.json 2> gpurun_out/r2_bench_i.err; python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2_bench_i.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phase_ms'], d['gpu_launches'])
print(d['roofline']['frac'], d['roofline']['ms'], d['roofline_hbm']['triangular_solve']['ms'])
PY
