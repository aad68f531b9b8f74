cd $GRAFT_REPO_ROOT
NG=${1:-8}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/prof_dist.py > gpurun_out/r2_prof_dist${NG}d.txt 2>&1; grep -E "rank 0|^ +[0-9]+ own" gpurun_out/r2_prof_dist${NG}d.txt | head -20 | cut -c1-220
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/r2_bench_${NG}gpu_final.json 2> gpurun_out/r2_bench_${NG}gpu_final.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2_bench_${NG}gpu_final.json') if l.startswith('{')][0]
c=d['c4']
print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])
print('c4 factor', round(c['factor_ms'],2), 'solve', round(c['solve_ms_8rhs_1refine'],2), c['inertia'], c['scaled_residual_inf'], c['factor_tflops_aggregate'], c.get('frac_of_aggregate_fp64_peak'))
PY
