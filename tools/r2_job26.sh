cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_dist.py -q -x > gpurun_out/r2_t26.log 2>&1; tail -3 gpurun_out/r2_t26.log
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/r2_c4_n1.json 2> gpurun_out/r2_c4_n1.err; tail -2 gpurun_out/r2_c4_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c4 --steps 3 --warmup 3 > gpurun_out/r2_c4_n2.json 2> gpurun_out/r2_c4_n2.err; tail -2 gpurun_out/r2_c4_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_c4_n1.json', 'gpurun_out/r2_c4_n2.json'):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][0]
        print(f, d['n_gpus'], 'factor', round(d['factor_ms'],2), 'solve', round(d['solve_ms_8rhs_1refine'],2), d['inertia'], d['scaled_residual_inf'], d.get('pipeline_on_one_rank'), (d.get('native') or {}).get('scaled_residual_inf'))
    except Exception as e: print(f, e)
PY
