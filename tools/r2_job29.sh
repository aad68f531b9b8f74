cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_dist.py -q -x -k nccl > gpurun_out/r2_t29.log 2>&1; tail -2 gpurun_out/r2_t29.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/prof_dist.py > gpurun_out/r2_prof_dist2b.txt 2>&1; grep -E "rank|^ +[0-9] own" gpurun_out/r2_prof_dist2b.txt | head -12 | cut -c1-220
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c4 --steps 5 --warmup 3 > gpurun_out/r2_c4_n2b.json 2> gpurun_out/r2_c4_n2b.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_c4_n2b.json',):
    d=[json.loads(l) for l in open(f) if l.startswith('{')][0]
    print(f, d['n_gpus'], 'factor', round(d['factor_ms'],2), 'solve', round(d['solve_ms_8rhs_1refine'],2), d['inertia'], d['scaled_residual_inf'])
PY
