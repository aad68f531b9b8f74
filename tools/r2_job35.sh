cd $GRAFT_REPO_ROOT
NG=${1:-2}
timeout 300 python -m pytest tests/test_gpu_dist.py -q -x 2>&1 | tail -2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 tools/check_dist.py 16384 2>&1 | grep -E "rank|Error|error" | cut -c1-300 | head -8
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/prof_dist.py > gpurun_out/r2_prof_dist${NG}c.txt 2>&1; grep -E "rank|^ +[0-9]+ own" gpurun_out/r2_prof_dist${NG}c.txt | head -8 | cut -c1-220
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --workload c4 --steps 5 --warmup 3 > gpurun_out/r2_c4_n${NG}c.json 2> gpurun_out/r2_c4_n${NG}c.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2_c4_n${NG}c.json') if l.startswith('{')][0]
print(d['n_gpus'], 'factor', round(d['factor_ms'],2), 'solve', round(d['solve_ms_8rhs_1refine'],2), d['inertia'], d['scaled_residual_inf'])
PY
