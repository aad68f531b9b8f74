#!/bin/bash
# config 4 alone on N GPUs of one node: bash tools/c4_run.sh N  ->  gpurun_out/c4_nN.json
cd ${GRAFT_REPO_ROOT:-.}
NG=${1:-4}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --workload c4 --steps 5 --warmup 3 > gpurun_out/c4_n${NG}.json 2> gpurun_out/c4_n${NG}.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/c4_n${NG}.json') if l.startswith('{')][0]
print(d['n_gpus'], 'factor', round(d['factor_ms'],2), 'solve', round(d['solve_ms_8rhs_1refine'],2), d['inertia'], d['scaled_residual_inf'], round(d['factor_tflops_aggregate'],1), d.get('solve_frac_of_hbm_peak'))
PY
