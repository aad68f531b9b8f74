set -x
cd $GRAFT_REPO_ROOT
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --traj-steps 0 --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -3 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
print(json.dumps(d['c4']))
PY
