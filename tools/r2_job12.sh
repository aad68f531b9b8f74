set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_dist.py -q -k nccl > gpurun_out/r2_t12.log 2>&1; tail -8 gpurun_out/r2_t12.log
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --workload c4 --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_c4_${N}gpu.json 2> gpurun_out/r2_c4_${N}gpu.err; cat gpurun_out/r2_c4_${N}gpu.json; tail -5 gpurun_out/r2_c4_${N}gpu.err
