cd $GRAFT_REPO_ROOT
NG=${1:-4}
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -k "ldlt" 2>&1 | tail -2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 tools/check_dist.py 16384 2>&1 | grep -E "rank|Error|error" | cut -c1-400 | head -20
