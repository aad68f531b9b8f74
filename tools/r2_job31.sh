cd $GRAFT_REPO_ROOT
NG=${1:-4}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 tools/check_dist.py 8192 2>&1 | grep -E "rank|Error|error" | head -20
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 tools/check_dist.py 16384 2>&1 | grep -E "rank|Error|error" | head -20
