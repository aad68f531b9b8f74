cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/prof_dist.py > gpurun_out/r2_prof_dist2.txt 2>&1; tail -32 gpurun_out/r2_prof_dist2.txt
