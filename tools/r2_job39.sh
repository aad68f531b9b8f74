cd $GRAFT_REPO_ROOT
NG=8
for dfr in 0 1; do
echo "== DEFER=$dfr"
B200IPM_DIST_DEFER=$dfr timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2951$dfr tools/prof_dist.py > gpurun_out/r2_prof_dist8_defer$dfr.txt 2>&1; grep -E "rank 0|^ +(8|9|16|17|24|25) own" gpurun_out/r2_prof_dist8_defer$dfr.txt | cut -c1-220
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 tools/check_dist.py 16384 2>&1 | grep -E "rank|Error|error" | cut -c1-300 | head -6
