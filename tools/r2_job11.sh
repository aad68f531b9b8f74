set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_dist.py -q > gpurun_out/r2_t11.log 2>&1; tail -5 gpurun_out/r2_t11.log
timeout 600 python bench.py --workload c4 --steps 3 --warmup 1 > gpurun_out/r2_c4_1gpu.json 2> gpurun_out/r2_c4_1gpu.err; cat gpurun_out/r2_c4_1gpu.json; tail -3 gpurun_out/r2_c4_1gpu.err
