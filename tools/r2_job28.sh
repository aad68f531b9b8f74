cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_dist.py -q -x > gpurun_out/r2_t28.log 2>&1; tail -3 gpurun_out/r2_t28.log
timeout 300 python tools/prof_dist.py > gpurun_out/r2_prof_dist1.txt 2>&1; tail -14 gpurun_out/r2_prof_dist1.txt
