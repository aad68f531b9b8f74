set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dist.py -q -x -k "ldlt or block" > gpurun_out/r2_t4.log 2>&1; tail -5 gpurun_out/r2_t4.log
python tools/prof_solve.py 2>&1 | tail -3
B200IPM_SOLVE_GROUPED=0 python tools/prof_solve.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_engine.py -q -x > gpurun_out/r2_t4b.log 2>&1; tail -5 gpurun_out/r2_t4b.log
