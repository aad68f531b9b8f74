set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "tcgen05_trailing" > gpurun_out/r2_t9.log 2>&1; tail -15 gpurun_out/r2_t9.log
timeout 300 python tools/prof_solve.py 2>&1 | tail -2
B200IPM_LDLT_TC=0 timeout 300 python tools/prof_solve.py 2>&1 | tail -2
timeout 300 python tools/trace_factor.py 3 2>&1 | tail -12
