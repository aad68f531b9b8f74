set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -x > gpurun_out/r2_t17.log 2>&1; tail -3 gpurun_out/r2_t17.log
timeout 300 python tools/trace_factor.py 3 2>&1 | tail -10
