set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_batch.py -q -s > gpurun_out/r2_t7.log 2>&1; tail -40 gpurun_out/r2_t7.log
python tools/trace_factor.py 3 > gpurun_out/r2_trace3.log 2>&1; tail -14 gpurun_out/r2_trace3.log
