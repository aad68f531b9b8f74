set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_lbfgs.py tests/test_gpu_slots.py tests/test_gpu_engine.py -q --durations=8 > gpurun_out/r2_t2.log 2>&1
tail -60 gpurun_out/r2_t2.log
