set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_slots.py tests/test_gpu_lbfgs.py tests/test_gpu_batch.py -q > gpurun_out/r2_t10.log 2>&1; tail -25 gpurun_out/r2_t10.log
timeout 900 python -m pytest tests/test_gpu_headline.py -q -k "config5_full or syrk" > gpurun_out/r2_t10b.log 2>&1; tail -5 gpurun_out/r2_t10b.log
