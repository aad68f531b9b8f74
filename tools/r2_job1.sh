set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
nproc; free -g | head -2
timeout 1500 python -m pytest tests/test_gpu_headline.py tests/test_gpu_slots.py tests/test_gpu_engine.py::test_callable_mode_nonconvex_delta_trace -q --durations=15 > gpurun_out/r2_t1.log 2>&1
tail -40 gpurun_out/r2_t1.log
