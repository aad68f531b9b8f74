cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_slots.py tests/test_gpu_lbfgs.py -q -x 2>&1 | tail -8
