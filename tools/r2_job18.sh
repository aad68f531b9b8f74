set -x
cd $GRAFT_REPO_ROOT
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DTILE_PROF -o /tmp/libprof.so pyipm_b200/csrc/b200ipm.cu
B200IPM_LIB=/tmp/libprof.so python tools/tile_prof.py 2>&1 | tail -40
timeout 900 python -m pytest tests/test_gpu_engine.py -q -x -k "adversarial or own_init" > gpurun_out/r2_t18.log 2>&1; tail -12 gpurun_out/r2_t18.log
