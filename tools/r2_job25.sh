cd $GRAFT_REPO_ROOT
timeout 200 python tools/trace_factor.py 3 0 250 1100 2>&1 | tail -n +2 > gpurun_out/r2_trace25.txt; head -12 gpurun_out/r2_trace25.txt
