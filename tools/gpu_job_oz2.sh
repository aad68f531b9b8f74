#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/oz_probe3.log
: > $L
for cfg in "mid 2" "big 0" "big 1" "big 2"; do
  timeout 120 python tools/oz_probe.py $cfg >> $L 2>&1 || echo "FAILED($?): $cfg" >> $L
done
cat $L
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -x -q -k "tcgen05" 2>&1 | tail -12
