#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/oz_probe2.log
: > $L
for cfg in "mid 0" "mid 1" "big 0" "big 1"; do
  timeout 120 python tools/oz_probe.py $cfg >> $L 2>&1 || echo "FAILED($?): $cfg" >> $L
done
cat $L
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -x -q -k "tcgen05" 2>&1 | tail -12
