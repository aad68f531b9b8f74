#!/bin/bash
# first hardware run of the tcgen05 path: descriptor probe, then regression tests + bench
mkdir -p gpurun_out
L=gpurun_out/oz_probe.log
: > $L
for cfg in "tiny 0 128 256" "tiny 0 256 128" "small 0 128 256" "small 0 256 128" "tiny 1 128 256" "tiny 1 256 128"; do
  timeout 90 python tools/oz_probe.py $cfg >> $L 2>&1 || echo "FAILED($?): $cfg" >> $L
done
cat $L
