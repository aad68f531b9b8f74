#!/bin/bash
# Round-2 evidence run (one GPU): smoke, full GPU test suite, default bench line, steady-state launch list of one
# Newton step at config 3 and one `ncu --set full` capture per kernel family of that step.  Outputs -> gpurun_out/${TAG}_*.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
TAG=${1:-r2e}
WHAT=${2:-all}
if [ "$WHAT" = all ] || [ "$WHAT" = tests ]; then
  timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_tests.log
  tail -3 gpurun_out/${TAG}_tests.log
fi
if [ "$WHAT" = all ] || [ "$WHAT" = bench ]; then
  timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench.json") if l.startswith('{')][0]
print(round(d["value"],2), "steps/s  e2e", round(d["e2e"]["value"],2), {k: round(v,3) for k,v in d["phase_ms"].items()})
print("roofline", {k: d["roofline"][k] for k in ("achieved","peak","frac","ms")}, "solve", d["roofline_hbm"]["triangular_solve"]["ms"])
print("clocks", d["clocks"], "launches", d["gpu_launches"], "c4", d.get("c4"))
print("traj", d.get("trajectory",{}).get("mean_ms"))
PY
fi
if [ "$WHAT" = all ] || [ "$WHAT" = ncu ]; then
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_step.py > gpurun_out/${TAG}_prof_step.log 2>&1
  python tools/make_traffic.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_traffic.json
  # one `ncu --set full` capture per kernel family; the reports are summarised ON THE BOX (gpurun copies back at most 64 MiB)
  # and only two of them travel for source-level reading
  mkdir -p /tmp/ncu_reps
  for k in oz_syrk_kernel ldlt_tile_kernel ldlt_mini_kernel ldlt_panel_kernel gemm_nt_sub64_tma_kernel ldlt_fwd256_kernel \
           ldlt_bwd256_kernel ldlt_blockinv_kernel gemv_n_kernel oz_slice_kernel oz_syrk_kernel_64; do
    name=$k
    skip=0
    # the in-place LDL^T update variant <64,6> of the tcgen05 kernel: the third oz_syrk launch of a step
    if [ $k = oz_syrk_kernel_64 ]; then k=oz_syrk_kernel; skip=2; fi
    timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$k" --launch-skip $skip -c 1 -f \
        -o /tmp/ncu_reps/${TAG}_ncu_$name python tools/prof_step.py > /dev/null 2>&1
    python tools/summarize_ncu.py /tmp/ncu_reps/${TAG}_ncu_$name.ncu-rep "Round 2 ncu --set full: $k (first launch of a steady-state Newton step at config 3)" \
        > gpurun_out/${TAG}_ncu_$name.md 2>/dev/null
  done
  cp /tmp/ncu_reps/${TAG}_ncu_oz_syrk_kernel_64.ncu-rep /tmp/ncu_reps/${TAG}_ncu_ldlt_fwd256_kernel.ncu-rep gpurun_out/ 2>/dev/null
fi
ls -la gpurun_out | grep ${TAG}_ | tail -20
