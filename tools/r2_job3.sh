set -x
cd $GRAFT_REPO_ROOT
python tools/make_c3_state.py > gpurun_out/r2_state.log 2>&1
tail -2 gpurun_out/r2_state.log
mkdir -p tests/golden && cp gpurun_out/c3_state_after3.npz tests/golden/c3_state_after3.npz
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -c 3000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
( time python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2>&1 | tail -3
cat gpurun_out/r2_bench_ref.json; tail -3 gpurun_out/r2_bench_ref.err
