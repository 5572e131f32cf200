#!/bin/bash
# round-2 run I: GPU suite and the default bench line on the final tree (K1 with the parallel table build)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests10.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2_gpu_tests10.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke4.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r2_smoke4.log
BND_DEBUG_TIMING=1 BND_BENCH_TRACE=1 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc $?"; cut -c1-1000 gpurun_out/r2_bench8.json; grep -E "trace|to_bedgraph:" gpurun_out/r2_bench8.err | tail -4
