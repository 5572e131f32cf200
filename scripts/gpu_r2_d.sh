#!/bin/bash
# round-2 run D: default bench line after the teardown fix, BAM writer timing with overlapped write and per-field lags, writer tests
set -u
mkdir -p gpurun_out
python -m pytest tests/test_bam_writers.py tests/test_outputs.py -m gpu -x -q > gpurun_out/r2_gpu_tests7.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2_gpu_tests7.log
BND_DEBUG_TIMING=1 BND_BENCH_TRACE=1 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc $?"; cut -c1-1100 gpurun_out/r2_bench5.json; grep -E "trace|to_bedgraph:|destroy" gpurun_out/r2_bench5.err | tail -6
timeout 600 python scripts/bam_writer_bench.py 10000000 > gpurun_out/r2_bam_writers_10M_c.json 2> gpurun_out/r2_bam_writers_10M_c.err; echo "writer bench rc $?"; cat gpurun_out/r2_bam_writers_10M_c.json; grep write_bam gpurun_out/r2_bam_writers_10M_c.err | tail -5
