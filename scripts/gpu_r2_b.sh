#!/bin/bash
# round-2 run B: the new tools on the GPU (parity tests), BAM writer timing, config 4 bench line
set -u
mkdir -p gpurun_out
python -m pytest tests/test_bam_writers.py tests/test_bigwig_tools.py tests/test_host.py -m gpu -x -q > gpurun_out/r2_gpu_tests5.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/r2_gpu_tests5.log
timeout 600 python scripts/bam_writer_bench.py 10000000 > gpurun_out/r2_bam_writers_10M.json 2> gpurun_out/r2_bam_writers_10M.err; echo "writer bench rc $?"; cat gpurun_out/r2_bam_writers_10M.json; grep write_bam gpurun_out/r2_bam_writers_10M.err | tail -8
timeout 600 python bench.py --config 4 --pairs 20000000 --steps 2 --warmup 3 > gpurun_out/r2_bench_cfg4_20M.json 2> gpurun_out/r2_bench_cfg4_20M.err; echo "cfg4 rc $?"; cat gpurun_out/r2_bench_cfg4_20M.json | cut -c1-1500
