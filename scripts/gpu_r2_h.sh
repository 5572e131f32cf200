#!/bin/bash
# round-2 run H: the GPU suite on the final tree, split at the BASELINE size with exact-length flushes
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests9.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2_gpu_tests9.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo "bench rc $?"; cut -c1-300 gpurun_out/r2_bench7.json
timeout 600 python scripts/bam_writer_full.py > gpurun_out/r2_bam_split_full2.json 2> gpurun_out/r2_bam_split_full2.err; echo "full split rc $?"; cat gpurun_out/r2_bam_split_full2.json; grep write_bam gpurun_out/r2_bam_split_full2.err | tail -2
