#!/bin/bash
# round-2 run E: K1 occupancy A/B (33 and 36 warps per SM against the shipped 30) and a larger segment, 40 M pairs, no profiler
set -u
mkdir -p gpurun_out
export PAIRS=40000000
for v in 0 4 5; do BND_INFLATE_VARIANT=$v python scripts/sweep_pipeline.py 2>&1 | tail -1; done | tee gpurun_out/r2_k1_occupancy.log
BND_SEGMENT_BYTES=150994944 python scripts/sweep_pipeline.py 2>&1 | tail -1 | tee -a gpurun_out/r2_k1_occupancy.log
BND_SEGMENT_BYTES=201326592 BND_INFLATE_VARIANT=0 python scripts/sweep_pipeline.py 2>&1 | tail -1 | tee -a gpurun_out/r2_k1_occupancy.log
