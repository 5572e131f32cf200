#!/bin/bash
# round-2 run E: K1 occupancy A/B (33 and 36 warps per SM against the shipped 30) and a larger segment, 40 M pairs, no profiler
set -u
mkdir -p gpurun_out
export PAIRS=40000000
for v in 0 13 14 15 16; do BND_INFLATE_VARIANT=$v python scripts/sweep_pipeline.py 2>&1 | tail -1; done | tee gpurun_out/r2_k1_ab_run4.log
