#!/bin/bash
# round-2 run E: K1 A/B timing at 40 M pairs, no profiler (profiles/r2_sweeps.md).  The sweep ran in four calls over shapes that
# were compiled in for the purpose (warps per SM, span bits, match-list entries) and removed again; what is left to compare
# are the shipped shape (0) and the previous default (4).
set -u
mkdir -p gpurun_out
export PAIRS=40000000
for v in 0 4; do BND_INFLATE_VARIANT=$v python scripts/sweep_pipeline.py 2>&1 | tail -1; done | tee gpurun_out/r2_k1_ab.log
