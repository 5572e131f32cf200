#!/bin/bash
# round-2 run C: the whole GPU suite after the text kernel / teardown / writer changes, the default bench line, BAM writer timing
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests6.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2_gpu_tests6.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke2.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r2_smoke2.log
BND_DEBUG_TIMING=1 BND_BENCH_TRACE=1 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc $?"; cut -c1-900 gpurun_out/r2_bench4.json; grep -E "trace|to_bedgraph:|destroy" gpurun_out/r2_bench4.err | tail -8
timeout 600 python scripts/bam_writer_bench.py 10000000 > gpurun_out/r2_bam_writers_10M_b.json 2> gpurun_out/r2_bam_writers_10M_b.err; echo "writer bench rc $?"; cat gpurun_out/r2_bam_writers_10M_b.json; grep write_bam gpurun_out/r2_bam_writers_10M_b.err | tail -4
