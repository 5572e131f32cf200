#!/bin/bash
# round-2 final evidence run: GPU suite, default bench line, full-size split, launch list and one full capture of the shipped K1
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests8.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2_gpu_tests8.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke3.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r2_smoke3.log
BND_DEBUG_TIMING=1 BND_BENCH_TRACE=1 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench rc $?"; cut -c1-1000 gpurun_out/r2_bench6.json; grep -E "trace|to_bedgraph:" gpurun_out/r2_bench6.err | tail -4
timeout 600 python scripts/bam_writer_full.py > gpurun_out/r2_bam_split_full.json 2> gpurun_out/r2_bam_split_full.err; echo "full split rc $?"; cat gpurun_out/r2_bam_split_full.json; grep write_bam gpurun_out/r2_bam_split_full.err | tail -2
CMD="python bench.py --pairs 10000000 --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
$CMD > gpurun_out/r2_plain10M_c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_10Mpairs_final.csv $CMD > gpurun_out/r2_ncu_launches2.log 2>&1
echo "launch list rc $?"
$CMD > gpurun_out/r2_plain10M_d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bgzf_inflate_spec -s 30 -c 2 -o gpurun_out/r2_k1_bench_final $CMD > gpurun_out/r2_ncu_full2.log 2>&1
echo "full capture rc $?"
