#!/bin/bash
# round-2 evidence run: GPU tests, the default bench line, the ncu launch list of a bench command and one full capture of K1
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests4.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2_gpu_tests4.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r2_smoke.log
BND_DEBUG_TIMING=1 BND_BENCH_TRACE=1 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc $?"; cat gpurun_out/r2_bench3.json
CMD="python bench.py --pairs 10000000 --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
$CMD > gpurun_out/r2_plain10M.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_10Mpairs.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc $?"
$CMD > gpurun_out/r2_plain10M_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bgzf_inflate_spec -s 30 -c 2 -o gpurun_out/r2_k1_bench $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo "full capture rc $?"
ls -la gpurun_out/*.ncu-rep | tail -3
