#!/bin/bash
# round-2 run G (2 GPUs): sharded bam-coverage after the round's host changes; page-table entries of the BAM mapping dropped per
# piece (the old behaviour) against kept
set -u
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline"
BND_DROP_PTES=1 BND_BENCH_TRACE=1 $RUN > gpurun_out/r2_cfg2_n2_drop.json 2> gpurun_out/r2_cfg2_n2_drop.err; echo "rc $?"; cut -c1-200 gpurun_out/r2_cfg2_n2_drop.json; python - <<'PY'
import json
for f in ("drop",):
    d=json.loads(open(f"gpurun_out/r2_cfg2_n2_{f}.json").read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6)
PY
grep "trace rank" gpurun_out/r2_cfg2_n2_drop.err | tail -2
BND_BENCH_TRACE=1 $RUN > gpurun_out/r2_cfg2_n2_keep.json 2> gpurun_out/r2_cfg2_n2_keep.err; echo "rc $?"; python - <<'PY'
import json
for f in ("keep",):
    d=json.loads(open(f"gpurun_out/r2_cfg2_n2_{f}.json").read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6, d.get("parity"))
PY
grep "trace rank" gpurun_out/r2_cfg2_n2_keep.err | tail -2
