"""multi-bam-coverage with one BAM per rank (BASELINE config 5's multi-GPU shape).

torchrun --nproc-per-node N scripts/multi_bam_dist.py --bams a.bam b.bam ... --bin-size 1 -o out.tsv [--backend nccl|gloo]

Every rank piles up its own BAM (raw counts, no collapse), the per-bin change flags are OR-ed over ranks with one
all-reduce(MAX) (NCCL over NVLink on GPU tensors; the emulator/gloo build uses CPU tensors), every rank compacts its
column at the shared run heads and rank 0 gathers the columns and writes the TSV of MultiBamPileup::to_tsv
(bamnado/src/coverage_analysis.rs:855-864).  `--emul` loads the test-only emulator build (tests/emul) instead of the GPU
library; it exists for the CPU test-suite.
"""
import argparse, ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from bamnado_b200 import _native

ap = argparse.ArgumentParser()
ap.add_argument("--bams", nargs="+", required=True)
ap.add_argument("--bin-size", type=int, default=50)
ap.add_argument("-o", "--output", required=True)
ap.add_argument("--backend", default="nccl")
ap.add_argument("--emul", action="store_true")
ap.add_argument("--min-mapq", type=int, default=20)
args = ap.parse_args()

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
use_cuda = args.backend == "nccl"
if use_cuda:
    torch.cuda.set_device(local)
dist.init_process_group(args.backend, **({"device_id": torch.device("cuda", local)} if use_cuda else {}))
if len(args.bams) != world:
    raise SystemExit("one BAM per rank: pass exactly WORLD_SIZE BAMs")
if args.emul:
    from bamnado_b200.build import build_emul
    N = _native.Native(ctypes.CDLL(str(build_emul())))
else:
    N = _native.lib()
dev = torch.device("cuda", local) if use_cuda else torch.device("cpu")
t0 = time.perf_counter()
pile = dict(bam_path=args.bams[rank], bin_size=args.bin_size, collapse=False, device=local if use_cuda else -1)
filt = dict(min_mapq=args.min_mapq, min_length=20, max_length=1000)
with N.pileup(pile, filt) as h:
    n_bins = h.multi_prepare()
    nb = torch.tensor([n_bins], dtype=torch.int64, device=dev)
    lo, hi = nb.clone(), nb.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if int(lo) != int(hi):
        raise SystemExit("the BAMs do not share one bin grid (different reference dictionaries or chunk lengths)")
    flags = torch.empty(max(n_bins, 1), dtype=torch.uint8, device=dev)
    h.multi_change_flags(flags.data_ptr())
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)  # OR of the change masks: the only exchange proportional to the bins
    if use_cuda:
        torch.cuda.synchronize()
    rows, vals = h.multi_compact(flags.data_ptr())
    ref_names, _ = h.refs()
vals_t = torch.from_numpy(vals.astype(np.int64)).to(dev)
gathered = [torch.empty_like(vals_t) for _ in range(world)] if rank == 0 else None
dist.gather(vals_t, gathered, dst=0)
if rank == 0:
    cols = [g.cpu().numpy() for g in gathered]
    order = np.lexsort((rows["start"], np.array([ref_names[r] for r in rows["ref_id"]], dtype=object)))
    with open(args.output, "w") as f:
        f.write("chrom\tstart\tend\t" + "\t".join(args.bams) + "\n")
        for k in order:
            f.write(f"{ref_names[rows['ref_id'][k]]}\t{rows['start'][k]}\t{rows['stop'][k]}\t" + "\t".join(str(int(c[k])) for c in cols) + "\n")
    print(f"multi-bam-coverage x{world}: {len(rows)} rows, {n_bins} bins, {time.perf_counter() - t0:.3f} s", flush=True)
dist.destroy_process_group()
