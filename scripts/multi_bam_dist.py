"""multi-bam-coverage with one BAM per rank (BASELINE config 5's multi-GPU shape): thin launcher of bamnado_b200.multi_dist.

torchrun --nproc-per-node N scripts/multi_bam_dist.py --bams a.bam b.bam ... --bin-size 1 -o out.tsv [--backend nccl|gloo]

`--emul` loads the test-only emulator build (tests/emul) instead of the GPU library; it exists for the CPU test-suite.
"""
import argparse, ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bamnado_b200 import _native, multi_dist

ap = argparse.ArgumentParser()
ap.add_argument("--bams", nargs="+", required=True)
ap.add_argument("--bin-size", type=int, default=50)
ap.add_argument("-o", "--output", required=True)
ap.add_argument("--backend", default="nccl")
ap.add_argument("--emul", action="store_true")
ap.add_argument("--min-mapq", type=int, default=20)
ap.add_argument("--slice-rows", type=int, default=64 << 20, help="rows formatted per pass (bounds the TSV text held in HBM)")
args = ap.parse_args()

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
use_cuda = args.backend == "nccl"
if use_cuda:
    torch.cuda.set_device(local)
dist.init_process_group(args.backend, **({"device_id": torch.device("cuda", local)} if use_cuda else {}))
if args.emul:
    from bamnado_b200.build import build_emul
    N = _native.Native(ctypes.CDLL(str(build_emul())))
else:
    N = _native.lib()
dev = torch.device("cuda", local) if use_cuda else torch.device("cpu")
t0 = time.perf_counter()
if rank == 0 and os.path.exists(args.output):
    os.unlink(args.output)
dist.barrier()
res = multi_dist.multi_bam_to_tsv(N, args.bams, dict(bin_size=args.bin_size), dict(min_mapq=args.min_mapq, min_length=20, max_length=1000),
                                  args.output, rank=rank, world=world, device=dev, slice_rows=args.slice_rows)
if rank == 0:
    print(f"multi-bam-coverage x{world}: {res['rows']} rows, {res['bins']} bins, {res['nvlink_bytes']} B exchanged, {time.perf_counter() - t0:.3f} s", flush=True)
dist.destroy_process_group()
