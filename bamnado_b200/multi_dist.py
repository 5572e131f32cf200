"""multi-bam-coverage with one BAM per rank (MultiBamPileup::to_tsv, bamnado/src/coverage_analysis.rs:783-864).

Every rank piles up its own BAM (raw counts, no collapse) on its own GPU.  What crosses NVLink:

  masks    one BIT per bin "my count changes here" (hg38 at bin 1: 387 MB per rank), all-gathered and OR-ed on the device
  columns  every rank compacts its column at the shared row heads and receives rows [k n / world, (k + 1) n / world) of every
           column (all-to-all, 4 B per row and BAM)
  tables   n_refs x u64 text bytes per reference, all-gathered, so that every rank knows where its pieces go

Every rank then formats its slice of rows on the device and fills its pieces of the ONE TSV file.  torch.distributed is the
plumbing only (NCCL on CUDA tensors; gloo on CPU tensors for the emulator build used by the CPU test-suite).
"""
from __future__ import annotations

import time

import torch
import torch.distributed as dist


def multi_bam_to_tsv(N, bams: list, pile: dict, filt: dict, out_path, *, rank: int, world: int, device: torch.device, timings: dict | None = None) -> dict:
    """Runs steps 1-7 of include/bamnado_b200.h's one-BAM-per-rank protocol.  `pile` holds bin_size etc. (collapse is forced
    off like MultiBamPileup requires); returns {"rows", "bins", "nvlink_bytes", "tsv_bytes"} (job-wide)."""
    if len(bams) != world:
        raise ValueError("one BAM per rank: pass exactly WORLD_SIZE BAMs")
    cuda = device.type == "cuda"

    def sync():
        if cuda:
            torch.cuda.synchronize(device)

    t0 = time.perf_counter()
    cfg = dict(pile, bam_path=str(bams[rank]), collapse=False, device=device.index if cuda else -1, shard_rank=0, shard_world=1)
    with N.pileup(cfg, filt) as h:
        n_bins = h.multi_prepare()
        t1 = time.perf_counter()
        nb = torch.tensor([n_bins, -n_bins], dtype=torch.int64, device=device)
        dist.all_reduce(nb, op=dist.ReduceOp.MAX)
        if int(nb[0]) != n_bins or int(nb[1]) != -n_bins:
            raise RuntimeError("the BAMs do not share one bin grid (different reference dictionaries or chunk lengths)")
        n_words = (n_bins + 31) // 32
        masks = torch.empty((world, max(n_words, 1)), dtype=torch.int32, device=device)
        mine = masks[rank]
        h.multi_change_bits(mine.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(masks.view(-1), mine.clone())
            sync()
            N.multi_or_bits(device.index if cuda else -1, masks.data_ptr(), masks.data_ptr(), world, max(n_words, 1))
        n_rows, vals_ptr = h.multi_compact_column(masks.data_ptr())
        del masks
        # column slices: rank k formats rows [bounds[k], bounds[k + 1])
        bounds = [n_rows * k // world for k in range(world + 1)]
        stride = max(max(bounds[k + 1] - bounds[k] for k in range(world)), 1)
        col = _wrap_u32(vals_ptr, n_rows, device)
        send = torch.zeros((world, stride), dtype=torch.int32, device=device)
        for k in range(world):
            send[k, : bounds[k + 1] - bounds[k]] = col[bounds[k]:bounds[k + 1]]
        recv = torch.empty_like(send)
        if world > 1:
            dist.all_to_all_single(recv.view(-1), send.view(-1))
        else:
            recv.copy_(send)
        sync()
        t2 = time.perf_counter()
        sizes = h.multi_format_tsv(recv.data_ptr(), stride, world, bounds[rank], bounds[rank + 1])
        t = torch.tensor(sizes, dtype=torch.int64, device=device)
        allv = [torch.empty_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(allv, t)
        else:
            allv = [t]
        tables = [[int(x) for x in v.tolist()] for v in allv]
        header = "chrom\tstart\tend\t" + "\t".join(str(b) for b in bams) + "\n"
        h.multi_write_tsv_pieces(out_path, header, tables, rank)
        if world > 1:
            dist.barrier()
        t3 = time.perf_counter()
    if timings is not None:
        timings.update(pileup_s=t1 - t0, exchange_s=t2 - t1, text_s=t3 - t2)
    mask_bytes = 4 * n_words * (world - 1) * world            # all-gather: every rank receives world - 1 masks
    col_bytes = 4 * n_rows * (world - 1)                       # all-to-all: every column leaves its rank except its own slice
    return {"rows": n_rows, "bins": n_bins, "nvlink_bytes": mask_bytes + col_bytes if world > 1 else 0,
            "tsv_bytes": len(header.encode()) + sum(sum(t) for t in tables)}


def _wrap_u32(ptr: int, n: int, device: torch.device) -> torch.Tensor:
    """A torch view (int32) of n u32 values the library owns at `ptr` (device memory, or host memory with the emulator build)."""
    if n == 0:
        return torch.zeros(0, dtype=torch.int32, device=device)
    if device.type == "cuda":
        class _Arr:  # __cuda_array_interface__: zero-copy view of library-owned device memory
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, True), "version": 3}
        return torch.as_tensor(_Arr(), device=device)
    import ctypes

    import numpy as np

    buf = (ctypes.c_int32 * n).from_address(ptr)
    return torch.from_numpy(np.ctypeslib.as_array(buf))
