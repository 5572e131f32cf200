"""multi-bam-coverage sharded by BAM across ranks (MultiBamPileup::to_tsv, bamnado/src/coverage_analysis.rs:783-864).

Rank r piles up BAMs [r k, (r + 1) k) of the list (k = n_bams / world; one BAM per GPU at world = n_bams) on its own GPU: raw
counts, no collapse.  What crosses NVLink:

  masks    one BIT per bin "one of my BAMs changes here" (hg38 at bin 1: 387 MB per rank), all-gathered and OR-ed on the device
  columns  every rank compacts its columns at the shared row heads and receives rows [j n / world, (j + 1) n / world) of every
           column (one all-to-all per local column, 4 B per row and BAM, straight out of the compacted column: no send copy)
  tables   n_refs x u64 text bytes per reference, all-gathered, so that every rank knows where its pieces go

Every rank then formats its slice of rows on the device and fills its pieces of the ONE TSV file.  torch.distributed is the
plumbing only (NCCL on CUDA tensors; gloo on CPU tensors for the emulator build used by the CPU test-suite).
"""
from __future__ import annotations

import time

import torch
import torch.distributed as dist


class MultiJob:
    """Handles of this rank's BAMs, kept open so that a bench can stage the compressed bytes once (resident timing)."""

    def __init__(self, N, bams: list, pile: dict, filt: dict, *, rank: int, world: int, device: torch.device, slice_rows: int = 64 << 20):
        if len(bams) % world:
            raise ValueError("the number of BAMs must be a multiple of the number of ranks")
        self.N, self.bams, self.rank, self.world, self.device = N, [str(b) for b in bams], rank, world, device
        self.k = len(bams) // world
        self.slice_rows = max(1, slice_rows)
        self.cuda = device.type == "cuda"
        mine = self.bams[rank * self.k:(rank + 1) * self.k]
        self.handles = [N.pileup(dict(pile, bam_path=b, collapse=False, device=device.index if self.cuda else -1, shard_rank=0, shard_world=1), filt)
                        for b in mine]

    def close(self):
        for h in self.handles:
            h.close()

    def stage(self):
        for h in self.handles:
            h.stage()

    def _sync(self):
        if self.cuda:
            torch.cuda.synchronize(self.device)

    def table(self, resident: bool = False) -> dict:
        """Steps 1-5: pile-ups, masks, shared heads, column slices.  Leaves this rank's slice of all columns in self.recv."""
        N, world, rank, k, dev = self.N, self.world, self.rank, self.k, self.device
        t0 = time.perf_counter()
        n_bins = 0
        for h in self.handles:
            n_bins = h.multi_prepare(resident)
        t1 = time.perf_counter()
        if world > 1:
            nb = torch.tensor([n_bins, -n_bins], dtype=torch.int64, device=dev)
            dist.all_reduce(nb, op=dist.ReduceOp.MAX)
            if int(nb[0]) != n_bins or int(nb[1]) != -n_bins:
                raise RuntimeError("the BAMs do not share one bin grid (different reference dictionaries or chunk lengths)")
        n_words = max((n_bins + 31) // 32, 1)
        masks = torch.empty((world, n_words), dtype=torch.int32, device=dev)
        N.multi_change_bits(self.handles, masks[rank].data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(masks.view(-1), masks[rank].clone())
            self._sync()
            N.multi_or_bits(dev.index if self.cuda else -1, masks.data_ptr(), masks.data_ptr(), world, n_words)
        n_rows, vals_ptr = N.multi_compact_columns(self.handles, masks.data_ptr())
        del masks
        bounds = [n_rows * j // world for j in range(world + 1)]
        lo, hi = bounds[rank], bounds[rank + 1]
        n_mine = hi - lo
        if world == 1:
            self.cols_ptr, self.stride, self.col_order, self.recv = vals_ptr, max(n_rows, 1), None, None
        else:
            cols = _wrap_u32(vals_ptr, k * n_rows, dev).view(k, n_rows) if n_rows else torch.zeros((k, 0), dtype=torch.int32, device=dev)
            # recv[c][i] = rows [lo, hi) of local column c of rank i = BAM i k + c
            self.recv = torch.empty((k, world, max(n_mine, 1)), dtype=torch.int32, device=dev)
            in_split = [bounds[j + 1] - bounds[j] for j in range(world)]
            for c in range(k):
                out = self.recv[c].view(-1)[: world * n_mine] if n_mine else self.recv[c].view(-1)[:0]
                dist.all_to_all_single(out, cols[c], output_split_sizes=[n_mine] * world, input_split_sizes=in_split)
            self._sync()
            self.cols_ptr, self.stride = self.recv.data_ptr(), max(n_mine, 1)
            self.col_order = [(g % k) * world + g // k for g in range(world * k)]
        self.lo, self.hi, self.n_rows, self.n_bins = lo, hi, n_rows, n_bins
        t2 = time.perf_counter()
        mask_bytes = 4 * n_words * (world - 1) * world  # all-gather: every rank receives world - 1 masks
        col_bytes = 4 * n_rows * k * (world - 1)        # all-to-all: every column leaves its rank except its own slice
        return {"rows": n_rows, "bins": n_bins, "nvlink_bytes": mask_bytes + col_bytes if world > 1 else 0, "pileup_s": t1 - t0, "exchange_s": t2 - t1}

    def _slices(self) -> list:
        """This rank's rows cut into sub-slices of at most `slice_rows` rows (bounds the text held in HBM); the same count on every rank."""
        world, n = self.world, self.n_rows
        biggest = max(n * (j + 1) // world - n * j // world for j in range(world))
        n_sub = max(1, -(-biggest // self.slice_rows))
        mine = self.hi - self.lo
        return [(self.lo + mine * s // n_sub, self.lo + mine * (s + 1) // n_sub) for s in range(n_sub)]

    def _format(self, lo: int, hi: int, sizes_only: bool) -> list:
        # dev_cols holds this rank's rows from self.lo on: the sub-slice starts (lo - self.lo) rows into every column
        return self.handles[0].multi_format_tsv(self.cols_ptr + 4 * (lo - self.lo), self.stride, self.world * self.k, lo, hi, self.col_order, sizes_only)

    def format(self) -> list:
        """Step 6 without keeping text (resident timing): TSV text of every sub-slice is produced on the device."""
        return [self._format(lo, hi, False) for lo, hi in self._slices()]

    def write(self, out_path) -> int:
        """Steps 6 + 7: byte tables of all sub-slices (virtual ranks), all-gather, then format + fill piece by piece.  Returns the TSV size."""
        world, dev = self.world, self.device
        subs = self._slices()
        mine = [self._format(lo, hi, True) for lo, hi in subs]
        if world > 1:
            t = torch.tensor(mine, dtype=torch.int64, device=dev).view(-1)
            allv = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            tables = [row for v in allv for row in v.view(len(subs), -1).tolist()]
        else:
            tables = mine
        header = "chrom\tstart\tend\t" + "\t".join(self.bams) + "\n"
        for s, (lo, hi) in enumerate(subs):
            self._format(lo, hi, False)
            self.handles[0].multi_write_tsv_pieces(out_path, header, tables, self.rank * len(subs) + s)
        self.recv = None
        if world > 1:
            dist.barrier()
        return len(header.encode()) + sum(sum(t) for t in tables)


def multi_bam_to_tsv(N, bams: list, pile: dict, filt: dict, out_path, *, rank: int, world: int, device: torch.device, slice_rows: int = 64 << 20) -> dict:
    """The whole job: {"rows", "bins", "nvlink_bytes", "tsv_bytes", "pileup_s", "exchange_s", "text_s"} (job-wide)."""
    job = MultiJob(N, bams, pile, filt, rank=rank, world=world, device=device, slice_rows=slice_rows)
    try:
        res = job.table()
        t0 = time.perf_counter()
        res["tsv_bytes"] = job.write(out_path)
        res["text_s"] = time.perf_counter() - t0
    finally:
        job.close()
    return res


def _wrap_u32(ptr: int, n: int, device: torch.device) -> torch.Tensor:
    """A torch view (int32) of n u32 values the library owns at `ptr` (device memory, or host memory with the emulator build)."""
    if n == 0:
        return torch.zeros(0, dtype=torch.int32, device=device)
    if device.type == "cuda":
        class _Arr:  # __cuda_array_interface__: zero-copy view of library-owned device memory
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}

        return torch.as_tensor(_Arr(), device=device)
    import ctypes

    import numpy as np

    buf = (ctypes.c_int32 * n).from_address(ptr)
    return torch.from_numpy(np.ctypeslib.as_array(buf))
