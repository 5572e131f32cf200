// collapse.cuh — K6: `collapse-bedgraph` run merging and the multi-BAM equal-bin collapse.
//
// collapse_adjacent_bins (bamnado/src/bedgraph_utils.rs:5-45): rows sorted by (chrom, start); a new group starts at the
// first row, when the chromosome changes, when start != previous end, or when |score - previous score| > 1e-6 (against
// the previous ROW, so drift chains); a group reports min start, max end and its first score.
// collapse_equal_bins (bamnado/src/coverage_analysis.rs:81-134): a new group starts when the chromosome or ANY sample's
// score differs from the previous row (exact equality, no contiguity test).
// Both are head-flag + scan + stream compaction: HBM streaming, no tensor-core shape.
#pragma once
#include "common.cuh"
#include "finalize.cuh"

namespace bnd {

constexpr int COL_THREADS = 256;

// ---- collapse-bedgraph -------------------------------------------------------------------------------------------
__device__ __forceinline__ bool bdg_is_head(const BdgArgs &a, uint64_t i) {
  if (i == 0) return true;
  if (a.chrom[i] != a.chrom[i - 1]) return true;
  if (a.start[i] != a.end[i - 1]) return true;
  double d = a.score[i] - a.score[i - 1];
  return fabs(d) > 1e-6;
}

// rows must be sorted by (chrom, start); flags out-of-order neighbours in a.unsorted
__global__ void __launch_bounds__(COL_THREADS) bdg_sorted_check_kernel(BdgArgs a) {
  int bad = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)COL_THREADS + threadIdx.x + 1; i < a.n; i += (uint64_t)gridDim.x * COL_THREADS)
    if (a.chrom[i] < a.chrom[i - 1] || (a.chrom[i] == a.chrom[i - 1] && a.start[i] < a.start[i - 1])) bad = 1;
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(a.unsorted, 1u);
}

__global__ void __launch_bounds__(COL_THREADS) bdg_count_kernel(BdgArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    uint64_t i = (uint64_t)t * COL_THREADS + threadIdx.x;
    uint64_t h = i < a.n && bdg_is_head(a, i) ? 1 : 0, total;
    block_exclusive_scan<COL_THREADS, uint64_t>(h, scratch, total);
    if (threadIdx.x == 0) a.tile_heads[t] = total;
  }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) u64_scan_kernel(uint64_t *v, uint32_t n, unsigned long long *total_out) {
  __shared__ uint64_t scratch[THREADS / 32 + 1];
  __shared__ uint64_t running;
  const int tid = (int)threadIdx.x;
  if (tid == 0) running = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += THREADS) {
    uint32_t k = base + (uint32_t)tid;
    uint64_t x = k < n ? v[k] : 0ull, total;
    uint64_t ex = block_exclusive_scan<THREADS, uint64_t>(x, scratch, total);
    if (k < n) v[k] = running + ex;
    __syncthreads();
    if (tid == 0) running += total;
    __syncthreads();
  }
  if (tid == 0) *total_out = running;
}

__global__ void __launch_bounds__(COL_THREADS) bdg_emit_kernel(BdgArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    uint64_t i = (uint64_t)t * COL_THREADS + threadIdx.x;
    bool in = i < a.n;
    uint64_t h = in && bdg_is_head(a, i) ? 1 : 0, total;
    uint64_t ex = block_exclusive_scan<COL_THREADS, uint64_t>(h, scratch, total);
    if (in) {
      uint64_t gid = a.tile_heads[t] + ex + h - 1;  // group of row i (heads open a new group)
      if (h) {
        a.out_chrom[gid] = a.chrom[i];
        a.out_score[gid] = a.score[i];
      }
      atomicMin(&a.out_start[gid], (long long)a.start[i]);
      atomicMax(&a.out_end[gid], (long long)a.end[i]);
    }
  }
}

// ---- multi-BAM equal-bin collapse --------------------------------------------------------------------------------
__device__ __forceinline__ bool multi_is_head(const MultiArgs &a, uint64_t i, bool ref_start) {
  if (a.flags) return a.flags[i] != 0;
  if (ref_start) return true;
  for (int b = 0; b < a.n_bams; b++)
    if (a.counts[b][i] != a.counts[b][i - 1]) return true;
  return false;
}

__device__ __forceinline__ bool bin_is_ref_start(const GeomDev &g, uint64_t gbin) {
  int32_t r = ref_of_bin(g, gbin);
  return __ldg(g.ref_first_bin + r) == gbin;
}

// per-bin head flags of the BAMs held by this rank; ranks OR them together (all-reduce MAX) to get the shared run heads
__global__ void __launch_bounds__(COL_THREADS) multi_flags_kernel(MultiArgs a) {
  for (uint64_t i = blockIdx.x * (uint64_t)COL_THREADS + threadIdx.x; i < a.n_bins; i += (uint64_t)gridDim.x * COL_THREADS) {
    bool h = i == 0 || bin_is_ref_start(a.g, i + a.g.bin_lo);
    for (int b = 0; b < a.n_bams && !h; b++) h = a.counts[b][i] != a.counts[b][i - 1];
    a.flags_out[i] = h ? 1 : 0;
  }
}

__global__ void __launch_bounds__(COL_THREADS) multi_count_kernel(MultiArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    uint64_t i = (uint64_t)t * COL_THREADS + threadIdx.x;
    uint64_t h = i < a.n_bins && multi_is_head(a, i, i == 0 || bin_is_ref_start(a.g, i + a.g.bin_lo)) ? 1 : 0, total;
    block_exclusive_scan<COL_THREADS, uint64_t>(h, scratch, total);
    if (threadIdx.x == 0) a.tile_heads[t] = total;
  }
}

__global__ void __launch_bounds__(COL_THREADS) multi_emit_kernel(MultiArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    uint64_t i = (uint64_t)t * COL_THREADS + threadIdx.x;
    bool in = i < a.n_bins;
    uint64_t h = in && multi_is_head(a, i, i == 0 || bin_is_ref_start(a.g, i + a.g.bin_lo)) ? 1 : 0, total;
    uint64_t ex = block_exclusive_scan<COL_THREADS, uint64_t>(h, scratch, total);
    if (h) {
      uint64_t j = a.tile_heads[t] + ex;
      a.run_bin[j] = i + a.g.bin_lo;
      for (int b = 0; b < a.n_bams; b++) a.out_vals[(uint64_t)b * a.n_runs_cap + j] = a.counts[b][i];
    }
  }
}

// rows of the multi-BAM table: [start of the head bin, stop of the last bin of the group)
__global__ void __launch_bounds__(256) multi_rows_kernel(RunArgs a) {
  for (uint64_t j = blockIdx.x * 256ull + threadIdx.x; j < a.n_runs; j += (uint64_t)gridDim.x * 256ull) {
    uint64_t gb = a.run_bin[j];
    uint64_t nx = j + 1 < a.n_runs ? a.run_bin[j + 1] : a.g.bin_hi;
    RunRow r;
    uint64_t stop_unused, start_unused;
    run_coords(a.g, gb, gb + 1, r.ref_id, r.start, stop_unused);
    int32_t ref2;
    run_coords(a.g, nx - 1, nx, ref2, start_unused, r.stop);
    r.val = 0;
    a.rows[j] = r;
  }
}

}  // namespace bnd
