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
__device__ __forceinline__ bool bin_is_ref_start(const GeomDev &g, uint64_t gbin) {
  int32_t r = ref_of_bin(g, gbin);
  return __ldg(g.ref_first_bin + r) == gbin;
}

__device__ __forceinline__ bool multi_local_head(const MultiArgs &a, uint64_t i) {
  if (i == 0 || bin_is_ref_start(a.g, i + a.g.bin_lo)) return true;
  for (int b = 0; b < a.n_bams; b++)
    if (a.counts[b][i] != a.counts[b][i - 1]) return true;
  return false;
}

__device__ __forceinline__ bool multi_is_head(const MultiArgs &a, uint64_t i) {
  if (a.bits) return (a.bits[i >> 5] >> (i & 31u)) & 1u;
  return multi_local_head(a, i);
}

// Head flags of the BAMs held by this rank, one bit per bin (a warp ballot packs 32 bins into a word).  With one BAM per
// rank the masks of all ranks are OR-ed (multi_or_kernel over the gathered masks) to get the shared run heads: 1 bit per
// bin on the wire instead of a byte (hg38 at bin 1: 387 MB instead of 3.1 GB).
__global__ void __launch_bounds__(COL_THREADS) multi_bits_kernel(MultiArgs a) {
  const uint64_t n_words = (a.n_bins + 31) / 32;
  const uint64_t warp = (blockIdx.x * (uint64_t)COL_THREADS + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * COL_THREADS) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint64_t w = warp; w < n_words; w += n_warps) {
    const uint64_t i = w * 32 + lane;
    const bool h = i < a.n_bins && multi_local_head(a, i);
    const unsigned m = __ballot_sync(0xffffffffu, h);
    if (lane == 0) a.bits_out[w] = m;
  }
}

// dst[w] = OR over the n_src masks laid out back to back (src[k * n_words + w]); dst may alias mask 0
__global__ void __launch_bounds__(COL_THREADS) multi_or_kernel(uint32_t *dst, const uint32_t *src, uint32_t n_src, uint64_t n_words) {
  for (uint64_t w = blockIdx.x * (uint64_t)COL_THREADS + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * COL_THREADS) {
    uint32_t v = 0;
    for (uint32_t k = 0; k < n_src; k++) v |= src[(uint64_t)k * n_words + w];
    dst[w] = v;
  }
}

__global__ void __launch_bounds__(COL_THREADS) multi_count_kernel(MultiArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    uint64_t i = (uint64_t)t * COL_THREADS + threadIdx.x;
    uint64_t h = i < a.n_bins && multi_is_head(a, i) ? 1 : 0, total;
    block_exclusive_scan<COL_THREADS, uint64_t>(h, scratch, total);
    if (threadIdx.x == 0) a.tile_heads[t] = total;
  }
}

__global__ void __launch_bounds__(COL_THREADS) multi_emit_kernel(MultiArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    uint64_t i = (uint64_t)t * COL_THREADS + threadIdx.x;
    bool in = i < a.n_bins;
    uint64_t h = in && multi_is_head(a, i) ? 1 : 0, total;
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

// ---- TSV text of the multi-BAM table on the device ---------------------------------------------------------------------
__device__ __forceinline__ void multi_row_coords(const GeomDev &g, uint64_t gb, uint64_t nx, int32_t &ref, uint64_t &start, uint64_t &stop) {
  uint64_t unused;
  int32_t ref2;
  run_coords(g, gb, gb + 1, ref, start, unused);
  run_coords(g, nx - 1, nx, ref2, unused, stop);
}

__global__ void __launch_bounds__(TEXT_THREADS) multi_text_len_kernel(MultiTextArgs a) {
  __shared__ uint64_t s_part[TEXT_THREADS / 32];
  const int tid = (int)threadIdx.x;
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t k = (uint64_t)t * TEXT_THREADS + (uint64_t)tid;  // row inside the slice
    uint32_t len = 0;
    if (k < a.n_rows) {
      const uint64_t j = a.row_lo + k;
      const uint64_t gb = a.run_bin[j], nx = j + 1 < a.n_runs ? a.run_bin[j + 1] : a.g.bin_hi;
      int32_t ref;
      uint64_t start, stop;
      multi_row_coords(a.g, gb, nx, ref, start, stop);
      len = (a.name_off[ref + 1] - a.name_off[ref]) + dec_len(start) + dec_len(stop) + 3;
      for (int c = 0; c < a.n_cols; c++) len += 1 + dec_len(a.vals[(uint64_t)a.col_pos[c] * a.stride + k]);
      a.row_len[k] = (uint16_t)len;
    }
    uint64_t s = warp_sum((uint64_t)len);
    if ((tid & 31) == 0) s_part[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
      uint64_t S = 0;
      for (int w = 0; w < TEXT_THREADS / 32; w++) S += s_part[w];
      a.tile_off[t] = S;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(TEXT_THREADS) multi_text_write_kernel(MultiTextArgs a) {
  __shared__ uint64_t scratch[TEXT_THREADS / 32 + 1];
  const int tid = (int)threadIdx.x;
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t k = (uint64_t)t * TEXT_THREADS + (uint64_t)tid;
    uint64_t len = k < a.n_rows ? a.row_len[k] : 0, total;
    const uint64_t off = a.tile_off[t] + block_exclusive_scan<TEXT_THREADS, uint64_t>(len, scratch, total);
    if (k < a.n_rows) {
      const uint64_t j = a.row_lo + k;
      const uint64_t gb = a.run_bin[j], nx = j + 1 < a.n_runs ? a.run_bin[j + 1] : a.g.bin_hi;
      int32_t ref;
      uint64_t start, stop;
      multi_row_coords(a.g, gb, nx, ref, start, stop);
      // first row of a reference inside this slice: where that reference's text begins (a reference's first bin is always a head)
      if (k == 0 || gb == tmax<uint64_t>(__ldg(a.g.ref_first_bin + ref), a.g.bin_lo)) a.ref_text_off[ref] = off;
      if (!a.text) continue;  // sizes only: the caller wants the bytes per reference, not the text
      char *p = a.text + off;
      for (uint32_t q = a.name_off[ref]; q < a.name_off[ref + 1]; q++) *p++ = a.names[q];
      *p++ = '\t';
      p = put_dec(p, start);
      *p++ = '\t';
      p = put_dec(p, stop);
      for (int c = 0; c < a.n_cols; c++) {
        *p++ = '\t';
        p = put_dec(p, a.vals[(uint64_t)a.col_pos[c] * a.stride + k]);
      }
      *p++ = '\n';
    }
  }
}

}  // namespace bnd
