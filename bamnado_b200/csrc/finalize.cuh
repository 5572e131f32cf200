// finalize.cuh — K5: prefix scan of the bin difference array, run (collapse) detection and stream compaction, plus
// the consumers of the compacted runs (row expansion, dense per-base signal, bedGraph text).
//
// Replaces bin_intervals' per-bin loop with its in-chunk collapse (bamnado/src/bam_utils.rs:57-90), the
// score * factor multiply of pileup_normalised (bamnado/src/coverage_analysis.rs:601-611), the row layout of
// pileup_to_polars (:550-598), the dense fill of get_signal_for_chromosome (bamnado-python/src/lib.rs:294-300) and
// the text rows of write_bedgraph (coverage_analysis.rs:43-75).
//
// The diff array is a flat int32 array over all bins of this shard; K4 placed compensation entries at chunk
// starts, so ONE unsegmented inclusive prefix sum yields every bin's count.  Run heads are bins whose diff is
// non-zero (the count changes there) or that start a chunk (runs never cross chunks); without collapse every
// bin is a head.  The head set does not depend on the scan carry, so pass A counts heads and sums diffs per tile,
// a single-CTA pass scans the tile aggregates, and pass B redoes the local scan and writes (bin, count) per head.
// Traffic: 2 x 4 B read per bin + 12 B written per run.  HBM-bound streaming, no tensor-core shape.
#pragma once
#include "common.cuh"
#include "pileup.cuh"

namespace bnd {


// Position of a global bin inside the chunk geometry.
struct BinPos {
  int32_t ref;
  uint64_t local;  // bin index inside the reference
  uint64_t rem;    // bin index inside its chunk
  uint64_t next_ref_first;  // global index of the first bin of the next reference that has bins
};

__device__ __forceinline__ int32_t ref_of_bin(const GeomDev &g, uint64_t gbin) {
  // last r with ref_first_bin[r] <= gbin (references without bins share their successor's first bin)
  int32_t lo = 0, hi = g.n_ref;  // search in [0, n_ref)
  while (hi - lo > 1) {
    int32_t mid = (lo + hi) >> 1;
    if (__ldg(g.ref_first_bin + mid) <= gbin) lo = mid;
    else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void bin_locate(const GeomDev &g, uint64_t gbin, BinPos &p) {
  p.ref = ref_of_bin(g, gbin);
  p.local = gbin - __ldg(g.ref_first_bin + p.ref);
  p.rem = g.nbf ? p.local % g.nbf : 0;
  p.next_ref_first = __ldg(g.ref_first_bin + p.ref + 1);
}

// Per-thread walk over FIN_IPT consecutive bins computing the head flags as a bit mask.
__device__ __forceinline__ uint32_t head_mask(const FinArgs &a, uint64_t first, uint32_t n, const int32_t *d) {
  if (n == 0) return 0;
  if (!a.collapse) return n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
  const GeomDev &g = a.g;
  BinPos p;
  bin_locate(g, first + g.bin_lo, p);
  uint32_t m = 0;
  uint64_t gbin = first + g.bin_lo;
#pragma unroll 1
  for (uint32_t i = 0; i < n; i++, gbin++) {
    if (gbin == p.next_ref_first) {  // entering the next reference that has bins
      bin_locate(g, gbin, p);
    }
    if (p.rem == 0 || d[i] != 0) m |= 1u << i;
    if (++p.rem == g.nbf) p.rem = 0;
  }
  return m;
}

__device__ __forceinline__ void load_bins(const FinArgs &a, uint64_t first, uint32_t n, int32_t *d) {
  if (n == FIN_IPT) {
    const int4 *src = reinterpret_cast<const int4 *>(a.diff + first);
#pragma unroll
    for (int q = 0; q < FIN_IPT / 4; q++) {
      int4 v = src[q];
      d[4 * q] = v.x, d[4 * q + 1] = v.y, d[4 * q + 2] = v.z, d[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < FIN_IPT; i++) d[i] = (uint32_t)i < n ? a.diff[first + i] : 0;
  }
}

__global__ void __launch_bounds__(FIN_THREADS) bins_reduce_kernel(FinArgs a) {
  __shared__ int32_t s_sum[FIN_THREADS / 32];
  __shared__ uint32_t s_heads[FIN_THREADS / 32];
  const int tid = (int)threadIdx.x;
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    uint64_t first = (uint64_t)t * FIN_TILE + (uint64_t)tid * FIN_IPT;
    uint32_t n = first < a.n_bins ? (uint32_t)tmin<uint64_t>(FIN_IPT, a.n_bins - first) : 0;
    int32_t d[FIN_IPT];
    load_bins(a, first, n, d);
    int32_t sum = 0;
#pragma unroll
    for (int i = 0; i < FIN_IPT; i++) sum += d[i];
    uint32_t heads = (uint32_t)__popc(head_mask(a, first, n, d));
    sum = warp_sum(sum);
    heads = warp_sum(heads);
    if ((tid & 31) == 0) {
      s_sum[tid >> 5] = sum;
      s_heads[tid >> 5] = heads;
    }
    __syncthreads();
    if (tid == 0) {
      int32_t S = 0;
      uint32_t H = 0;
      for (int w = 0; w < FIN_THREADS / 32; w++) S += s_sum[w], H += s_heads[w];
      a.tile_sum[t] = S;
      a.tile_heads[t] = H;
    }
    __syncthreads();
  }
}

// Single CTA: exclusive scans of tile_sum (wrapping int32) and tile_heads (u64), in place.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) tile_scan_kernel(FinArgs a) {
  __shared__ uint64_t scratch[THREADS / 32 + 1];
  __shared__ uint64_t run_sum, run_heads;
  const int tid = (int)threadIdx.x;
  if (tid == 0) run_sum = 0, run_heads = 0;
  __syncthreads();
  for (uint32_t base = 0; base < a.n_tiles; base += THREADS) {
    uint32_t k = base + (uint32_t)tid;
    uint64_t s = k < a.n_tiles ? (uint64_t)(uint32_t)a.tile_sum[k] : 0ull;
    uint64_t h = k < a.n_tiles ? a.tile_heads[k] : 0ull;
    uint64_t ts, th;
    uint64_t es = block_exclusive_scan<THREADS, uint64_t>(s, scratch, ts);
    uint64_t eh = block_exclusive_scan<THREADS, uint64_t>(h, scratch, th);
    if (k < a.n_tiles) {
      a.tile_sum[k] = (int32_t)(uint32_t)(run_sum + es);
      a.tile_heads[k] = run_heads + eh;
    }
    __syncthreads();
    if (tid == 0) run_sum += ts, run_heads += th;
    __syncthreads();
  }
  if (tid == 0) a.fs->n_runs = run_heads;
}

__global__ void __launch_bounds__(FIN_THREADS) bins_emit_kernel(FinArgs a) {
  __shared__ uint64_t scratch[FIN_THREADS / 32 + 1];
  __shared__ unsigned int s_max;
  const int tid = (int)threadIdx.x;
  if (tid == 0) s_max = 0;
  __syncthreads();
  unsigned int my_max = 0;
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    uint64_t first = (uint64_t)t * FIN_TILE + (uint64_t)tid * FIN_IPT;
    uint32_t n = first < a.n_bins ? (uint32_t)tmin<uint64_t>(FIN_IPT, a.n_bins - first) : 0;
    int32_t d[FIN_IPT];
    load_bins(a, first, n, d);
    uint32_t hm = head_mask(a, first, n, d);
    int32_t sum = 0;
#pragma unroll
    for (int i = 0; i < FIN_IPT; i++) sum += d[i];
    // two CTA scans: the diff sums (only the low 32 bits matter: counts wrap like the reference's `as u32`) and
    // the head counts (output positions)
    uint64_t ts, th;
    uint64_t es = block_exclusive_scan<FIN_THREADS, uint64_t>((uint64_t)(uint32_t)sum, scratch, ts);
    uint64_t eh = block_exclusive_scan<FIN_THREADS, uint64_t>((uint64_t)__popc(hm), scratch, th);
    uint32_t run = (uint32_t)a.tile_sum[t] + (uint32_t)es;  // count just before this thread's first bin
    uint64_t pos = a.tile_heads[t] + eh;
#pragma unroll
    for (int i = 0; i < FIN_IPT; i++) {
      if ((uint32_t)i < n) {
        run += (uint32_t)d[i];
        if (a.counts) a.counts[first + i] = run;
        if ((hm >> i) & 1u) {
          if (a.run_bin) {
            a.run_bin[pos] = first + (uint64_t)i + a.g.bin_lo;
            a.run_val[pos] = run;
          }
          pos++;
          my_max = tmax(my_max, run);
        }
      }
    }
  }
  my_max = tmax(my_max, (unsigned int)__shfl_xor_sync(0xffffffffu, my_max, 16));
  my_max = tmax(my_max, (unsigned int)__shfl_xor_sync(0xffffffffu, my_max, 8));
  my_max = tmax(my_max, (unsigned int)__shfl_xor_sync(0xffffffffu, my_max, 4));
  my_max = tmax(my_max, (unsigned int)__shfl_xor_sync(0xffffffffu, my_max, 2));
  my_max = tmax(my_max, (unsigned int)__shfl_xor_sync(0xffffffffu, my_max, 1));
  if ((tid & 31) == 0) atomicMax(&s_max, my_max);
  __syncthreads();
  if (tid == 0 && s_max) atomicMax(&a.fs->max_val, s_max);
}

// ---- consumers of the compacted runs ---------------------------------------------------------------------------

// Genomic [start, stop) of run j (1-based half-open as the reference keeps them).
__device__ __forceinline__ void run_coords(const GeomDev &g, uint64_t gbin, uint64_t next_gbin, int32_t &ref, uint64_t &start, uint64_t &stop) {
  ref = ref_of_bin(g, gbin);
  const uint64_t fb = __ldg(g.ref_first_bin + ref), nfb = __ldg(g.ref_first_bin + ref + 1);
  const uint64_t local = gbin - fb;
  const uint64_t c = local / g.nbf, j = local - c * g.nbf;
  const uint64_t cs = c * g.L + 1, ce = tmin<uint64_t>(cs + g.L - 1, __ldg(g.ref_len + ref));
  start = cs + j * g.bin;
  const uint64_t chunk_end_bin = tmin<uint64_t>(fb + (c + 1) * g.nbf, nfb);
  stop = next_gbin < chunk_end_bin ? cs + (next_gbin - fb - c * g.nbf) * g.bin : ce;
}


__global__ void __launch_bounds__(256) expand_runs_kernel(RunArgs a) {
  for (uint64_t j = blockIdx.x * 256ull + threadIdx.x; j < a.n_runs; j += (uint64_t)gridDim.x * 256ull) {
    uint64_t gb = a.run_bin[j];
    uint64_t nx = j + 1 < a.n_runs ? a.run_bin[j + 1] : a.g.bin_hi;
    RunRow r;
    run_coords(a.g, gb, nx, r.ref_id, r.start, r.stop);
    r.val = a.run_val[j];
    a.rows[j] = r;
  }
}

// get_signal_for_chromosome: out[i] = count(bin holding 1-based coordinate i) as f32 * scale, for the runs'
// [start, stop) clipped to [0, chrom_size); index 0 and every chunk-terminal base stay 0.

__global__ void __launch_bounds__(256) dense_signal_kernel(SignalArgs a) {
  const GeomDev &g = a.g;
  const uint64_t fb = __ldg(g.ref_first_bin + a.ref) - g.bin_lo;
  for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < a.chrom_size; i += (uint64_t)gridDim.x * 256ull) {
    float v = 0.f;
    if (i >= 1 && g.nbf > 0) {
      uint64_t c = (i - 1) / g.L;
      uint64_t cs = c * g.L + 1, ce = tmin<uint64_t>(cs + g.L - 1, a.chrom_size);
      if (i < ce) v = (float)a.counts[fb + c * g.nbf + (i - cs) / g.bin] * a.scale;
    }
    a.out[i] = v;
  }
}

// ---- bedGraph text on the device --------------------------------------------------------------------------------
// Row text = name \t start \t stop \t score \n.  Scores come from a host-built table indexed by the raw count
// (count * factor formatted once per distinct count; the number of distinct counts is tiny next to the rows).

__device__ __forceinline__ uint32_t dec_len(uint64_t v) {
  uint32_t n = 1;
  while (v >= 10) v /= 10, n++;
  return n;
}
__device__ __forceinline__ char *put_dec(char *p, uint64_t v) {
  char tmp[20];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

__global__ void __launch_bounds__(TEXT_THREADS) text_len_kernel(TextArgs a) {
  __shared__ uint64_t s_part[TEXT_THREADS / 32];
  const int tid = (int)threadIdx.x;
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    uint64_t j = (uint64_t)t * TEXT_THREADS + (uint64_t)tid;
    uint32_t len = 0;
    if (j < a.n_runs) {
      uint64_t gb = a.run_bin[j];
      uint64_t nx = j + 1 < a.n_runs ? a.run_bin[j + 1] : a.g.bin_hi;
      int32_t ref;
      uint64_t start, stop;
      run_coords(a.g, gb, nx, ref, start, stop);
      if (!a.ref_skip[ref]) {
        uint32_t v = a.run_val[j];
        uint32_t sl = v < a.n_scores ? (uint32_t)(uint8_t)a.score_text[(size_t)v * SCORE_W] : 0u;
        len = (a.name_off[ref + 1] - a.name_off[ref]) + dec_len(start) + dec_len(stop) + sl + 4;
      }
      a.row_len[j] = (uint16_t)len;
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

template <int THREADS>
__global__ void __launch_bounds__(THREADS) text_scan_kernel(TextArgs a) {
  __shared__ uint64_t scratch[THREADS / 32 + 1];
  __shared__ uint64_t running;
  const int tid = (int)threadIdx.x;
  if (tid == 0) running = 0;
  __syncthreads();
  for (uint32_t base = 0; base < a.n_tiles; base += THREADS) {
    uint32_t k = base + (uint32_t)tid;
    uint64_t v = k < a.n_tiles ? a.tile_off[k] : 0ull, total;
    uint64_t ex = block_exclusive_scan<THREADS, uint64_t>(v, scratch, total);
    if (k < a.n_tiles) a.tile_off[k] = running + ex;
    __syncthreads();
    if (tid == 0) running += total;
    __syncthreads();
  }
  if (tid == 0) a.fs->text_bytes = running;
}

// A tile's rows (256 x ~43 bytes) are formatted into shared memory, at the same 16-byte phase as their place in the output, and
// leave as one bulk store of the copy engine (plus the ragged ends): the per-thread byte stores of the first version reached 150 GB/s (14 ms for the 2.1 GB of the
// BASELINE run, profiles/r2_launches_bench_10Mpairs_summary.txt).  Tiles longer than the staging buffer take the direct path.
constexpr uint32_t TEXT_STAGE = 24 * 1024;
__device__ __forceinline__ void text_format_row(const TextArgs &a, char *p, int32_t ref, uint64_t start, uint64_t stop, uint32_t val) {
  const uint32_t n0 = a.name_off[ref], n1 = a.name_off[ref + 1];
  for (uint32_t k = n0; k < n1; k++) *p++ = a.names[k];
  *p++ = '\t';
  p = put_dec(p, start);
  *p++ = '\t';
  p = put_dec(p, stop);
  *p++ = '\t';
  const char *sc = a.score_text + (size_t)val * SCORE_W;
  const uint32_t sl = (uint8_t)sc[0];
  for (uint32_t k = 0; k < sl; k++) *p++ = sc[1 + k];
  *p++ = '\n';
}

__global__ void __launch_bounds__(TEXT_THREADS) text_write_kernel(TextArgs a) {
  __shared__ uint64_t scratch[TEXT_THREADS / 32 + 1];
  __shared__ __align__(16) char stage[TEXT_STAGE + 16];
  const int tid = (int)threadIdx.x;
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    uint64_t j = (uint64_t)t * TEXT_THREADS + (uint64_t)tid;
    uint64_t len = j < a.n_runs ? a.row_len[j] : 0, total;
    const uint64_t tile_off = a.tile_off[t];
    const uint64_t off = tile_off + block_exclusive_scan<TEXT_THREADS, uint64_t>(len, scratch, total);
    const uint32_t phase = (uint32_t)((uintptr_t)(a.text + tile_off) & 15u);
    const bool staged = total <= TEXT_STAGE;
    if (j < a.n_runs) {
      uint64_t gb = a.run_bin[j];
      uint64_t nx = j + 1 < a.n_runs ? a.run_bin[j + 1] : a.g.bin_hi;
      int32_t ref;
      uint64_t start, stop;
      run_coords(a.g, gb, nx, ref, start, stop);
      // first row of a reference records where that reference's text begins
      if (gb == tmax<uint64_t>(__ldg(a.g.ref_first_bin + ref), a.g.bin_lo)) a.ref_text_off[ref] = off;  // first row of the reference in this shard
      if (len) text_format_row(a, staged ? stage + phase + (off - tile_off) : a.text + off, ref, start, stop, a.run_val[j]);
    }
    if (staged) bulk_store_fence();  // the rows written above become visible to the copy engine
    __syncthreads();
    if (staged && total) {
      char *dst = a.text + tile_off - phase;  // 16-byte aligned; stage[k] belongs at dst[k] for k in [phase, phase + total)
      const uint32_t end = phase + (uint32_t)total;
      const uint32_t v0 = phase ? 16u : 0u, v1 = end & ~15u;  // whole 16-byte lines inside the tile
      // the aligned middle of the tile (~11 KB) leaves as ONE bulk store of the copy engine, the ragged ends by hand
      if (tid == 0 && v1 > v0) bulk_store_s2g(dst + v0, smem_addr(stage + v0), v1 - v0);
      for (uint32_t k = phase + (uint32_t)tid; k < tmin<uint32_t>(v0, end); k += TEXT_THREADS) dst[k] = stage[k];
      for (uint32_t k = tmax<uint32_t>(v1, v0) + (uint32_t)tid; k < end; k += TEXT_THREADS) dst[k] = stage[k];
    }
    __syncthreads();
  }
}

// One thread per bin of the pile-up's own layout (per reference: chunk-major, `nbf` bins per full chunk).  The bin that
// starts at base `start` lands in grid row (start - 1) / bin of its chromosome (bin_counts.rs:258-260); bins the chunks do
// not produce (the single-base bin after a chunk's last base) keep the matrix's zero.
__global__ void bin_matrix_kernel(BinMatrixArgs a) {
  const uint64_t n = a.g.bin_hi - a.g.bin_lo;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t gb = a.g.bin_lo + t;
    int lo = 0, hi = a.g.n_ref;  // last reference whose first bin is <= gb
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (a.g.ref_first_bin[mid] <= gb)
        lo = mid;
      else
        hi = mid;
    }
    const uint64_t row0 = a.grid_row0[lo];
    if (row0 == ~0ull || a.g.nbf == 0) continue;
    const uint64_t j = gb - a.g.ref_first_bin[lo];
    const uint64_t row = row0 + (j / a.g.nbf) * a.bins_per_chunk_grid + (j % a.g.nbf);
    a.matrix[row * a.n_samples + a.sample] = a.counts[t];
  }
}

}  // namespace bnd
