// bigwig_tools.cuh — `bigwig-compare` and `bigwig-aggregate` on the device (bamnado/src/bigwig_compare.rs).
//
// The reference loads every chromosome's intervals with bigtools, sweeps them on one core per chromosome and integrates the
// result over fixed-width bins.  Here the sections of every input are inflated by the zlib variant of K1 (inflate_spec.cuh),
// turned into one item array per file, and every output bin is computed by its own thread, which replays exactly the additions
// the reference's sweep makes for that bin, in the same order and in f64 without contraction - so sums, means and ratios are
// bit-identical; only LogRatio goes through the device's log() (within 1e-6 relative of libm's).
//
//   bwr_headers / bwr_emit_items   section headers -> host (item counts, chromosome ranges); sections -> {start, end, value}
//                                  for bedGraph (type 1), variableStep (2) and fixedStep (3) sections
//   bwt_bins                       one thread per bin: compare (subtraction, ratio, log-ratio), sum / mean of per-file bin means,
//                                  max / min of the swept signals
//   bwt_median_count / _fill       weighted lower median: (value, bases) pairs per bin counted, scanned, collected, heap-sorted
//   bwt_heads / bwt_emit           consecutive equal bins of a chromosome merge into one item (push_bigwig_values_for_bins)
// All of it is integer / float streaming work bound by HBM latency (binary searches) and bandwidth; no tensor-core shape.
#pragma once
#include "bigwig.cuh"

namespace bnd {

__global__ void __launch_bounds__(256) bwr_headers_kernel(const uint8_t *data, const uint64_t *sec_off, uint32_t n_secs, uint32_t *headers) {
  for (uint32_t s = blockIdx.x * 256u + threadIdx.x; s < n_secs; s += gridDim.x * 256u) {
    const uint32_t *h = reinterpret_cast<const uint32_t *>(data + sec_off[s]);  // sections start on 16-byte lines of the inflate buffer
    for (int k = 0; k < 6; k++) headers[(size_t)s * 6 + k] = h[k];
  }
}

// one warp per section
__global__ void __launch_bounds__(256) bwr_emit_items_kernel(BwrEmitArgs a) {
  const int lane = (int)(threadIdx.x & 31u);
  for (uint32_t s = (uint32_t)((blockIdx.x * 256ull + threadIdx.x) >> 5); s < a.n_secs; s += (uint32_t)((gridDim.x * 256ull) >> 5)) {
    const uint8_t *sec = a.data + a.sec_off[s];
    const uint32_t *h = reinterpret_cast<const uint32_t *>(sec);
    const uint32_t chrom_start = h[1], step = h[3], span = h[4], type = h[5] & 0xffu, n = h[5] >> 16;
    const uint32_t *body = h + 6;
    BwItem *out = a.items + a.item_off[s];
    unsigned bad = 0;
    for (uint32_t i = (uint32_t)lane; i < n; i += 32) {
      BwItem it;
      if (type == 1) {
        it.start = body[3 * i], it.end = body[3 * i + 1], it.val = __uint_as_float(body[3 * i + 2]);
      } else if (type == 2) {
        it.start = body[2 * i], it.end = it.start + span, it.val = __uint_as_float(body[2 * i + 1]);
      } else {
        it.start = chrom_start + i * step, it.end = it.start + span, it.val = __uint_as_float(body[i]);
      }
      if (it.end <= it.start) bad = 1;
      out[i] = it;
    }
    if (__ballot_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.bad, 1u);
  }
}

// items of one chromosome must be sorted and disjoint for the per-bin binary searches (bigtools' own reader only sorts defensively)
__global__ void __launch_bounds__(256) bwr_check_sorted_kernel(const BwItem *items, const uint64_t *range, uint32_t n_chrom, unsigned int *bad) {
  for (uint32_t c = blockIdx.y; c < n_chrom; c += gridDim.y) {
    const uint64_t i0 = range[2 * c], i1 = range[2 * c + 1];
    for (uint64_t i = i0 + 1 + blockIdx.x * 256ull + threadIdx.x; i < i1; i += (uint64_t)gridDim.x * 256ull)
      if (items[i].start < items[i - 1].end) atomicOr(bad, 2u);
  }
}

// ---- per-bin replay of the reference's sweeps ------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t bwt_first_end_after(const BwItem *it, uint64_t lo, uint64_t hi, uint32_t pos) {
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (it[mid].end <= pos) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// StepSignalCursor::fetch (bigwig_compare.rs:160-177): value at pos and the next position where it can change
__device__ __forceinline__ void bwt_fetch(const BwItem *it, uint64_t &idx, uint64_t hi, uint32_t pos, uint32_t chrom_len, float &v, uint32_t &next) {
  while (idx < hi && it[idx].end <= pos) idx++;
  if (idx < hi) {
    const BwItem iv = it[idx];
    if (iv.start <= pos && pos < iv.end) {
      v = iv.val;
      next = iv.end < chrom_len ? iv.end : chrom_len;
    } else {
      v = 0.f;
      next = iv.start < chrom_len ? iv.start : chrom_len;
    }
  } else {
    v = 0.f;
    next = chrom_len;
  }
}

__device__ __forceinline__ uint32_t bwt_chrom_of_bin(const BwtArgs &a, uint64_t b) {
  uint32_t lo = 0, hi = a.n_chrom;  // last chromosome whose first bin is <= b
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a.bin_first[mid] <= b) lo = mid;
    else hi = mid;
  }
  return lo;
}

// BinAccumulator over one file (binned_means_single_bw, :190-206): intervals in order, zero values skipped, f64 sums
__device__ __forceinline__ float bwt_single_mean(const BwItem *it, uint64_t i0, uint64_t i1, uint32_t bs, uint32_t be, uint32_t chrom_len) {
  double sum = 0.0;
  for (uint64_t i = bwt_first_end_after(it, i0, i1, bs); i < i1; i++) {
    const BwItem iv = it[i];
    if (iv.start >= be) break;
    if (iv.val == 0.f) continue;
    const uint32_t s = iv.start > bs ? iv.start : bs, e0 = iv.end < chrom_len ? iv.end : chrom_len, e = e0 < be ? e0 : be;
    if (e > s) sum = __dadd_rn(sum, __dmul_rn((double)iv.val, (double)(e - s)));
  }
  return (float)__ddiv_rn(sum, (double)(be - bs));
}

__global__ void __launch_bounds__(256) bwt_bins_kernel(BwtArgs a) {
  for (uint64_t b = blockIdx.x * 256ull + threadIdx.x; b < a.n_bins; b += (uint64_t)gridDim.x * 256ull) {
    const uint32_t c = bwt_chrom_of_bin(a, b);
    const uint32_t chrom_len = a.chrom_len[c];
    const uint32_t bs = (uint32_t)(b - a.bin_first[c]) * a.bin_size;
    const uint32_t be = bs + a.bin_size < chrom_len && bs + a.bin_size > bs ? bs + a.bin_size : chrom_len;
    float out;
    if (a.mode == BWT_SUM || a.mode == BWT_MEAN) {
      // binned_sum_or_mean (:676-698): every file's bin mean (f32), scaled and shifted in f64, added in file order
      double acc = 0.0;
      for (int f = 0; f < a.n_files; f++) {
        const float m = bwt_single_mean(a.items[f], a.range[f][2 * c], a.range[f][2 * c + 1], bs, be, chrom_len);
        acc = __dadd_rn(acc, __dadd_rn(__dmul_rn((double)m, a.scale[f]), a.pseudocount));
      }
      out = (float)__ddiv_rn(acc, a.mode == BWT_MEAN ? (double)a.n_files : 1.0);
    } else {
      // binned_compare_two_bw (:226-268) / binned_extrema_multi_bw (:276-328): the sweep restricted to this bin
      uint64_t idx[BWT_MAX_FILES];
      for (int f = 0; f < a.n_files; f++) idx[f] = bwt_first_end_after(a.items[f], a.range[f][2 * c], a.range[f][2 * c + 1], bs);
      double sum = 0.0;
      uint32_t pos = bs;
      while (pos < be) {
        uint32_t next = chrom_len;
        double val;
        if (a.mode <= BWT_LOGRATIO) {
          float v1, v2;
          uint32_t n1, n2;
          bwt_fetch(a.items[0], idx[0], a.range[0][2 * c + 1], pos, chrom_len, v1, n1);
          bwt_fetch(a.items[1], idx[1], a.range[1][2 * c + 1], pos, chrom_len, v2, n2);
          next = n1 < n2 ? n1 : n2;
          const double s1 = __dmul_rn((double)v1, a.scale[0]), s2 = __dmul_rn((double)v2, a.scale[1]), pc = a.pseudocount;
          if (a.mode == BWT_SUBTRACTION) val = __dsub_rn(s1, s2);
          else if (a.mode == BWT_RATIO) val = __ddiv_rn(s1, __dadd_rn(s2, pc));
          else val = log(__ddiv_rn(__dadd_rn(s1, pc), __dadd_rn(s2, pc)));
        } else {
          const bool want_max = a.mode == BWT_MAX;
          val = want_max ? -INFINITY : INFINITY;
          for (int f = 0; f < a.n_files; f++) {
            float v;
            uint32_t n;
            bwt_fetch(a.items[f], idx[f], a.range[f][2 * c + 1], pos, chrom_len, v, n);
            next = n < next ? n : next;
            const double vv = __dadd_rn(__dmul_rn((double)v, a.scale[f]), a.pseudocount);
            val = want_max ? fmax(val, vv) : fmin(val, vv);
          }
        }
        if (next <= pos) next = pos + 1 < chrom_len ? pos + 1 : chrom_len;
        const uint32_t seg_end = next < be ? next : be;
        sum = __dadd_rn(sum, __dmul_rn(val, (double)(seg_end - pos)));
        pos = seg_end;
      }
      out = (float)__ddiv_rn(sum, (double)(be - bs));
    }
    a.bins[b] = out;
  }
}

// ---- weighted median (binned_median_multi_bw, :334-470) ----------------------------------------------------------------------
// collect_weighted_values_for_bw_in_bin: the bin's bases of one file as (value, bases) runs, gaps carrying `gap_value`
template <bool STORE>
__device__ __forceinline__ uint32_t bwt_collect(const BwItem *it, uint64_t i0, uint64_t i1, uint32_t bs, uint32_t be, double scale, double pc, float gap_value,
                                                float *vals, uint32_t *w) {
  uint32_t n = 0, pos = bs;
  uint64_t i = bwt_first_end_after(it, i0, i1, bs);
  while (pos < be) {
    if (i >= i1) {
      if (STORE) vals[n] = gap_value, w[n] = be - pos;
      n++;
      break;
    }
    const BwItem iv = it[i];
    if (iv.start > pos) {
      const uint32_t gap_end = iv.start < be ? iv.start : be;
      if (STORE) vals[n] = gap_value, w[n] = gap_end - pos;
      n++;
      pos = gap_end;
      continue;
    }
    const uint32_t seg_end = iv.end < be ? iv.end : be;  // iv.start <= pos < iv.end (sorted, disjoint items; first end after pos)
    if (STORE) vals[n] = (float)__dadd_rn(__dmul_rn((double)iv.val, scale), pc), w[n] = seg_end - pos;
    n++;
    pos = seg_end;
    if (pos >= iv.end) i++;
  }
  return n;
}

__global__ void __launch_bounds__(256) bwt_median_count_kernel(BwtArgs a) {
  for (uint64_t b = blockIdx.x * 256ull + threadIdx.x; b < a.n_bins; b += (uint64_t)gridDim.x * 256ull) {
    const uint32_t c = bwt_chrom_of_bin(a, b);
    const uint32_t chrom_len = a.chrom_len[c];
    const uint32_t bs = (uint32_t)(b - a.bin_first[c]) * a.bin_size;
    const uint32_t be = bs + a.bin_size < chrom_len && bs + a.bin_size > bs ? bs + a.bin_size : chrom_len;
    uint64_t n = 0;
    for (int f = 0; f < a.n_files; f++)
      n += bwt_collect<false>(a.items[f], a.range[f][2 * c], a.range[f][2 * c + 1], bs, be, a.scale[f], a.pseudocount, 0.f, nullptr, nullptr);
    a.pair_off[b] = n;
  }
}

__device__ __forceinline__ void bwt_sift(float *v, uint32_t *w, uint64_t root, uint64_t n) {
  for (;;) {
    uint64_t child = 2 * root + 1;
    if (child >= n) return;
    if (child + 1 < n && v[child] < v[child + 1]) child++;
    if (!(v[root] < v[child])) return;
    const float tv = v[root];
    const uint32_t tw = w[root];
    v[root] = v[child], w[root] = w[child];
    v[child] = tv, w[child] = tw;
    root = child;
  }
}

__global__ void __launch_bounds__(256) bwt_median_fill_kernel(BwtArgs a) {
  const float gap_value = (float)a.pseudocount;
  for (uint64_t b = blockIdx.x * 256ull + threadIdx.x; b < a.n_bins; b += (uint64_t)gridDim.x * 256ull) {
    const uint32_t c = bwt_chrom_of_bin(a, b);
    const uint32_t chrom_len = a.chrom_len[c];
    const uint32_t bs = (uint32_t)(b - a.bin_first[c]) * a.bin_size;
    const uint32_t be = bs + a.bin_size < chrom_len && bs + a.bin_size > bs ? bs + a.bin_size : chrom_len;
    float *v = a.pair_val + a.pair_off[b];
    uint32_t *w = a.pair_w + a.pair_off[b];
    uint64_t n = 0;
    for (int f = 0; f < a.n_files; f++)
      n += bwt_collect<true>(a.items[f], a.range[f][2 * c], a.range[f][2 * c + 1], bs, be, a.scale[f], a.pseudocount, gap_value, v + n, w + n);
    // heap sort by value (the reference's stable sort and this one agree on which VALUE carries the median weight)
    for (uint64_t k = n / 2; k-- > 0;) bwt_sift(v, w, k, n);
    for (uint64_t m = n; m > 1; m--) {
      const float tv = v[0];
      const uint32_t tw = w[0];
      v[0] = v[m - 1], w[0] = w[m - 1];
      v[m - 1] = tv, w[m - 1] = tw;
      bwt_sift(v, w, 0, m - 1);
    }
    uint64_t total = 0;
    for (uint64_t k = 0; k < n; k++) total += w[k];
    float med = 0.f;
    if (total) {
      const uint64_t half = (total + 1) / 2;  // lower median
      uint64_t cum = 0;
      for (uint64_t k = 0; k < n; k++) {
        cum += w[k];
        if (cum >= half) {
          med = v[k];
          break;
        }
      }
    }
    a.bins[b] = med;
  }
}

// ---- bins -> items (push_bigwig_values_for_bins, :520-566) -------------------------------------------------------------------
__device__ __forceinline__ bool bwt_is_head(const BwtArgs &a, uint64_t b) {
  const uint32_t c = bwt_chrom_of_bin(a, b);
  return b == a.bin_first[c] || a.bins[b] != a.bins[b - 1];
}

__global__ void __launch_bounds__(COL_THREADS) bwt_heads_kernel(BwtArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t b = (uint64_t)t * COL_THREADS + threadIdx.x;
    uint64_t h = b < a.n_bins && bwt_is_head(a, b) ? 1 : 0, total;
    block_exclusive_scan<COL_THREADS, uint64_t>(h, scratch, total);
    if (threadIdx.x == 0) a.tile_count[t] = total;
  }
}

__global__ void __launch_bounds__(COL_THREADS) bwt_emit_kernel(BwtArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t b = (uint64_t)t * COL_THREADS + threadIdx.x;
    const bool in = b < a.n_bins;
    uint64_t h = in && bwt_is_head(a, b) ? 1 : 0, total;
    const uint64_t j = a.tile_count[t] + block_exclusive_scan<COL_THREADS, uint64_t>(h, scratch, total);
    if (h) {
      const uint32_t c = bwt_chrom_of_bin(a, b);
      const uint64_t cb_end = a.bin_first[c + 1];
      uint64_t e = b + 1;  // the run ends where the next head of this chromosome is (short forward scan: runs are what the output stores)
      while (e < cb_end && a.bins[e] == a.bins[b]) e++;
      BwItem it;
      it.start = (uint32_t)(b - a.bin_first[c]) * a.bin_size;
      const uint64_t end64 = (e - a.bin_first[c]) * (uint64_t)a.bin_size;
      it.end = end64 < a.chrom_len[c] ? (uint32_t)end64 : a.chrom_len[c];
      it.val = a.bins[b];
      a.out_items[j] = it;
      if (b == a.bin_first[c]) a.chrom_item_first[c] = j;
    }
  }
}

}  // namespace bnd
