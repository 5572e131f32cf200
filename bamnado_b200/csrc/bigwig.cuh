// bigwig.cuh — BigWig encoding on the device: items, data sections, zlib (DEFLATE) compression, zoom summaries.
//
// Replaces the bigtools writer the reference drives on a tokio pool (bamnado/src/coverage_analysis.rs:642-759): rows in BAM
// header order, both ends shifted by -1, score = (count as f64 * factor) as f32, 1 024 items per section, zlib-compressed
// sections, zoom levels.  The host (outputs.cpp) only lays the file out: header, chromosome B+ tree, R-tree indexes built from
// the per-section bounds computed here.
//
//   bw_row_first   first run of every reference (binary search in the compacted runs)
//   bw_items       runs -> {start - 1, stop - 1, f32 score} of the references that are kept
//   bw_deflate     one CTA per section: the section's bytes are assembled in shared memory and written as ONE fixed-Huffman
//                  DEFLATE block inside a zlib stream.  Sections are arrays of fixed-size records, so matches are searched at
//                  a few fixed lags only (same field of the previous record, previous field of this record) and never cross a
//                  record: every record is parsed independently by its own thread - no serial walk, no hash table - a block
//                  scan of the per-thread bit counts gives every thread its bit offset, and the bits are OR-ed into a shared
//                  memory image of the stream.  Adler-32, section bounds and the summary statistics come out of the same pass.
//   bw_compact     packs the sections back to back (offsets from a scan of their sizes)
//   bw_zoom_*      one thread per zoom window.  First level: binary search of the first item overlapping the window, sequential
//                  f32 accumulation in item order (about ten items per window); every further level folds the four finer windows
//                  of the level below (deterministic order); then stream compaction of the windows that hold data
// All of it is byte/integer streaming work: HBM- and shared-memory-bound, no tensor-core shape.
#pragma once
#include "common.cuh"
#include "finalize.cuh"

namespace bnd {

__global__ void bw_row_first_kernel(const uint64_t *run_bin, uint64_t n_runs, GeomDev g, uint64_t *row_first) {
  const int r = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (r > g.n_ref) return;
  const uint64_t key = g.ref_first_bin[r];  // [n_ref + 1]
  uint64_t lo = 0, hi = n_runs;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (run_bin[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  row_first[r] = lo;
}

__global__ void __launch_bounds__(256) bw_items_kernel(BwItemArgs a) {
  for (uint64_t j = blockIdx.x * 256ull + threadIdx.x; j < a.n_runs; j += (uint64_t)gridDim.x * 256ull) {
    const uint64_t gb = a.run_bin[j];
    const uint64_t nx = j + 1 < a.n_runs ? a.run_bin[j + 1] : a.g.bin_hi;
    int32_t ref;
    uint64_t start, stop;
    run_coords(a.g, gb, nx, ref, start, stop);
    if (a.ref_skip[ref]) continue;
    if (start == 0 || stop - 1 > 0xffffffffull) {
      atomicOr(a.range_error, 1u);
      continue;
    }
    BwItem it;
    it.start = (uint32_t)(start - 1);
    it.end = (uint32_t)(stop - 1);
    it.val = (float)((double)a.run_val[j] * a.factor);
    a.items[j - a.row_first[ref] + a.item_first[ref]] = it;
  }
}

// ---- fixed-Huffman DEFLATE (RFC 1951 3.2.6) --------------------------------------------------------------------------
// code of a literal/length symbol, bit-reversed for the LSB-first stream: value | bits << 16
__device__ __forceinline__ uint32_t fx_litlen(uint32_t sym) {
  uint32_t code, n;
  if (sym < 144) code = 0x30 + sym, n = 8;
  else if (sym < 256) code = 0x190 + (sym - 144), n = 9;
  else if (sym < 280) code = sym - 256, n = 7;
  else code = 0xC0 + (sym - 280), n = 8;
  return (__brev(code) >> (32 - n)) | (n << 16);
}
// length 3..258 -> code + extra bits, ready to be OR-ed into the stream: value | bits << 16
__device__ __forceinline__ uint32_t fx_length(uint32_t len) {
  uint32_t sym, eb = 0, extra = 0;
  if (len <= 10) {
    sym = 254 + len;
  } else if (len == 258) {
    sym = 285;
  } else {
    const uint32_t x = len - 3;
    const uint32_t msb = 31u - (uint32_t)__clz((int)x);
    eb = msb - 2;
    sym = 257 + 4 * eb + 4 + ((x >> eb) & 3u);
    extra = x & ((1u << eb) - 1u);
  }
  const uint32_t c = fx_litlen(sym);
  return ((c & 0xffffu) | (extra << (c >> 16))) | (((c >> 16) + eb) << 16);
}
// distance 1..32768 -> 5-bit code (reversed) + extra bits: value | bits << 16
__device__ __forceinline__ uint32_t fx_distance(uint32_t d) {
  uint32_t code, eb = 0, extra = 0;
  if (d <= 4) {
    code = d - 1;
  } else {
    const uint32_t x = d - 1;
    const uint32_t msb = 31u - (uint32_t)__clz((int)x);
    eb = msb - 1;
    code = 2 * msb + ((x >> eb) & 1u);
    extra = x & ((1u << eb) - 1u);
  }
  return ((__brev(code) >> 27) | (extra << 5)) | ((5 + eb) << 16);
}

__device__ __forceinline__ void bw_put(uint32_t *out, uint32_t bitpos, uint32_t value, uint32_t nbits) {
  const uint32_t w = bitpos >> 5, sh = bitpos & 31u;
  atomicOr(&out[w], value << sh);
  if (sh + nbits > 32) atomicOr(&out[w + 1], value >> (32 - sh));
}

// Greedy parse of raw[u0, u1): at every byte the longest match among the fixed lags (>= 3 bytes, inside the unit, source may
// lie in earlier units) or a literal.  Returns the bits used; EMIT writes them at `bitpos`.
template <bool EMIT>
__device__ __forceinline__ uint32_t bw_parse_unit(const uint8_t *raw, uint32_t u0, uint32_t u1, const uint32_t *lags, uint32_t *out, uint32_t bitpos) {
  uint32_t bits = 0, p = u0;
#pragma unroll 1
  while (p < u1) {
    uint32_t best = 0, best_d = 0;
    const uint32_t room = u1 - p;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t d = lags[k];
      if (d == 0 || d > p) continue;
      uint32_t l = 0;
      while (l < room && raw[p + l] == raw[p + l - d]) l++;
      if (l > best) best = l, best_d = d;
    }
    if (best >= 3) {
      const uint32_t lc = fx_length(best), dc = fx_distance(best_d);
      if (EMIT) {
        bw_put(out, bitpos + bits, lc & 0xffffu, lc >> 16);
        bw_put(out, bitpos + bits + (lc >> 16), dc & 0xffffu, dc >> 16);
      }
      bits += (lc >> 16) + (dc >> 16);
      p += best;
    } else {
      const uint32_t c = fx_litlen(raw[p]);
      if (EMIT) bw_put(out, bitpos + bits, c & 0xffffu, c >> 16);
      bits += c >> 16;
      p++;
    }
  }
  return bits;
}

// One CTA per section.  Dynamic shared memory: raw bytes (header + count * unit, padded to 4) | stream image (words).
__global__ void __launch_bounds__(BW_THREADS) bw_deflate_kernel(BwDeflateArgs a) {
  BND_DYN_SMEM(smem);
  __shared__ uint64_t scratch[BW_THREADS / 32 + 1];
  __shared__ double s_red[5][BW_THREADS / 32];
  const int tid = (int)threadIdx.x;
  const uint32_t raw_cap = (a.header_bytes + a.unit_bytes * BW_ITEMS + 3u) & ~3u;
  uint8_t *raw = smem;
  uint32_t *out = reinterpret_cast<uint32_t *>(smem + raw_cap);
  const uint32_t out_words = (bw_comp_bound(a.header_bytes + a.unit_bytes * BW_ITEMS) + 3) / 4 + 1;
  for (uint32_t s = blockIdx.x; s < a.n_sections; s += gridDim.x) {
    const BwSection sec = a.sections[s];
    const uint32_t n_raw = a.header_bytes + sec.count * a.unit_bytes;
    // ---- assemble the section ----
    {
      const uint32_t *src = reinterpret_cast<const uint32_t *>(a.units + sec.first * a.unit_bytes);  // units are 4-byte aligned
      uint32_t *dst = reinterpret_cast<uint32_t *>(raw + a.header_bytes);
      for (uint32_t i = (uint32_t)tid; i < sec.count * a.unit_bytes / 4; i += BW_THREADS) dst[i] = src[i];
      if (a.compress)
        for (uint32_t i = (uint32_t)tid; i < out_words; i += BW_THREADS) out[i] = 0;
    }
    __syncthreads();
    uint32_t first_start, last_end;
    if (a.unit_bytes == BW_ITEM_BYTES) {
      const BwItem *it = reinterpret_cast<const BwItem *>(raw + a.header_bytes);
      first_start = it[0].start;
      last_end = it[sec.count - 1].end;
    } else {
      const uint32_t *z = reinterpret_cast<const uint32_t *>(raw + a.header_bytes);
      first_start = z[1];
      last_end = z[(sec.count - 1) * 8 + 2];
    }
    if (a.header_bytes && tid == 0) {  // bedGraph section header: chromId, chromStart, chromEnd, itemStep, itemSpan, type 1, reserved, itemCount
      uint32_t *h = reinterpret_cast<uint32_t *>(raw);
      h[0] = sec.chrom, h[1] = first_start, h[2] = last_end, h[3] = 0, h[4] = 0;
      h[5] = 1u | (sec.count << 16);
    }
    if (tid == 0) a.bounds[2 * s] = first_start, a.bounds[2 * s + 1] = last_end;
    // ---- summary statistics of a data section (fixed reduction order: reproducible) ----
    if (a.stats && a.unit_bytes == BW_ITEM_BYTES) {
      const BwItem *it = reinterpret_cast<const BwItem *>(raw + a.header_bytes);
      double cov = 0, mn = 1e300, mx = -1e300, sum = 0, sq = 0;
      for (uint32_t i = (uint32_t)tid; i < sec.count; i += BW_THREADS) {
        const double w = (double)(it[i].end - it[i].start), v = (double)it[i].val;
        cov += w, sum += v * w, sq += v * v * w;
        mn = v < mn ? v : mn, mx = v > mx ? v : mx;
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        cov += __shfl_xor_sync(0xffffffffu, cov, d);
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        sq += __shfl_xor_sync(0xffffffffu, sq, d);
        const double m1 = __shfl_xor_sync(0xffffffffu, mn, d), m2 = __shfl_xor_sync(0xffffffffu, mx, d);
        mn = m1 < mn ? m1 : mn, mx = m2 > mx ? m2 : mx;
      }
      if ((tid & 31) == 0) s_red[0][tid >> 5] = cov, s_red[1][tid >> 5] = mn, s_red[2][tid >> 5] = mx, s_red[3][tid >> 5] = sum, s_red[4][tid >> 5] = sq;
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < BW_THREADS / 32; w++) {
          cov += s_red[0][w], sum += s_red[3][w], sq += s_red[4][w];
          mn = s_red[1][w] < mn ? s_red[1][w] : mn, mx = s_red[2][w] > mx ? s_red[2][w] : mx;
        }
        double *o = a.stats + 5 * (size_t)s;
        o[0] = cov, o[1] = mn, o[2] = mx, o[3] = sum, o[4] = sq;
      }
    }
    __syncthreads();
    uint8_t *dst = a.staging + (size_t)s * a.stride;
    if (!a.compress) {
      for (uint32_t i = (uint32_t)tid; i < (n_raw + 3) / 4; i += BW_THREADS) reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<uint32_t *>(raw)[i];
      if (tid == 0) a.comp_size[s] = n_raw;
      __syncthreads();
      continue;
    }
    // ---- every thread owns a contiguous run of units (thread 0 also the header) ----
    const uint32_t upt = (sec.count + BW_THREADS - 1) / BW_THREADS;
    const uint32_t k0 = tmin<uint32_t>((uint32_t)tid * upt, sec.count), k1 = tmin<uint32_t>(k0 + upt, sec.count);
    uint32_t lags[4] = {a.dist[0], a.dist[1], a.dist[2], a.dist[3]};
    uint32_t my_bits = 0;
    if (tid == 0 && a.header_bytes) my_bits += bw_parse_unit<false>(raw, 0, a.header_bytes, lags, out, 0);
    for (uint32_t k = k0; k < k1; k++) my_bits += bw_parse_unit<false>(raw, a.header_bytes + k * a.unit_bytes, a.header_bytes + (k + 1) * a.unit_bytes, lags, out, 0);
    uint64_t total;
    const uint64_t ex = block_exclusive_scan<BW_THREADS, uint64_t>((uint64_t)my_bits, scratch, total);
    uint32_t bp = 16 + 3 + (uint32_t)ex;  // zlib header (2 bytes), BFINAL = 1 + BTYPE = 01
    if (tid == 0) {
      bw_put(out, 0, 0x9C78u, 16);  // CMF 0x78, FLG 0x9C
      bw_put(out, 16, 3u, 3);
      if (a.header_bytes) bp += bw_parse_unit<true>(raw, 0, a.header_bytes, lags, out, bp);
    }
    for (uint32_t k = k0; k < k1; k++) bp += bw_parse_unit<true>(raw, a.header_bytes + k * a.unit_bytes, a.header_bytes + (k + 1) * a.unit_bytes, lags, out, bp);
    // Adler-32 of the raw bytes: a = 1 + sum b_i, b = n + sum (n - i) b_i, both mod 65521
    uint64_t sa = 0, sb = 0;
    for (uint32_t i = (uint32_t)tid; i < n_raw; i += BW_THREADS) {
      const uint64_t x = raw[i];
      sa += x;
      sb += (uint64_t)(n_raw - i) * x;
    }
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    __syncthreads();  // scan scratch is free again
    if ((tid & 31) == 0) scratch[tid >> 5] = sa;
    __syncthreads();
    uint64_t ta = 0;
    for (int w = 0; w < BW_THREADS / 32; w++) ta += scratch[w];
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = sb;
    __syncthreads();
    uint64_t tb = 0;
    for (int w = 0; w < BW_THREADS / 32; w++) tb += scratch[w];
    const uint32_t end_bits = 19 + (uint32_t)total + 7;  // + end-of-block (symbol 256 = seven zero bits: nothing to OR)
    const uint32_t body = (end_bits + 7) / 8;
    const uint32_t csize = body + 4;
    __syncthreads();  // all bits are in the image
    if (tid == 0) {
      const uint32_t ad_a = (uint32_t)((1 + ta) % 65521u), ad_b = (uint32_t)((n_raw + tb) % 65521u);
      const uint32_t adler = (ad_b << 16) | ad_a;
      uint8_t *ob = reinterpret_cast<uint8_t *>(out);
      ob[body] = (uint8_t)(adler >> 24), ob[body + 1] = (uint8_t)(adler >> 16), ob[body + 2] = (uint8_t)(adler >> 8), ob[body + 3] = (uint8_t)adler;
      a.comp_size[s] = csize;
    }
    __syncthreads();
    for (uint32_t i = (uint32_t)tid; i < (csize + 3) / 4; i += BW_THREADS) reinterpret_cast<uint32_t *>(dst)[i] = out[i];
    __syncthreads();
  }
}

// Sections back to back: one warp per section, byte copies widened to words where both sides are aligned.
__global__ void __launch_bounds__(256) bw_compact_kernel(BwCompactArgs a) {
  const int lane = (int)(threadIdx.x & 31u);
  for (uint32_t s = (uint32_t)((blockIdx.x * 256ull + threadIdx.x) >> 5); s < a.n_sections; s += (uint32_t)((gridDim.x * 256ull) >> 5)) {
    const uint8_t *src = a.staging + (size_t)s * a.stride;
    uint8_t *dst = a.blob + a.offset[s];
    const uint32_t n = a.comp_size[s];
    const uint32_t lead = tmin<uint32_t>((uint32_t)((4 - ((uintptr_t)dst & 3)) & 3), n);
    for (uint32_t i = (uint32_t)lane; i < lead; i += 32) dst[i] = src[i];
    // destination word j holds source bytes lead + 4j ..: assembled from two aligned source words
    const uint32_t nw = (n - lead) / 4;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src);
    uint32_t *dw = reinterpret_cast<uint32_t *>(dst + lead);
    const uint32_t sh = (lead & 3u) * 8;
    for (uint32_t j = (uint32_t)lane; j < nw; j += 32) {
      const uint32_t w0 = sw[(lead >> 2) + j], w1 = sh ? sw[(lead >> 2) + j + 1] : 0u;
      dw[j] = sh ? __funnelshift_r(w0, w1, sh) : w0;
    }
    for (uint32_t i = lead + nw * 4 + (uint32_t)lane; i < n; i += 32) dst[i] = src[i];
  }
}

// ---- zoom levels -------------------------------------------------------------------------------------------------------
// One thread per window of `reduction` bases.  A window that overlaps at least one item yields one summary record:
// covered bases, min, max, sum and sum of squares weighted by the overlap, accumulated in f32 in item order.
__global__ void __launch_bounds__(BW_THREADS) bw_zoom_windows_kernel(BwZoomArgs a) {
  __shared__ uint64_t scratch[BW_THREADS / 32 + 1];
  const int tid = (int)threadIdx.x;
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t wi = (uint64_t)t * BW_THREADS + (uint64_t)tid;
    uint64_t flag = 0;
    if (wi < a.n_windows) {
      uint32_t lo = 0, hi = a.n_chrom;  // last chromosome whose first window is <= wi
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.win_first[mid] <= wi) lo = mid;
        else hi = mid;
      }
      const uint32_t c = lo;
      if (a.prev_recs) {
        // level z from level z - 1 (reduction x4, windows of both levels start at base 0 of the chromosome): fold the up to four
        // finer windows that hold data, in order
        const uint64_t w = wi - a.win_first[c];
        const uint64_t p0 = a.prev_win_first[c] + 4 * w, p1 = tmin<uint64_t>(p0 + 4, a.prev_win_first[c + 1]);
        uint32_t rs = 0, re = 0, valid = 0;
        float mn = 0.f, mx = 0.f, sum = 0.f, sq = 0.f;
        for (uint64_t q = p0; q < p1; q++) {
          if (!a.prev_flags[q]) continue;
          const uint32_t *r = reinterpret_cast<const uint32_t *>(a.prev_recs + q * BW_ZOOM_BYTES);
          const float qmn = __uint_as_float(r[4]), qmx = __uint_as_float(r[5]), qs = __uint_as_float(r[6]), qq = __uint_as_float(r[7]);
          if (!flag) {
            rs = r[1], mn = qmn, mx = qmx, sum = qs, sq = qq;
            flag = 1;
          } else {
            mn = fminf(mn, qmn), mx = fmaxf(mx, qmx);
            sum = __fadd_rn(sum, qs), sq = __fadd_rn(sq, qq);
          }
          re = r[2];
          valid += r[3];
        }
        if (flag) {
          uint32_t *r = reinterpret_cast<uint32_t *>(a.recs + wi * BW_ZOOM_BYTES);
          r[0] = c, r[1] = rs, r[2] = re, r[3] = valid;
          r[4] = __float_as_uint(mn), r[5] = __float_as_uint(mx), r[6] = __float_as_uint(sum), r[7] = __float_as_uint(sq);
        }
        a.flags[wi] = (uint8_t)flag;
      } else {
      const uint64_t ws = (wi - a.win_first[c]) * (uint64_t)a.reduction, we = ws + a.reduction;
      uint64_t i0 = a.item_first[c], i1 = a.item_first[c + 1];
      uint64_t l = i0, h = i1;  // first item whose end lies beyond the window start
      while (l < h) {
        const uint64_t mid = (l + h) >> 1;
        if ((uint64_t)a.items[mid].end <= ws) l = mid + 1;
        else h = mid;
      }
      if (l < i1 && (uint64_t)a.items[l].start < we) {
        flag = 1;
        uint32_t rs = 0, re = 0, valid = 0;
        float mn = 0.f, mx = 0.f, sum = 0.f, sq = 0.f;
        bool first = true;
        for (uint64_t k = l; k < i1; k++) {
          const BwItem it = a.items[k];
          if ((uint64_t)it.start >= we) break;
          const uint32_t s0 = (uint64_t)it.start > ws ? it.start : (uint32_t)ws, e0 = (uint64_t)it.end < we ? it.end : (uint32_t)we;
          const float n = (float)(e0 - s0);
          if (first) {
            rs = s0, mn = mx = it.val;
            sum = __fmul_rn(it.val, n);
            sq = __fmul_rn(__fmul_rn(it.val, it.val), n);
            first = false;
          } else {
            mn = fminf(mn, it.val), mx = fmaxf(mx, it.val);
            sum = __fadd_rn(sum, __fmul_rn(it.val, n));
            sq = __fadd_rn(sq, __fmul_rn(__fmul_rn(it.val, it.val), n));
          }
          re = e0;
          valid += e0 - s0;
        }
        uint32_t *r = reinterpret_cast<uint32_t *>(a.recs + wi * BW_ZOOM_BYTES);
        r[0] = c, r[1] = rs, r[2] = re, r[3] = valid;
        r[4] = __float_as_uint(mn), r[5] = __float_as_uint(mx), r[6] = __float_as_uint(sum), r[7] = __float_as_uint(sq);
      }
      a.flags[wi] = (uint8_t)flag;
      }
    }
    uint64_t total;
    block_exclusive_scan<BW_THREADS, uint64_t>(flag, scratch, total);
    if (tid == 0) a.tile_count[t] = total;
  }
}

__global__ void __launch_bounds__(BW_THREADS) bw_zoom_compact_kernel(BwZoomArgs a) {
  __shared__ uint64_t scratch[BW_THREADS / 32 + 1];
  const int tid = (int)threadIdx.x;
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t wi = (uint64_t)t * BW_THREADS + (uint64_t)tid;
    const uint64_t flag = wi < a.n_windows ? a.flags[wi] : 0;
    uint64_t total;
    const uint64_t pos = a.tile_count[t] + block_exclusive_scan<BW_THREADS, uint64_t>(flag, scratch, total);
    if (flag) {
      const uint4 *src = reinterpret_cast<const uint4 *>(a.recs + wi * BW_ZOOM_BYTES);
      uint4 *dst = reinterpret_cast<uint4 *>(a.out + pos * BW_ZOOM_BYTES);
      dst[0] = src[0], dst[1] = src[1];
    }
  }
}

// first compacted record of every chromosome (and the total at [n_chrom])
__global__ void bw_zoom_chrom_first_kernel(BwZoomArgs a, unsigned long long total) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > a.n_chrom) return;
  if (c == a.n_chrom) {
    a.chrom_first_rec[c] = total;
    return;
  }
  const uint64_t wi = a.win_first[c];
  if (wi >= a.n_windows) {
    a.chrom_first_rec[c] = total;
    return;
  }
  const uint64_t t = wi / BW_THREADS;
  uint64_t pos = a.tile_count[t];
  for (uint64_t k = t * BW_THREADS; k < wi; k++) pos += a.flags[k];
  a.chrom_first_rec[c] = pos;
}

}  // namespace bnd
