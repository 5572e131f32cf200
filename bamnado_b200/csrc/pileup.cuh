// pileup.cuh — K3 + K4: per-record field parse, read filter, interval emission and +1/-1 scatter.
//
// One thread per BAM record.  Replaces, for every (record, genomic chunk) incidence the reference visits
// (bamnado/src/coverage_analysis.rs:461-484):
//   * the noodles `query` intersection test (record alignment [pos1, pos1+span-1] vs the 1-based inclusive chunk),
//   * BamReadFilter::is_valid                     (bamnado/src/read_filter.rs:570-803)  -> 13 counters,
//   * IntervalMaker::coords / coords_read / coords_fragment / coords_shifted / truncate
//                                                  (bamnado/src/genomic_intervals.rs:206-402),
//   * Lapper::new + bin_intervals' per-bin Lapper::count (bamnado/src/bam_utils.rs:48-91), restated as a
//     difference array: count_i = #{start < e_{i+1}} - #{stop <= e_i} over the chunk's bin edges e_i, i.e. +1 at
//     the first bin whose right edge exceeds `start`, -1 at the first bin whose left edge reaches `stop`, and a
//     compensating entry at the first bin of the next chunk so that one plain prefix sum over the whole flat bin
//     array reproduces the per-chunk counts (a read is visited once per chunk it overlaps).
// Scatter goes to a shared-memory window of bins private to the CTA (the BAM is coordinate sorted, so a CTA's
// records fall into a narrow window) and is flushed with global atomics; bins outside the window use global
// atomics directly.  The kernel is a byte/integer gather: HBM- and atomic-bound, no tensor-core shape.
#pragma once
#include "common.cuh"
#include "records.cuh"

namespace bnd {


// ---- aux field lookup: noodles `data.get(tag)` = first field with that tag ---------------------------------------
// returns 0 not found, 1 found as Z string (val/len set), 2 found with another type, -1 malformed (an Err upstream)
__device__ __forceinline__ int aux_find_string(const uint8_t *a, const uint8_t *e, char t0, char t1, const uint8_t *&val, uint32_t &len) {
  if (a > e) return -1;
#pragma unroll 1
  while (a < e) {
    if (e - a < 3) return -1;
    char a0 = (char)a[0], a1 = (char)a[1];
    uint8_t ty = a[2];
    a += 3;
    uint64_t sz = 0;
    bool is_z = false;
    switch (ty) {
      case 'A': case 'c': case 'C': sz = 1; break;
      case 's': case 'S': sz = 2; break;
      case 'i': case 'I': case 'f': sz = 4; break;
      case 'Z': case 'H': {
        const uint8_t *q = a;
        while (q < e && *q != 0) q++;
        if (q >= e) return -1;
        sz = (uint64_t)(q - a) + 1;
        is_z = ty == 'Z';
        break;
      }
      case 'B': {
        if (e - a < 5) return -1;
        uint8_t st = a[0];
        uint32_t n = (uint32_t)a[1] | ((uint32_t)a[2] << 8) | ((uint32_t)a[3] << 16) | ((uint32_t)a[4] << 24);
        uint64_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : (st == 'i' || st == 'I' || st == 'f') ? 4 : 0;
        if (!es) return -1;
        sz = 5 + (uint64_t)n * es;
        break;
      }
      default: return -1;
    }
    if ((uint64_t)(e - a) < sz) return -1;
    if (a0 == t0 && a1 == t1) {
      if (is_z) {
        val = a;
        len = (uint32_t)(sz - 1);
        return 1;
      }
      return 2;
    }
    a += sz;
  }
  return 0;
}

__device__ __forceinline__ bool bytes_equal(const uint8_t *a, const char *b, uint32_t n) {
  for (uint32_t i = 0; i < n; i++)
    if (a[i] != (uint8_t)b[i]) return false;
  return true;
}

__host__ __device__ __forceinline__ uint64_t fnv1a64(const uint8_t *p, uint32_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (uint32_t i = 0; i < n; i++) {
    h ^= p[i];
    h *= 0x100000001b3ull;
  }
  return h;
}

__device__ __forceinline__ bool barcode_allowed(const FilterDev &f, const uint8_t *val, uint32_t len) {
  uint32_t slot = (uint32_t)(fnv1a64(val, len) >> 17) & f.bc_mask;
#pragma unroll 1
  for (;;) {
    uint32_t e = __ldg(f.bc_table + slot);
    if (e == 0) return false;
    uint32_t o0 = __ldg(f.bc_off + e - 1), o1 = __ldg(f.bc_off + e);
    if (o1 - o0 == len && bytes_equal(val, f.bc_pool + o0, len)) return true;
    slot = (slot + 1) & f.bc_mask;
  }
}

template <class T>
__device__ __forceinline__ uint64_t lower_bound_dev(const T *a, uint64_t n, T key) {  // first i with a[i] >= key
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ uint64_t sat_sub64(uint64_t a, uint64_t b) { return a > b ? a - b : 0; }
__device__ __forceinline__ uint64_t sat_add64(uint64_t a, uint64_t b) { return a + b < a ? ~0ull : a + b; }

struct RecView {
  const uint8_t *r;  // points at block_size
  uint32_t bs;
  int32_t ref, pos0, next_pos0, tlen;
  uint32_t lrn, mapq, n_cigar, flag, l_seq;
};

__device__ __forceinline__ void rec_load(const uint8_t *r, RecView &v) {
  v.r = r;
  v.bs = ld_u32_unaligned(r);
  v.ref = (int32_t)ld_u32_unaligned(r + 4);
  v.pos0 = (int32_t)ld_u32_unaligned(r + 8);
  uint32_t w3 = ld_u32_unaligned(r + 12), w4 = ld_u32_unaligned(r + 16);
  v.lrn = w3 & 0xff;
  v.mapq = (w3 >> 8) & 0xff;
  v.n_cigar = w4 & 0xffff;
  v.flag = w4 >> 16;
  v.l_seq = ld_u32_unaligned(r + 20);
  v.next_pos0 = (int32_t)ld_u32_unaligned(r + 28);
  v.tlen = (int32_t)ld_u32_unaligned(r + 32);
}

// alignment_span (genomic_intervals.rs:27-44): M, D, N, =, X consume the reference
__device__ __forceinline__ uint64_t rec_span(const RecView &v) {
  const uint8_t *c = v.r + 36 + v.lrn;
  uint64_t span = 0;
#pragma unroll 1
  for (uint32_t i = 0; i < v.n_cigar; i++) {
    uint32_t w = ld_u32_unaligned(c + 4 * (size_t)i);
    uint32_t op = w & 0xf;
    if ((0x18Du >> op) & 1u) span += w >> 4;  // ops 0,2,3,7,8
  }
  return span;
}

// BamReadFilter::is_valid.  Returns 1 valid, 0 filtered (counter index in `fail`), -1 Err (dropped, no counter).
__device__ __forceinline__ int filter_is_valid(const FilterDev &f, const RecView &v, int &fail) {
  const uint32_t flag = v.flag;
  if (f.ex_dup && (flag & 0x400)) return fail = ST_FAIL_DUPLICATE, 0;
  if (f.ex_sec && (flag & 0x100)) return fail = ST_FAIL_SECONDARY, 0;
  if (f.ex_sup && (flag & 0x800)) return fail = ST_FAIL_SUPPLEMENTARY, 0;
  const bool rev = (flag & 0x10) != 0;
  if (rev && f.strand == 1) return fail = ST_FAIL_STRAND, 0;
  if (!rev && f.strand == 2) return fail = ST_FAIL_STRAND, 0;
  if (f.proper_pair && !(flag & 0x2)) return fail = ST_FAIL_PROPER_PAIR, 0;
  if (flag & 0x4) return fail = ST_FAIL_MAPQ, 0;
  if (v.mapq != 255 && v.mapq < f.min_mapq) return fail = ST_FAIL_MAPQ, 0;  // 255 == missing == passes
  const uint64_t l_seq = v.l_seq;
  if (l_seq < f.min_length || l_seq > f.max_length) return fail = ST_FAIL_LENGTH, 0;
  if (f.has_min_frag || f.has_max_frag) {
    int32_t t = v.tlen;
    uint32_t at = t < 0 ? (uint32_t)(-(int64_t)t) : (uint32_t)t;
    if (f.has_min_frag && at < f.min_frag) return fail = ST_FAIL_FRAGMENT_LENGTH, 0;
    if (f.has_max_frag && at > f.max_frag) return fail = ST_FAIL_FRAGMENT_LENGTH, 0;
  }
  if (v.ref < 0) return -1;
  if (v.pos0 < 0) return fail = ST_FAIL_MAPQ, 0;
  if (f.has_blacklist) {
    uint64_t start = (uint64_t)v.pos0 + 1, end = start + l_seq;
    uint64_t o0 = __ldg(f.bl_off + v.ref), o1 = __ldg(f.bl_off + v.ref + 1);
    if (o1 > o0) {
      uint64_t first = lower_bound_dev<uint64_t>(f.bl_stops + o0, o1 - o0, start + 1);
      uint64_t last = lower_bound_dev<uint64_t>(f.bl_starts + o0, o1 - o0, end);
      if (last != first) return fail = ST_FAIL_BLACKLIST, 0;  // Lapper::count(start, end) > 0 in wrapping usize
    }
  }
  if (f.has_barcodes || f.has_rg || f.has_tag) {
    const uint8_t *aux = v.r + 36 + v.lrn + 4 * (size_t)v.n_cigar + ((size_t)v.l_seq + 1) / 2 + v.l_seq;
    const uint8_t *end = v.r + 4 + v.bs;
    const uint8_t *val = nullptr;
    uint32_t len = 0;
    if (f.has_barcodes) {
      int rc = aux_find_string(aux, end, 'C', 'B', val, len);
      if (rc != 1 || !barcode_allowed(f, val, len)) return fail = ST_FAIL_BARCODE, 0;
    }
    if (f.has_rg) {
      int rc = aux_find_string(aux, end, 'R', 'G', val, len);
      if (rc == 0) return fail = ST_FAIL_READ_GROUP, 0;
      if (rc < 0) return -1;
      if (rc == 2 || len != (uint32_t)f.rg_len || !bytes_equal(val, f.rg, len)) return fail = ST_FAIL_READ_GROUP, 0;
    }
    if (f.has_tag) {
      if (!f.tag_name_ok) return fail = ST_FAIL_TAG, 0;
      int rc = aux_find_string(aux, end, f.tag0, f.tag1, val, len);
      if (rc != 1 || len != (uint32_t)f.tagv_len || !bytes_equal(val, f.tagv, len)) return fail = ST_FAIL_TAG, 0;
    }
  }
  return 1;
}

// IntervalMaker::coords after a passing filter.  false = no interval (the read still counts as retained).
__device__ __forceinline__ bool make_interval(const PileDev &P, const GeomDev &g, const RecView &v, uint64_t span, uint64_t &s_out, uint64_t &e_out) {
  uint64_t start, end;
  const uint32_t flag = v.flag;
  if (P.use_fragment) {
    if (!(flag & 0x40)) return false;
    int32_t t = v.tlen;
    if (t > 0) {
      start = (uint64_t)v.pos0 + 1;
      end = start + (uint64_t)t;
    } else if (t < 0) {
      if (v.next_pos0 < 0) return false;
      start = (uint64_t)v.next_pos0 + 1;
      end = start + (uint64_t)(-(int64_t)t);
    } else {
      return false;
    }
  } else {
    start = (uint64_t)v.pos0 + 1;
    end = start + span;
  }
  if (P.shift_on) {
    const bool rev = (flag & 0x10) != 0, first = (flag & 0x40) != 0;
    const int32_t five = P.shift[0], three = P.shift[1], five_r = P.shift[2], three_r = P.shift[3];
    int32_t ds, de;
    if (rev && first) ds = three, de = five_r;
    else if (rev && !first) ds = three_r, de = five_r;
    else if (!rev && first) ds = five, de = three;
    else ds = five, de = three_r;
    start = ds >= 0 ? sat_sub64(start, (uint64_t)ds) : sat_add64(start, (uint64_t)(-(int64_t)ds));
    end = de >= 0 ? sat_add64(end, (uint64_t)de) : sat_sub64(end, (uint64_t)(-(int64_t)de));
    start = tmax<uint64_t>(start, 1);
    end = tmin<uint64_t>(__ldg(g.ref_len + v.ref), end);
  }
  if (P.has_trunc) {
    if (P.t_has5) start = sat_add64(start, P.t5);
    if (P.t_has3) end = sat_sub64(end, P.t3);
  }
  s_out = start;
  e_out = end;
  return true;
}

__device__ __forceinline__ void bin_add(int32_t *diff, int32_t *win, uint64_t win_base, uint64_t idx, int32_t v) {
  uint64_t rel = idx - win_base;
  if (rel < PILE_WIN) atomicAdd(&win[rel], v);
  else atomicAdd(&diff[idx], v);
}

__global__ void __launch_bounds__(PILE_THREADS) pileup_kernel(PileArgs a) {
  __shared__ int32_t win[PILE_WIN];
  __shared__ unsigned int s_stats[ST_N_COUNTERS];
  __shared__ unsigned long long s_win_base;
  const int tid = (int)threadIdx.x;
  const uint64_t n_rec = a.st->n_records;
  const GeomDev &g = a.g;
  constexpr uint64_t TILE = (uint64_t)PILE_THREADS * PILE_RPT;
  if (tid < ST_N_COUNTERS) s_stats[tid] = 0;
  for (int i = tid; i < (int)PILE_WIN; i += PILE_THREADS) win[i] = 0;
  __syncthreads();
  for (uint64_t tile0 = blockIdx.x * TILE; tile0 < n_rec; tile0 += (uint64_t)gridDim.x * TILE) {
    // ---- window placement from the tile's first record ----
    if (tid == 0) {
      uint64_t base = ~0ull >> 1;  // far away: everything goes straight to global memory
      if (a.use_window) {
        RecView v;
        rec_load(a.buf + a.rec_off[tile0], v);
        if (v.ref >= 0 && v.ref < g.n_ref && v.pos0 >= 0 && g.nbf > 0) {
          uint64_t p1 = (uint64_t)v.pos0 + 1;
          uint64_t c = (p1 - 1) / g.L;
          uint64_t gb = __ldg(g.ref_first_bin + v.ref) + c * g.nbf + (p1 - (c * g.L + 1)) / g.bin;
          uint64_t rel = gb >= g.bin_lo ? gb - g.bin_lo : 0;
          base = rel > PILE_WIN_MARGIN ? rel - PILE_WIN_MARGIN : 0;
        }
      }
      s_win_base = base;
    }
    __syncthreads();
    const uint64_t win_base = s_win_base;
#pragma unroll 1
    for (int it = 0; it < PILE_RPT; it++) {
      const uint64_t i = tile0 + (uint64_t)it * PILE_THREADS + (uint64_t)tid;
      if (i >= n_rec) break;
      RecView v;
      rec_load(a.buf + a.rec_off[i], v);
      if (i == n_rec - 1) {
        unsigned long long key = ((unsigned long long)(uint32_t)(v.ref + 1) << 32) | (uint32_t)(v.pos0 + 1);
        if (v.ref < 0) key = ~0ull;  // unplaced reads sort last
        atomicMax(a.range_probe, key);
      }
      // noodles `query`/`intersects`: same reference, has start and a non-empty alignment, overlaps the chunk
      if (v.ref < 0 || v.ref >= g.n_ref || v.pos0 < 0) continue;
      if (32ull + v.lrn + 4ull * v.n_cigar > v.bs) continue;  // undecodable record: silently dropped upstream
      const uint64_t span = rec_span(v);
      if (span == 0) continue;
      const uint64_t a_start = (uint64_t)v.pos0 + 1, a_end = a_start + span - 1;
      const uint64_t len = __ldg(g.ref_len + v.ref);
      const uint64_t ch0 = __ldg(g.ref_first_chunk + v.ref), ch1 = __ldg(g.ref_first_chunk + v.ref + 1);
      if (ch1 == ch0 || a_start > len) continue;  // reference without chunks / read beyond the last chunk
      uint64_t c_first = (a_start - 1) / g.L;
      uint64_t c_last = tmin<uint64_t>((a_end - 1) / g.L, ch1 - ch0 - 1);
      // clip to the chunks this shard owns
      if (ch0 + c_first < g.chunk_lo) c_first = g.chunk_lo - ch0;
      if (ch0 + c_last >= g.chunk_hi) {
        if (g.chunk_hi <= ch0) continue;
        c_last = g.chunk_hi - ch0 - 1;
      }
      if (c_first > c_last) continue;
      const uint32_t n_inc = (uint32_t)(c_last - c_first + 1);
      int fail = -1;
      const int valid = filter_is_valid(a.f, v, fail);
      atomicAdd(&s_stats[ST_N_TOTAL], n_inc);
      if (valid == 0) atomicAdd(&s_stats[fail], n_inc);
      if (valid != 1) continue;
      uint64_t s, e;
      if (!make_interval(a.p, g, v, span, s, e)) continue;
      const uint64_t fb = __ldg(g.ref_first_bin + v.ref);
#pragma unroll 1
      for (uint64_t c = c_first; c <= c_last; c++) {
        const uint64_t cs = c * g.L + 1, ce = tmin<uint64_t>(cs + g.L - 1, len);
        const uint64_t nb = (ce - cs + g.bin - 1) / g.bin;
        const uint64_t gb0 = fb + c * g.nbf - g.bin_lo;
        int32_t net = 0;
        if (s < ce) {
          uint64_t ia = s <= cs ? 0 : (s - cs) / g.bin;
          bin_add(a.diff, win, win_base, gb0 + ia, 1);
          net += 1;
        }
        uint64_t ib = e <= cs ? 0 : (e - cs + g.bin - 1) / g.bin;
        if (ib < nb) {
          bin_add(a.diff, win, win_base, gb0 + ib, -1);
          net -= 1;
        }
        if (net != 0) bin_add(a.diff, win, win_base, gb0 + nb, -net);
      }
    }
    __syncthreads();
    // ---- flush the window ----
    if (a.use_window) {
      for (uint32_t i = (uint32_t)tid; i < PILE_WIN; i += PILE_THREADS) {
        int32_t d = win[i];
        if (d != 0) {
          atomicAdd(&a.diff[win_base + i], d);
          win[i] = 0;
        }
      }
    }
    __syncthreads();
  }
  if (tid < ST_N_COUNTERS && s_stats[tid]) atomicAdd(&a.stats[tid], (unsigned long long)s_stats[tid]);
  if (tid == 0 && blockIdx.x == 0) atomicAdd(a.range_probe + 1, (unsigned long long)n_rec);
}

}  // namespace bnd
