// bedgraph_parse.cuh — the reader of `collapse-bedgraph` on the device (bamnado/src/cli.rs:1149-1176): the raw bytes of the
// bedGraph go to HBM once; lines, fields, integers and scores are found there instead of in a getline / strtod loop.
//
//   bdg_newline_count / bdg_line_starts   newline positions: per-tile counts, scan, offsets of every line start
//   bdg_parse_lines      one thread per line: trim, skip blank and '#' lines and lines with fewer than four fields
//                        (`l.trim()`, `split_whitespace`, `parts.len() < 4`), parse start / end as i64 and the score as f64.
//                        A decimal with at most 19 significant digits whose value is an exact double times / over an exact power
//                        of ten (|exponent| <= 22, mantissa < 2^53: one correctly rounded IEEE operation) is converted here;
//                        everything else (17-digit scores, inf / nan spellings) is appended to a short list the host converts.
//   bdg_compact_rows     drops the skipped lines (only launched when there are any)
//   bdg_name_heads / bdg_name_segments / bdg_assign_chrom
//                        rows whose chromosome name differs from the previous row's open a name segment; the host ranks the
//                        (few) segment names bytewise and every row gets its chromosome rank from its segment
//   bdg_gather_rows      applies the sort permutation when the rows are not already in (chrom, start) order
// All of it is byte / integer streaming work bound by HBM bandwidth; the collapse itself is collapse.cuh.
#pragma once
#include "collapse.cuh"

namespace bnd {

__device__ __forceinline__ bool bdg_is_ws(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }  // char::is_whitespace, ASCII part

__global__ void __launch_bounds__(BDG_THREADS) bdg_newline_count_kernel(BdgParseArgs a) {
  __shared__ uint64_t s_part[BDG_THREADS / 32];
  const int tid = (int)threadIdx.x;
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t p0 = (uint64_t)t * BDG_TILE + (uint64_t)tid * BDG_PER_THREAD;
    uint64_t n = 0;
    for (uint32_t k = 0; k < BDG_PER_THREAD; k++)
      if (p0 + k < a.n_bytes && a.text[p0 + k] == '\n') n++;
    n = warp_sum(n);
    if ((tid & 31) == 0) s_part[tid >> 5] = n;
    __syncthreads();
    if (tid == 0) {
      uint64_t S = 0;
      for (int w = 0; w < BDG_THREADS / 32; w++) S += s_part[w];
      a.tile_count[t] = S;
    }
    __syncthreads();
  }
}

// line_off[k] = first byte of line k (line 0 starts at byte 0; the caller writes the sentinel line_off[n_lines])
__global__ void __launch_bounds__(BDG_THREADS) bdg_line_starts_kernel(BdgParseArgs a) {
  __shared__ uint64_t scratch[BDG_THREADS / 32 + 1];
  const int tid = (int)threadIdx.x;
  if (blockIdx.x == 0 && tid == 0) a.line_off[0] = 0;
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t p0 = (uint64_t)t * BDG_TILE + (uint64_t)tid * BDG_PER_THREAD;
    uint64_t n = 0, total;
    for (uint32_t k = 0; k < BDG_PER_THREAD; k++)
      if (p0 + k < a.n_bytes && a.text[p0 + k] == '\n') n++;
    uint64_t j = a.tile_count[t] + block_exclusive_scan<BDG_THREADS, uint64_t>(n, scratch, total);
    for (uint32_t k = 0; k < BDG_PER_THREAD && n; k++)
      if (p0 + k < a.n_bytes && a.text[p0 + k] == '\n') {
        j++;
        if (j < a.n_lines) a.line_off[j] = p0 + k + 1;  // a file that ends in '\n' has no line after its last newline
      }
  }
}

// str::parse::<i64>: one optional sign, at least one digit, nothing else, no overflow
__device__ __forceinline__ bool bdg_parse_i64(const uint8_t *s, uint32_t n, long long &out) {
  uint32_t i = 0;
  bool neg = false;
  if (n && (s[0] == '+' || s[0] == '-')) neg = s[0] == '-', i = 1;
  if (i >= n) return false;
  unsigned long long v = 0;
  for (; i < n; i++) {
    const uint32_t d = (uint32_t)s[i] - '0';
    if (d > 9) return false;
    if (v > 922337203685477580ull || (v == 922337203685477580ull && d > (neg ? 8u : 7u))) return false;
    v = v * 10 + d;
  }
  out = neg ? (long long)(0ull - v) : (long long)v;
  return true;
}

// 0: not a number this routine converts (the host decides), 1: converted
__device__ __forceinline__ int bdg_parse_f64_fast(const uint8_t *s, uint32_t n, double &out) {
  uint32_t i = 0;
  bool neg = false;
  if (n && (s[0] == '+' || s[0] == '-')) neg = s[0] == '-', i = 1;
  unsigned long long m = 0;
  int nd = 0, digits = 0, e10 = 0;
  bool dot = false;
  for (; i < n; i++) {
    const uint8_t c = s[i];
    if (c == '.') {
      if (dot) return 0;
      dot = true;
      continue;
    }
    const uint32_t d = (uint32_t)c - '0';
    if (d > 9) break;
    digits++;
    if (nd == 0 && d == 0) {  // leading zero: only moves the decimal point
      if (dot) e10--;
      continue;
    }
    if (nd == 19) return 0;  // more significant digits than a u64 holds
    m = m * 10 + d;
    nd++;
    if (dot) e10--;
  }
  if (digits == 0) return 0;  // ".", "e5", "inf", "nan", ...
  if (i < n) {
    if (s[i] != 'e' && s[i] != 'E') return 0;
    i++;
    bool eneg = false;
    if (i < n && (s[i] == '+' || s[i] == '-')) eneg = s[i] == '-', i++;
    if (i >= n) return 0;
    int ex = 0;
    for (; i < n; i++) {
      const uint32_t d = (uint32_t)s[i] - '0';
      if (d > 9) return 0;
      if (ex < 100000) ex = ex * 10 + (int)d;
    }
    e10 += eneg ? -ex : ex;
  }
  double v;
  if (m == 0) {
    v = 0.0;
  } else {
    if (m >> 53 || e10 < -22 || e10 > 22) return 0;
    const double p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                            1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    v = e10 < 0 ? __ddiv_rn((double)m, p10[-e10]) : __dmul_rn((double)m, p10[e10]);
  }
  out = neg ? -v : v;
  return 1;
}

__global__ void __launch_bounds__(256) bdg_parse_lines_kernel(BdgParseArgs a) {
  for (uint64_t li = blockIdx.x * 256ull + threadIdx.x; li < a.n_lines; li += (uint64_t)gridDim.x * 256ull) {
    uint64_t p = a.line_off[li], e = a.line_off[li + 1] - 1;  // [p, e): the line without its '\n'
    if (e > a.n_bytes) e = a.n_bytes;
    while (p < e && bdg_is_ws(a.text[p])) p++;
    while (e > p && bdg_is_ws(a.text[e - 1])) e--;
    uint8_t st = 0;
    if (p < e && a.text[p] != '#') {
      uint64_t fs[4], fe[4];
      int nf = 0;
      uint64_t q = p;
      while (q < e && nf < 4) {
        while (q < e && bdg_is_ws(a.text[q])) q++;
        if (q >= e) break;
        fs[nf] = q;
        while (q < e && !bdg_is_ws(a.text[q])) q++;
        fe[nf++] = q;
      }
      if (nf == 4) {
        long long s = 0, en = 0;
        double sc = 0.0;
        const bool ok = fe[1] - fs[1] < 64 && fe[2] - fs[2] < 64 && bdg_parse_i64(a.text + fs[1], (uint32_t)(fe[1] - fs[1]), s) &&
                        bdg_parse_i64(a.text + fs[2], (uint32_t)(fe[2] - fs[2]), en);
        if (!ok) {
          st = 3;
          atomicMin(a.bad_line, (unsigned long long)li);
        } else {
          st = 1;
          const uint64_t sl = fe[3] - fs[3];
          if (sl >= 4096 || !bdg_parse_f64_fast(a.text + fs[3], (uint32_t)sl, sc)) {
            st = 2;
            const unsigned long long k = atomicAdd(a.n_slow, 1ull);
            if (k < a.slow_cap) a.slow[k] = BdgSlow{li, fs[3], (uint32_t)(sl < 0xffffffffull ? sl : 0xffffffffull), 0u};
          }
          a.start[li] = s;
          a.end[li] = en;
          a.score[li] = sc;
          a.name_off[li] = fs[0];
          a.name_len[li] = (uint32_t)(fe[0] - fs[0]);
        }
      }
    }
    a.state[li] = st;
  }
}

// scores the host converted: score[line] = value
__global__ void __launch_bounds__(256) bdg_scatter_slow_kernel(const BdgSlow *slow, const double *vals, uint64_t n, double *score) {
  for (uint64_t k = blockIdx.x * 256ull + threadIdx.x; k < n; k += (uint64_t)gridDim.x * 256ull) score[slow[k].line] = vals[k];
}

// ---- rows = lines that hold data -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(COL_THREADS) bdg_row_count_kernel(BdgRowArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t i = (uint64_t)t * COL_THREADS + threadIdx.x;
    uint64_t h = i < a.n_lines && a.state[i] != 0 ? 1 : 0, total;
    block_exclusive_scan<COL_THREADS, uint64_t>(h, scratch, total);
    if (threadIdx.x == 0) a.tile_count[t] = total;
  }
}

__global__ void __launch_bounds__(COL_THREADS) bdg_compact_rows_kernel(BdgRowArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t i = (uint64_t)t * COL_THREADS + threadIdx.x;
    const uint64_t h = i < a.n_lines && a.state[i] != 0 ? 1 : 0;
    uint64_t total;
    const uint64_t r = a.tile_count[t] + block_exclusive_scan<COL_THREADS, uint64_t>(h, scratch, total);
    if (h) {
      a.o_start[r] = a.start[i];
      a.o_end[r] = a.end[i];
      a.o_score[r] = a.score[i];
      a.o_name_off[r] = a.name_off[i];
      a.o_name_len[r] = a.name_len[i];
    }
  }
}

// ---- chromosome ranks ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool bdg_name_head(const BdgNameArgs &a, uint64_t r) {
  if (r == 0) return true;
  const uint32_t n = a.name_len[r];
  if (n != a.name_len[r - 1]) return true;
  const uint8_t *x = a.text + a.name_off[r], *y = a.text + a.name_off[r - 1];
  for (uint32_t k = 0; k < n; k++)
    if (x[k] != y[k]) return true;
  return false;
}

__global__ void __launch_bounds__(COL_THREADS) bdg_name_heads_kernel(BdgNameArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t r = (uint64_t)t * COL_THREADS + threadIdx.x;
    const uint8_t f = r < a.n && bdg_name_head(a, r) ? 1 : 0;
    if (r < a.n) a.flag[r] = f;
    uint64_t total;
    block_exclusive_scan<COL_THREADS, uint64_t>((uint64_t)f, scratch, total);
    if (threadIdx.x == 0) a.tile_count[t] = total;
  }
}

// first row, name offset and name length of every name segment
__global__ void __launch_bounds__(COL_THREADS) bdg_name_segments_kernel(BdgNameArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t r = (uint64_t)t * COL_THREADS + threadIdx.x;
    const uint64_t f = r < a.n ? a.flag[r] : 0;
    uint64_t total;
    const uint64_t s = a.tile_count[t] + block_exclusive_scan<COL_THREADS, uint64_t>(f, scratch, total);
    if (f) {
      a.seg_name_off[s] = a.name_off[r];
      a.seg_name_len[s] = a.name_len[r];
    }
  }
}

__global__ void __launch_bounds__(COL_THREADS) bdg_assign_chrom_kernel(BdgNameArgs a) {
  __shared__ uint64_t scratch[COL_THREADS / 32 + 1];
  for (uint32_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint64_t r = (uint64_t)t * COL_THREADS + threadIdx.x;
    const uint64_t f = r < a.n ? a.flag[r] : 0;
    uint64_t total;
    const uint64_t inc = a.tile_count[t] + block_exclusive_scan<COL_THREADS, uint64_t>(f, scratch, total) + f;  // segments opened up to and including r
    if (r < a.n) a.chrom[r] = a.seg_rank[inc - 1];
  }
}

__global__ void __launch_bounds__(256) bdg_gather_rows_kernel(const uint32_t *perm, uint64_t n, const uint32_t *chrom, const long long *start, const long long *end,
                                                             const double *score, uint32_t *o_chrom, long long *o_start, long long *o_end, double *o_score) {
  for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull) {
    const uint32_t p = perm[i];
    o_chrom[i] = chrom[p];
    o_start[i] = start[p];
    o_end[i] = end[p];
    o_score[i] = score[p];
  }
}

}  // namespace bnd
