// bam_write.cuh — the BAM-writing consumers of the read filter on the device: `split`, `split-exogenous`, `modify`
// (bamnado/src/bam_splitter.rs:54-120, spike_in_analysis.rs:233-400, bam_modifier.rs:55-230).
//
// The reference decodes every record on one core, asks BamReadFilter::is_valid, and hands the survivors to a noodles BAM
// writer (zlib DEFLATE on the host).  Here the records are already in HBM after K1/K2 (inflate + record index):
//   bamw_classify   one thread per record: the same filter code as the pile-up (pileup.cuh) picks the output class of the
//                   record (or none), the Tn5 rule of `modify` decides whether a shifted record is still written, counters of
//                   SplitStats; bytes per class and tile of 256 records
//   bamw_scan       exclusive scan of the tile sums per class, continued from the previous segment (class streams are
//                   appended to in file order)
//   bamw_gather     records copied into their class stream (warp-wide word copies), Tn5 fields patched in place
//                   (pos, bin, mate pos, tlen), first record start of every 65 280-byte block noted for the compressor
//   bamw_deflate    one CTA per BGZF block: DEFLATE with dynamic Huffman codes built per block (RFC 1951 3.2.7).  The LZ77
//                   parse is record-parallel: BAM records repeat the records before them field by field, so every record is
//                   parsed by its own thread against the same field (fixed part + name / CIGAR .. qualities / tags) of the
//                   last eight records and against lag 1, with no hash table and no serial walk; a block scan of the per-record bit counts gives every thread
//                   its bit offset.  gzip header with the BC field, CRC-32 (warp-interleaved tables of the inflate kernel)
//                   and ISIZE are written by the same kernel; incompressible blocks fall back to a stored block.
// All of it is byte / integer work bound by shared-memory and L2 latency; no tensor-core shape.
#pragma once
#include "inflate.cuh"
#include "pileup.cuh"

namespace bnd {

constexpr int BAMW_THREADS = 256;
constexpr uint32_t BAMW_MAX_UNITS = 2048;  // records of one block (>= 36 bytes each) + pieces of long records
constexpr uint32_t BAMW_PIECE = 1024;      // longer records are parsed in pieces of this many bytes
constexpr uint32_t BAMW_NONE = 0xffffffffu;
constexpr uint32_t BAMW_MAX_RECS = 1824;    // >= BGZF_RAW / 36 + 2
constexpr uint32_t BAMW_NOREC = 0xffffu;
constexpr int BAMW_BACK = 8;                // previous records whose fields are tried as match sources

// ---- record classes ----------------------------------------------------------------------------------------------------------
// UCSC binning scheme (SAM spec 5.3), the `bin` field a BAM encoder writes for [beg, end)
__device__ __forceinline__ uint32_t bam_reg2bin(int64_t beg, int64_t end) {
  --end;
  if (beg >> 14 == end >> 14) return (uint32_t)(4681 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (uint32_t)(585 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (uint32_t)(73 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (uint32_t)(9 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (uint32_t)(1 + (beg >> 26));
  return 0;
}

struct Tn5Patch {
  int32_t pos0, next_pos0, tlen;
  uint32_t bin;
};

// BamModifier::run with --tn5-shift (bam_modifier.rs:103-219), shift values [4, -5, 5, -4]: forward reads start 4 bases
// later, reverse reads end 5 bases earlier, |tlen| shrinks by 9, a reverse read's mate start moves by +4.
// 1 = written with the patch, 0 = not written (not a proper pair / shifted out of the contig), -1 = the reference's error.
__device__ __forceinline__ int tn5_adjust(const RecView &v, uint64_t chrom_len, uint64_t span, Tn5Patch &p) {
  if (!(v.flag & 0x2)) return 0;
  const bool rev = (v.flag & 0x10) != 0;
  int64_t start = v.pos0 < 0 ? 0 : (int64_t)v.pos0;
  int64_t end = start + (int64_t)v.l_seq;
  if (rev) end -= 5;
  else start += 4;
  if (!(start >= 0 && end <= (int64_t)chrom_len)) return 0;
  if (start > (int64_t)chrom_len) return 0;
  if (v.next_pos0 < 0) return -1;  // "Missing mate alignment start"
  p.pos0 = (int32_t)start;
  p.tlen = v.tlen < 0 ? (v.tlen > INT32_MAX - 9 ? INT32_MAX : v.tlen + 9) : v.tlen - 9;
  p.next_pos0 = v.next_pos0;
  if (rev) {
    const int64_t adj = (int64_t)v.next_pos0 + 4;
    if (adj >= 0 && adj <= (int64_t)chrom_len) p.next_pos0 = (int32_t)adj;
  }
  p.bin = bam_reg2bin(start, start + (int64_t)span);
  return 1;
}

// class of a record (0xff = not written); one counter bump per record like the reference's branches
__device__ __forceinline__ uint32_t bamw_class_of(const BamwArgs &a, const RecView &v, unsigned int *cnt, Tn5Patch &patch, bool &patched) {
  patched = false;
  const uint32_t flag = v.flag;
  if (a.mode == BAMW_SPLIT_EXOGENOUS) {  // BamSplitter::split (spike_in_analysis.rs:322-380): every record of the file, in file order
    atomicAdd(&cnt[BW_N_TOTAL], 1u);
    if (flag & 0x4) return atomicAdd(&cnt[BW_N_UNMAPPED], 1u), 3u;
    if (flag & 0x200) return atomicAdd(&cnt[BW_N_QCFAIL], 1u), 3u;
    if (flag & 0x400) return atomicAdd(&cnt[BW_N_DUPLICATE], 1u), 3u;
    if (flag & 0x100) return atomicAdd(&cnt[BW_N_SECONDARY], 1u), 3u;
    int fail = -1;
    const int valid = filter_is_valid(a.f, v, fail);
    if (valid < 0 || (valid == 1 && v.ref >= a.n_ref)) return atomicAdd(&cnt[BW_N_ERRORS], 1u), 0xffu;  // `?` ends the reference's run
    if (valid == 0) return atomicAdd(&cnt[BW_N_FILTERED], 1u), 0xffu;
    if (a.ref_is_exo[v.ref]) return atomicAdd(&cnt[BW_N_EXOGENOUS], 1u), 1u;
    return atomicAdd(&cnt[BW_N_ENDOGENOUS], 1u), 0u;
  }
  // split / modify: what the per-contig queries return (noodles `intersects` against [1, length]), then the filter
  if (v.ref < 0 || v.ref >= a.n_ref || v.pos0 < 0) return 0xffu;
  if (32ull + v.lrn + 4ull * v.n_cigar > v.bs) return 0xffu;
  const uint64_t span = rec_span(v);
  if (span == 0) return 0xffu;
  const uint64_t len = a.ref_len[v.ref];
  if ((uint64_t)v.pos0 + 1 > len) return 0xffu;
  atomicAdd(&cnt[BW_N_TOTAL], 1u);
  int fail = -1;
  const int valid = filter_is_valid(a.f, v, fail);
  if (valid < 0) return atomicAdd(&cnt[BW_N_ERRORS], 1u), 0xffu;
  if (valid == 0) return atomicAdd(&cnt[BW_N_FILTERED], 1u), 0xffu;
  if (a.mode == BAMW_MODIFY && a.tn5_shift) {
    const int r = tn5_adjust(v, len, span, patch);
    if (r < 0) return atomicAdd(&cnt[BW_N_ERRORS], 1u), 0xffu;
    if (r == 0) return atomicAdd(&cnt[BW_N_DROPPED_BY_SHIFT], 1u), 0xffu;
    patched = true;
  }
  atomicAdd(&cnt[BW_N_WRITTEN], 1u);
  return 0u;
}

__global__ void __launch_bounds__(BAMW_TILE) bamw_classify_kernel(BamwArgs a) {
  __shared__ unsigned int s_cnt[BW_N_COUNTERS];
  __shared__ unsigned long long s_sum[BAMW_MAX_CLASSES];
  const int tid = (int)threadIdx.x;
  const uint64_t n_rec = a.st->n_records;
  const uint64_t n_tiles = (n_rec + BAMW_TILE - 1) / BAMW_TILE;
  if (tid < BW_N_COUNTERS) s_cnt[tid] = 0;
  for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    if (tid < BAMW_MAX_CLASSES) s_sum[tid] = 0;
    __syncthreads();
    const uint64_t i = t * BAMW_TILE + (uint64_t)tid;
    if (i < n_rec) {
      RecView v;
      rec_load(a.buf + a.rec_off[i], v);
      Tn5Patch patch;
      bool patched;
      const uint32_t c = bamw_class_of(a, v, s_cnt, patch, patched);
      a.rec_cls[i] = (uint8_t)c;
      if (c < (uint32_t)BAMW_MAX_CLASSES) atomicAdd(&s_sum[c], (unsigned long long)v.bs + 4ull);
    }
    __syncthreads();
    if (tid < BAMW_MAX_CLASSES) a.tile_sum[t * BAMW_MAX_CLASSES + tid] = s_sum[tid];
    __syncthreads();
  }
  if (tid < BW_N_COUNTERS && s_cnt[tid]) atomicAdd(&a.counts[tid], (unsigned long long)s_cnt[tid]);
}

// one CTA: tile sums -> stream offsets, per class, continued from class_len
__global__ void __launch_bounds__(BAMW_THREADS) bamw_scan_kernel(BamwArgs a) {
  __shared__ unsigned long long scratch[BAMW_THREADS / 32 + 1];
  const uint64_t n_rec = a.st->n_records;
  const uint64_t n_tiles = (n_rec + BAMW_TILE - 1) / BAMW_TILE;
  for (int c = 0; c < a.n_classes; c++) {
    unsigned long long running = a.class_len[c];
    __syncthreads();
    for (uint64_t t0 = 0; t0 < n_tiles; t0 += BAMW_THREADS) {
      const uint64_t t = t0 + threadIdx.x;
      unsigned long long v = t < n_tiles ? a.tile_sum[t * BAMW_MAX_CLASSES + c] : 0ull, total;
      const unsigned long long ex = block_exclusive_scan<BAMW_THREADS, unsigned long long>(v, scratch, total);
      if (t < n_tiles) a.tile_sum[t * BAMW_MAX_CLASSES + c] = running + ex;
      running += total;
    }
    if (threadIdx.x == 0) a.class_len[c] = running;
  }
}

// n bytes from src to dst by one warp; both unaligned.  Reads whole aligned words around [src, src + n): the segment buffers
// carry 64 readable bytes behind their data.
__device__ __forceinline__ void warp_copy_bytes(uint8_t *dst, const uint8_t *src, uint32_t n, int lane) {
  const uint32_t lead = tmin<uint32_t>((uint32_t)((4 - ((uintptr_t)dst & 3)) & 3), n);
  for (uint32_t i = (uint32_t)lane; i < lead; i += 32) dst[i] = src[i];
  const uint8_t *s2 = src + lead;
  const uint32_t nw = (n - lead) / 4;
  const uint32_t *sw = reinterpret_cast<const uint32_t *>((uintptr_t)s2 & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)((uintptr_t)s2 & 3) * 8;
  uint32_t *dw = reinterpret_cast<uint32_t *>(dst + lead);
  for (uint32_t j = (uint32_t)lane; j < nw; j += 32) {
    const uint32_t w0 = sw[j];
    dw[j] = sh ? __funnelshift_r(w0, sw[j + 1], sh) : w0;
  }
  for (uint32_t i = lead + nw * 4 + (uint32_t)lane; i < n; i += 32) dst[i] = src[i];
}

__device__ __forceinline__ void st_u32_bytes(uint8_t *p, uint32_t v) {
  p[0] = (uint8_t)v, p[1] = (uint8_t)(v >> 8), p[2] = (uint8_t)(v >> 16), p[3] = (uint8_t)(v >> 24);
}

__global__ void __launch_bounds__(BAMW_TILE) bamw_gather_kernel(BamwArgs a) {
  __shared__ unsigned long long scratch[BAMW_TILE / 32 + 1];
  const int tid = (int)threadIdx.x, lane = tid & 31;
  const uint64_t n_rec = a.st->n_records;
  const uint64_t n_tiles = (n_rec + BAMW_TILE - 1) / BAMW_TILE;
  for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const uint64_t i = t * BAMW_TILE + (uint64_t)tid;
    uint32_t cls = 0xffu, size = 0;
    const uint8_t *src = nullptr;
    if (i < n_rec) {
      cls = a.rec_cls[i];
      src = a.buf + a.rec_off[i];
      if (cls < (uint32_t)BAMW_MAX_CLASSES) size = ld_u32_unaligned(src) + 4u;
    }
    unsigned long long dst_off = 0;
    for (int c = 0; c < a.n_classes; c++) {
      unsigned long long v = cls == (uint32_t)c ? (unsigned long long)size : 0ull, total;
      const unsigned long long ex = block_exclusive_scan<BAMW_TILE, unsigned long long>(v, scratch, total);
      if (cls == (uint32_t)c) dst_off = a.tile_sum[t * BAMW_MAX_CLASSES + c] + ex;
    }
    // the warp copies its 32 records one after the other
    for (int j = 0; j < 32; j++) {
      const uint32_t n = __shfl_sync(0xffffffffu, size, j);
      if (n == 0) continue;
      const uint32_t cj = __shfl_sync(0xffffffffu, cls, j);
      const unsigned long long oj = __shfl_sync(0xffffffffu, dst_off, j);
      const uint8_t *sj = reinterpret_cast<const uint8_t *>(__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)src, j));
      warp_copy_bytes(a.stream[cj] + oj, sj, n, lane);
    }
    __syncwarp();
    if (size) {
      uint8_t *dst = a.stream[cls] + dst_off;
      if (a.mode == BAMW_MODIFY && a.tn5_shift) {  // the same decision as in bamw_classify: this record is written shifted
        RecView v;
        rec_load(src, v);
        Tn5Patch p;
        if (tn5_adjust(v, a.ref_len[v.ref], rec_span(v), p) == 1) {
          st_u32_bytes(dst + 8, (uint32_t)p.pos0);
          dst[14] = (uint8_t)p.bin, dst[15] = (uint8_t)(p.bin >> 8);
          st_u32_bytes(dst + 28, (uint32_t)p.next_pos0);
          st_u32_bytes(dst + 32, (uint32_t)p.tlen);
        }
      }
      // the record behind this one is the first to start in its block when this one ends in (or at the end of) an earlier block
      const unsigned long long nx = dst_off + size;
      if (nx / BGZF_RAW > dst_off / BGZF_RAW) a.first_rec[cls][nx / BGZF_RAW] = (uint32_t)(nx % BGZF_RAW);
    }
    __syncthreads();
  }
}

// ---- DEFLATE with dynamic Huffman codes, one CTA per BGZF block --------------------------------------------------------------
__device__ __forceinline__ void dfl_len_code(uint32_t len, uint32_t &sym, uint32_t &eb, uint32_t &ev) {
  if (len == 258) {
    sym = 285, eb = 0, ev = 0;
    return;
  }
  const uint32_t l = len - 3;
  if (l < 8) {
    sym = 257 + l, eb = 0, ev = 0;
    return;
  }
  const uint32_t k = 31u - (uint32_t)__clz((int)l);  // 3..7
  eb = k - 2;
  sym = 257 + 4 * (k - 1) + ((l - (1u << k)) >> eb);
  ev = l & ((1u << eb) - 1u);
}
__device__ __forceinline__ void dfl_dist_code(uint32_t dist, uint32_t &sym, uint32_t &eb, uint32_t &ev) {
  const uint32_t d = dist - 1;
  if (d < 4) {
    sym = d, eb = 0, ev = 0;
    return;
  }
  const uint32_t k = 31u - (uint32_t)__clz((int)d);  // 2..14
  eb = k - 1;
  sym = 2 * k + ((d >> (k - 1)) & 1u);
  ev = d & ((1u << eb) - 1u);
}

// Bits appended LSB-first to a zeroed image of 32-bit words.  Words a writer fills completely are stored plainly; its first and
// last word may be shared with the neighbouring writers and are OR-ed in.
struct BitWriter {
  uint32_t *img;
  uint32_t word, nacc;
  uint64_t acc;
  bool first;
  __device__ __forceinline__ void start(uint32_t *image, uint32_t bitpos) {
    img = image, word = bitpos >> 5, nacc = bitpos & 31u, acc = 0, first = true;
  }
  __device__ __forceinline__ void put(uint32_t v, uint32_t n) {  // n <= 32, v < 2^n
    acc |= (uint64_t)v << nacc;
    nacc += n;
    if (nacc >= 32) {
      const uint32_t lo = (uint32_t)acc;
      if (first) atomicOr(img + word, lo), first = false;
      else img[word] = lo;
      word++;
      acc >>= 32;
      nacc -= 32;
    }
  }
  __device__ __forceinline__ void finish() {
    if (nacc) atomicOr(img + word, (uint32_t)acc);
  }
};

struct DflShared {
  uint32_t hist[288 + 32];     // literal/length and distance frequencies
  uint32_t sorted_freq[288];   // Moffat-Katajainen work array (literal/length), then scratch
  uint32_t sorted_freq_d[32];
  uint16_t sorted_sym[288], sorted_sym_d[32];
  uint16_t lit_code[288], dist_code[32];
  uint8_t lit_len[288], dist_len[32];
  uint16_t unit_start[BAMW_MAX_UNITS + 1];
  uint16_t unit_rec[BAMW_MAX_UNITS];        // record the unit is a piece of (BAMW_NOREC: the tail of a record that started in an earlier block)
  // per record of the block: where it starts, where its CIGAR starts, where its tags start (BAMW_NOREC: not known / outside the block)
  uint16_t rec_start[BAMW_MAX_RECS], rec_mid[BAMW_MAX_RECS], rec_aux[BAMW_MAX_RECS];
  uint8_t rle_sym[320], rle_ext[320];
  uint16_t cl_code[19];
  uint8_t cl_len[19];
  uint32_t n_units, n_lit, n_dist, n_rle, hlit, hdist, hclen, header_bits, crc, stored;
  unsigned long long scratch[BAMW_THREADS / 32 + 1];
};

// A[0..n) ascending frequencies -> code lengths in place (Moffat & Katajainen, "In-place calculation of minimum-redundancy
// codes"), then limited to max_len by moving the excess to shorter codes (Kraft sum kept at 1); len[sym] for the n used symbols.
__device__ inline void dfl_code_lengths(uint32_t *A, const uint16_t *sym, int n, int max_len, uint8_t *len) {
  if (n == 0) return;
  if (n == 1) {
    len[sym[0]] = 1;
    return;
  }
  A[0] += A[1];
  int root = 0, leaf = 2, next;
  for (next = 1; next < n - 1; next++) {
    if (leaf >= n || A[root] < A[leaf]) {
      A[next] = A[root];
      A[root++] = (uint32_t)next;
    } else {
      A[next] = A[leaf++];
    }
    if (leaf >= n || (root < next && A[root] < A[leaf])) {
      A[next] += A[root];
      A[root++] = (uint32_t)next;
    } else {
      A[next] += A[leaf++];
    }
  }
  A[n - 2] = 0;
  for (next = n - 3; next >= 0; next--) A[next] = A[A[next]] + 1;
  int avbl = 1, used = 0, dpth = 0;
  root = n - 2;
  next = n - 1;
  while (avbl > 0) {
    while (root >= 0 && (int)A[root] == dpth) {
      used++;
      root--;
    }
    while (avbl > used) {
      A[next--] = (uint32_t)dpth;
      avbl--;
    }
    avbl = 2 * used;
    dpth++;
    used = 0;
  }
  constexpr int MAXD = 48;
  int num[MAXD + 1];
  for (int i = 0; i <= MAXD; i++) num[i] = 0;
  for (int i = 0; i < n; i++) num[A[i] < (uint32_t)MAXD ? A[i] : MAXD]++;
  for (int i = max_len + 1; i <= MAXD; i++) num[max_len] += num[i];
  uint32_t total = 0;
  for (int i = max_len; i > 0; i--) total += (uint32_t)num[i] << (max_len - i);
  while (total > (1u << max_len)) {
    num[max_len]--;
    for (int i = max_len - 1; i > 0; i--)
      if (num[i]) {
        num[i]--;
        num[i + 1] += 2;
        break;
      }
    total--;
  }
  int j = n;  // sym[] ascends by frequency: the most frequent symbols take the shortest codes
  for (int i = 1; i <= max_len; i++)
    for (int l = num[i]; l > 0; l--) len[sym[--j]] = (uint8_t)i;
}

// canonical codes (RFC 1951 3.2.2), bit-reversed for the LSB-first stream
__device__ inline void dfl_assign_codes(const uint8_t *len, int n, int max_len, uint16_t *code) {
  uint32_t count[16], next[16];
  for (int i = 0; i < 16; i++) count[i] = 0;
  for (int i = 0; i < n; i++) count[len[i]]++;
  count[0] = 0;
  uint32_t c = 0;
  next[0] = 0;
  for (int l = 1; l <= max_len; l++) {
    c = (c + count[l - 1]) << 1;
    next[l] = c;
  }
  for (int i = 0; i < n; i++) {
    const uint32_t l = len[i];
    code[i] = l ? (uint16_t)(__brev(next[l]++) >> (32 - l)) : (uint16_t)0;
  }
}

// symbols with a non-zero frequency ranked by (frequency, symbol): the whole CTA takes part
__device__ __forceinline__ void dfl_rank_sort(const uint32_t *freq, int n, uint32_t *out_freq, uint16_t *out_sym, uint32_t *n_used) {
  for (int i = (int)threadIdx.x; i < n; i += BAMW_THREADS) {
    const uint32_t f = freq[i];
    if (!f) continue;
    const uint32_t key = (f << 9) | (uint32_t)i;
    uint32_t rank = 0;
    for (int j = 0; j < n; j++) {
      const uint32_t g = freq[j];
      rank += g && ((g << 9) | (uint32_t)j) < key;
    }
    out_freq[rank] = f;
    out_sym[rank] = (uint16_t)i;
    atomicAdd(n_used, 1u);
  }
}

__device__ __forceinline__ uint32_t dfl_token_bits(const DflShared &S, uint32_t tok) {
  if (!(tok >> 31)) return S.lit_len[tok];
  uint32_t ls, leb, lev, ds, deb, dev;
  dfl_len_code(((tok >> 16) & 0xffu) + 3u, ls, leb, lev);
  dfl_dist_code((tok & 0xffffu) + 1u, ds, deb, dev);
  return S.lit_len[ls] + leb + S.dist_len[ds] + deb;
}

// Dynamic shared memory: the block's raw bytes (BGZF_RAW + 16).
__global__ void __launch_bounds__(BAMW_THREADS) bamw_deflate_kernel(BamwDeflateArgs a) {
  BND_DYN_SMEM(dyn);
  uint8_t *raw = dyn;
  __shared__ DflShared S;
  const int tid = (int)threadIdx.x;
  uint32_t *tok = a.tokens + (size_t)blockIdx.x * BGZF_RAW;
  const uint8_t cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  __shared__ __align__(8) unsigned long long s_mbar;
  const saddr_t mbar = smem_addr(&s_mbar);
  if (tid == 0) mbar_init(mbar, 1);
  uint32_t phase = 0;
  __syncthreads();
  for (uint32_t b = blockIdx.x; b < a.n_blocks; b += gridDim.x) {
    const uint64_t base = (uint64_t)b * BGZF_RAW;
    const uint32_t n = (uint32_t)tmin<uint64_t>(BGZF_RAW, a.raw_len - base);
    uint32_t *img = reinterpret_cast<uint32_t *>(a.out + (size_t)b * BGZF_SLOT);
    // ---- 1. the block into shared memory: one bulk copy by the copy engine (the previous iteration ended with a barrier, nobody
    //         touches `raw` any more), the last bytes by hand; histograms cleared meanwhile ----
    {
      const uint32_t n16 = n & ~15u;
      if (tid == 0 && n16) bulk_copy_g2s(smem_addr(raw), a.raw + base, n16, mbar);
      for (uint32_t i = n16 + (uint32_t)tid; i < n; i += BAMW_THREADS) raw[i] = a.raw[base + i];
      for (int i = tid; i < 288 + 32; i += BAMW_THREADS) S.hist[i] = 0;
      for (int i = tid; i < 288; i += BAMW_THREADS) S.lit_len[i] = 0, S.lit_code[i] = 0;
      if (tid < 32) S.dist_len[tid] = 0, S.dist_code[tid] = 0;
      if (tid == 0) S.n_lit = S.n_dist = 0, S.stored = (uint32_t)a.store_only;
      if (n16) {
        mbar_wait(mbar, phase);
        phase ^= 1u;
      }
    }
    __syncthreads();
    // ---- 2. units (thread 0 follows the record chain of the block) and the CRC-32 (warp 1) ----
    if (tid == 0) {
      uint32_t k = 0;
      // [s, e) of record r in pieces
      auto add = [&](uint32_t s, uint32_t e, uint32_t r) {
        while (s < e && k < BAMW_MAX_UNITS - 1) {
          const uint32_t q = k == BAMW_MAX_UNITS - 2 ? e : tmin<uint32_t>(e, s + BAMW_PIECE);
          S.unit_start[k] = (uint16_t)s, S.unit_rec[k] = (uint16_t)r;
          k++;
          s = q;
        }
      };
      uint32_t first = a.first_rec ? a.first_rec[b] : BAMW_NONE;
      if (first >= n) first = n;
      add(0, first, BAMW_NOREC);
      // Records differ in name, CIGAR and read length, so one lag does not line every field up with an earlier record: the
      // fixed fields and the name repeat at the distance of the record starts, CIGAR / sequence / qualities at the distance of
      // the CIGAR starts, the tags at the distance of the aux starts.  The table below lets the parse try the last few records.
      uint32_t p = first, r = 0;
      while (p < n && r < BAMW_MAX_RECS) {
        uint32_t q = n, mid = BAMW_NOREC, aux = BAMW_NOREC;
        if (p + 4 <= n) {
          const uint32_t bs = (uint32_t)raw[p] | ((uint32_t)raw[p + 1] << 8) | ((uint32_t)raw[p + 2] << 16) | ((uint32_t)raw[p + 3] << 24);
          if (bs >= 32 && bs < REC_MAX_SIZE && (uint64_t)p + 4 + bs < (uint64_t)n) q = p + 4 + bs;
          if (bs >= 32 && p + 24 <= n) {
            const uint32_t lrn = raw[p + 12], nc = (uint32_t)raw[p + 16] | ((uint32_t)raw[p + 17] << 8);
            const uint32_t ls = (uint32_t)raw[p + 20] | ((uint32_t)raw[p + 21] << 8) | ((uint32_t)raw[p + 22] << 16) | ((uint32_t)raw[p + 23] << 24);
            const uint64_t m64 = (uint64_t)p + 36 + lrn, a64 = m64 + 4ull * nc + ((uint64_t)ls + 1) / 2 + ls;
            if (a64 <= (uint64_t)p + 4 + bs && m64 < n) mid = (uint32_t)m64, aux = a64 < n ? (uint32_t)a64 : BAMW_NOREC;
          }
        }
        S.rec_start[r] = (uint16_t)p, S.rec_mid[r] = (uint16_t)mid, S.rec_aux[r] = (uint16_t)aux;
        add(p, q, r);
        r++;
        p = q;
      }
      if (p < n) add(p, n, BAMW_NOREC);  // cannot happen (a record is at least 36 bytes); kept so that every byte belongs to a unit
      S.unit_start[k] = (uint16_t)n;
      S.n_units = k;
    }
    if (tid >= 32 && tid < 64) {
      Tile<32> tile;
      CrcStream cs;
      const uint32_t crc = cs.finish(tile, raw, n, a.crc);
      if (tile.lane == 0) S.crc = crc;
    }
    __syncthreads();
    const uint32_t U = S.n_units, K = (U + BAMW_THREADS - 1) / BAMW_THREADS;
    const uint32_t u0 = tmin<uint32_t>(U, (uint32_t)tid * K), u1 = tmin<uint32_t>(U, u0 + K);
    // ---- 3. greedy parse of every unit against the fields of the records before it and lag 1; tokens to scratch; frequencies ----
    if (!S.stored) {
      for (uint32_t u = u0; u < u1; u++) {
        const uint32_t s = S.unit_start[u], e = S.unit_start[u + 1];
        const uint32_t r = S.unit_rec[u];
        const uint32_t mid = r == BAMW_NOREC ? BAMW_NOREC : S.rec_mid[r], aux = r == BAMW_NOREC ? BAMW_NOREC : S.rec_aux[r];
        uint32_t p = s, k = 0;
        while (p < e) {
          const uint32_t maxm = tmin<uint32_t>(258u, e - p);
          uint32_t best = 0, bestd = 0;
          if (r != BAMW_NOREC) {
            // the same field of the last few records: 0 = fixed fields and name, 1 = CIGAR .. qualities, 2 = tags
            const int field = p < mid ? 0 : p < aux ? 1 : 2;
            const uint32_t here = field == 0 ? S.rec_start[r] : field == 1 ? mid : aux;
#pragma unroll 1
            for (int back = 1; back <= BAMW_BACK && (uint32_t)back <= r; back++) {
              const uint32_t there = field == 0 ? S.rec_start[r - back] : field == 1 ? S.rec_mid[r - back] : S.rec_aux[r - back];
              if (there == BAMW_NOREC) continue;
              const uint32_t lag = here - there;
              if (lag > 32768u || lag > p) continue;
              uint32_t m = 0;
              while (m < maxm && raw[p + m] == raw[p + m - lag]) m++;
              if (m >= 4 && m > best + (best ? 1u : 0u)) best = m, bestd = lag;  // a farther source pays a longer distance code
              if (best >= 32) break;  // long enough: the records further back are not asked
            }
          }
          if (p >= 1 && raw[p] == raw[p - 1]) {
            uint32_t m = 1;
            while (m < maxm && raw[p + m] == raw[p + m - 1]) m++;
            if (m >= 3 && m > best) best = m, bestd = 1;
          }
          uint32_t t;
          if (best) {
            uint32_t ls, leb, lev, ds, deb, dev;
            dfl_len_code(best, ls, leb, lev);
            dfl_dist_code(bestd, ds, deb, dev);
            atomicAdd(&S.hist[ls], 1u);
            atomicAdd(&S.hist[288 + ds], 1u);
            t = 0x80000000u | ((best - 3u) << 16) | (bestd - 1u);
            p += best;
          } else {
            t = raw[p];
            atomicAdd(&S.hist[t], 1u);
            p++;
          }
          tok[s + k++] = t;
        }
        // the token count of the unit is kept in the slot behind its last token when there is room, else it equals the byte count
        if (k < e - s) tok[s + k] = 0xffffffffu;
      }
      if (tid == 0) S.hist[256] = 1;
    }
    __syncthreads();
    // ---- 4. code lengths and codes: both alphabets ranked by the CTA, the trees built by one thread each ----
    if (!S.stored) {
      dfl_rank_sort(S.hist, 286, S.sorted_freq, S.sorted_sym, &S.n_lit);
      dfl_rank_sort(S.hist + 288, 30, S.sorted_freq_d, S.sorted_sym_d, &S.n_dist);
    }
    __syncthreads();
    if (!S.stored) {
      if (tid == 0) {
        dfl_code_lengths(S.sorted_freq, S.sorted_sym, (int)S.n_lit, 15, S.lit_len);
        dfl_assign_codes(S.lit_len, 286, 15, S.lit_code);
      }
      if (tid == 32) {
        if (S.n_dist == 0) S.dist_len[0] = 1;  // no match in the block: one unused 1-bit distance code keeps every inflater content
        else dfl_code_lengths(S.sorted_freq_d, S.sorted_sym_d, (int)S.n_dist, 15, S.dist_len);
        dfl_assign_codes(S.dist_len, 30, 15, S.dist_code);
      }
    }
    __syncthreads();
    // ---- 5. the block header: run-length coded code lengths and their own 7-bit code (thread 0); bits of every thread's units ----
    if (!S.stored && tid == 0) {
      int hlit = 286, hdist = 30;
      while (hlit > 257 && S.lit_len[hlit - 1] == 0) hlit--;
      while (hdist > 1 && S.dist_len[hdist - 1] == 0) hdist--;
      uint32_t clf[19];
      for (int i = 0; i < 19; i++) clf[i] = 0, S.cl_len[i] = 0;
      const int total = hlit + hdist;
      auto at = [&](int i) -> uint32_t { return i < hlit ? S.lit_len[i] : S.dist_len[i - hlit]; };
      uint32_t nr = 0;
      auto emit = [&](uint32_t sym, uint32_t ext) {
        S.rle_sym[nr] = (uint8_t)sym, S.rle_ext[nr] = (uint8_t)ext;
        nr++;
        clf[sym]++;
      };
      for (int i = 0; i < total;) {
        const uint32_t v = at(i);
        int run = 1;
        while (i + run < total && at(i + run) == v) run++;
        i += run;
        if (v == 0) {
          while (run >= 11) {
            const int r = run < 138 ? run : 138;
            emit(18, (uint32_t)(r - 11));
            run -= r;
          }
          if (run >= 3) emit(17, (uint32_t)(run - 3)), run = 0;
          while (run-- > 0) emit(0, 0);
        } else {
          emit(v, 0);
          run--;
          while (run >= 3) {
            const int r = run < 6 ? run : 6;
            emit(16, (uint32_t)(r - 3));
            run -= r;
          }
          while (run-- > 0) emit(v, 0);
        }
      }
      // the code-length alphabet: 19 symbols, at most 7 bits; sorted by insertion
      uint32_t f[19];
      uint16_t sy[19];
      int m = 0;
      for (int i = 0; i < 19; i++)
        if (clf[i]) {
          int j = m++;
          while (j > 0 && (f[j - 1] > clf[i])) f[j] = f[j - 1], sy[j] = sy[j - 1], j--;
          f[j] = clf[i], sy[j] = (uint16_t)i;
        }
      dfl_code_lengths(f, sy, m, 7, S.cl_len);
      if (m == 1) S.cl_len[sy[0] == 0 ? 1 : 0] = 1;  // the code-length code must be complete: a second 1-bit code that is never used
      dfl_assign_codes(S.cl_len, 19, 7, S.cl_code);
      int hclen = 19;
      while (hclen > 4 && S.cl_len[cl_order[hclen - 1]] == 0) hclen--;
      uint32_t hb = 3 + 5 + 5 + 4 + 3 * (uint32_t)hclen;
      for (uint32_t i = 0; i < nr; i++) {
        const uint32_t s = S.rle_sym[i];
        hb += S.cl_len[s] + (s == 16 ? 2u : s == 17 ? 3u : s == 18 ? 7u : 0u);
      }
      S.n_rle = nr, S.hlit = (uint32_t)hlit, S.hdist = (uint32_t)hdist, S.hclen = (uint32_t)hclen, S.header_bits = hb;
    }
    __syncthreads();
    uint32_t my_bits = 0;
    if (!S.stored)
      for (uint32_t u = u0; u < u1; u++) {
        const uint32_t s = S.unit_start[u], e = S.unit_start[u + 1];
        for (uint32_t k = 0; k < e - s; k++) {
          const uint32_t t = tok[s + k];
          if (t == 0xffffffffu) break;
          my_bits += dfl_token_bits(S, t);
        }
      }
    unsigned long long total_bits;
    const unsigned long long my_off = block_exclusive_scan<BAMW_THREADS, unsigned long long>((unsigned long long)my_bits, S.scratch, total_bits);
    const uint32_t data_bit0 = 18u * 8u + S.header_bits;
    const uint32_t dfl_bits = S.header_bits + (uint32_t)total_bits + S.lit_len[256];
    uint32_t cbytes = (dfl_bits + 7) / 8;
    const bool stored = S.stored || cbytes >= n + 5u || 18u + cbytes + 8u > BGZF_SLOT;
    if (stored) cbytes = n + 5u;
    const uint32_t total_bytes = 18u + cbytes + 8u;
    // ---- 6. the image: zeroed, then every thread writes the bits of its units; thread 0 adds headers, end of block and trailer ----
    for (uint32_t i = (uint32_t)tid; i < (total_bytes + 3) / 4; i += BAMW_THREADS) img[i] = 0;
    __syncthreads();
    if (tid == 0) {
      BitWriter w;
      w.start(img, 0);
      const uint8_t gz[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
      for (int i = 0; i < 16; i++) w.put(gz[i], 8);
      w.put(total_bytes - 1, 16);
      if (stored) {
        w.put(1, 8);  // BFINAL = 1, BTYPE = 00, padding to the byte
        w.put(n, 16);
        w.put(~n & 0xffffu, 16);
      } else {
        w.put(1, 1);
        w.put(2, 2);
        w.put(S.hlit - 257, 5);
        w.put(S.hdist - 1, 5);
        w.put(S.hclen - 4, 4);
        for (uint32_t i = 0; i < S.hclen; i++) w.put(S.cl_len[cl_order[i]], 3);
        for (uint32_t i = 0; i < S.n_rle; i++) {
          const uint32_t s = S.rle_sym[i];
          w.put(S.cl_code[s], S.cl_len[s]);
          if (s == 16) w.put(S.rle_ext[i], 2);
          else if (s == 17) w.put(S.rle_ext[i], 3);
          else if (s == 18) w.put(S.rle_ext[i], 7);
        }
      }
      w.finish();
      if (!stored) {
        BitWriter eob;
        eob.start(img, data_bit0 + (uint32_t)total_bits);
        eob.put(S.lit_code[256], S.lit_len[256]);
        eob.finish();
      }
      BitWriter tr;
      tr.start(img, (18u + cbytes) * 8u);
      tr.put(S.crc, 32);
      tr.put(n, 32);
      tr.finish();
      a.out_size[b] = total_bytes;
    }
    if (stored) {
      uint8_t *dst = a.out + (size_t)b * BGZF_SLOT + 23;  // bytes owned by this loop only (header and trailer words were OR-ed in)
      // the words that hold the first / last payload bytes are shared with header and trailer: those bytes go in with atomics
      for (uint32_t i = (uint32_t)tid; i < n; i += BAMW_THREADS) {
        const uint32_t byte = 23u + i;
        if (byte < 24u || byte >= ((18u + cbytes) & ~3u)) atomicOr(img + (byte >> 2), (uint32_t)raw[i] << ((byte & 3u) * 8u));
        else dst[i] = raw[i];
      }
    } else if (my_bits) {
      BitWriter w;
      w.start(img, data_bit0 + (uint32_t)my_off);
      for (uint32_t u = u0; u < u1; u++) {
        const uint32_t s = S.unit_start[u], e = S.unit_start[u + 1];
        for (uint32_t k = 0; k < e - s; k++) {
          const uint32_t t = tok[s + k];
          if (t == 0xffffffffu) break;
          if (!(t >> 31)) {
            w.put(S.lit_code[t], S.lit_len[t]);
          } else {
            uint32_t ls, leb, lev, ds, deb, dev;
            dfl_len_code(((t >> 16) & 0xffu) + 3u, ls, leb, lev);
            dfl_dist_code((t & 0xffffu) + 1u, ds, deb, dev);
            w.put(S.lit_code[ls] | (lev << S.lit_len[ls]), S.lit_len[ls] + leb);
            w.put(S.dist_code[ds] | (dev << S.dist_len[ds]), S.dist_len[ds] + deb);
          }
        }
      }
      w.finish();
    }
    __syncthreads();
  }
}

}  // namespace bnd
