// inflate.cuh — K1: per-BGZF-block DEFLATE inflate + CRC32 check, hand-written for sm_100a.
//
// Replaces noodles-bgzf block decoding (the reference calls it through
// bam::io::indexed_reader, bamnado/src/coverage_analysis.rs:456-466).  One tile of G lanes owns one BGZF
// block (<= 64 KiB in, <= 64 KiB out).  Lane 0 walks the Huffman stream through shared-memory lookup tables
// and emits a batch of tokens; the whole tile then executes the batch (literal stores, LZ77 copies), and at
// the end verifies ISIZE and the CRC32 with a lane-interleaved table CRC.  The work is byte/bit manipulation,
// HBM- and latency-bound; there is no GEMM shape in it, so no tensor-core path.
#pragma once
#include "common.cuh"
#include "kernel_args.h"

namespace bnd {


constexpr int LIT_BITS = 10;   // primary lookup width for literal/length codes
constexpr int DIST_BITS = 8;   // primary lookup width for distance codes
constexpr int NTOK = 32;       // tokens decoded per batch

// litlen LUT entry (u32): bits 0-3 code length, bits 4-5 kind (0 literal, 1 end-of-block, 2 length, 3 long/invalid),
// bits 8-11 extra-bit count, bits 16-31 value (literal byte or length base).
// dist LUT entry (u32): bits 0-3 code length, bit 4 long/invalid, bits 8-11 extra-bit count, bits 16-31 base.
constexpr uint32_t KIND_LIT = 0u << 4, KIND_EOB = 1u << 4, KIND_LEN = 2u << 4, KIND_LONG = 3u << 4;

struct HuffCanon {       // canonical description for the slow path (codes longer than the primary width)
  uint16_t count[16];    // number of codes of each length
  uint16_t sorted[288];  // symbols ordered by (length, symbol)
};

struct alignas(16) TileSmem {
  uint32_t lit_lut[1 << LIT_BITS];
  uint32_t dist_lut[1 << DIST_BITS];
  HuffCanon lit_canon;
  struct {
    uint16_t count[16];
    uint16_t sorted[32];
  } dist_canon;
  uint8_t lens[320];      // code lengths of the block being set up (litlen then dist)
  uint16_t codes[320];    // canonical code of each symbol
  uint32_t tok[NTOK];
  uint32_t scratch[4];
};


__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// ---- bit reader (used by lane 0 only) ----------------------------------------------------------------------
struct BitReader {
  const uint32_t *wp;   // next aligned word to fetch
  uint64_t buf;
  int cnt;
  uint32_t nextw;       // prefetched word
  const uint32_t *end;  // first word past the payload (for overrun detection)
  __device__ __forceinline__ void init(const uint8_t *p, const uint8_t *pend) {
    uintptr_t a = (uintptr_t)p;
    wp = (const uint32_t *)(a & ~(uintptr_t)3);
    int skip = (int)(a & 3) * 8;
    uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1);
    nextw = __ldg(wp + 2);
    wp += 3;
    buf = ((uint64_t)w1 << 32 | w0) >> skip;
    cnt = 64 - skip;
    end = (const uint32_t *)(((uintptr_t)pend + 3) & ~(uintptr_t)3);
  }
  __device__ __forceinline__ void ensure() {  // afterwards cnt > 32
    if (cnt <= 32) {
      buf |= (uint64_t)nextw << cnt;
      cnt += 32;
      nextw = __ldg(wp);
      wp++;
    }
  }
  __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)buf & ((1u << n) - 1u); }
  __device__ __forceinline__ void consume(int n) {
    buf >>= n;
    cnt -= n;
  }
  __device__ __forceinline__ uint32_t take(int n) {
    uint32_t v = peek(n);
    consume(n);
    return v;
  }
  // bytes consumed so far relative to the aligned origin are not tracked; overrun is detected from wp
  __device__ __forceinline__ bool overrun() const {
    // words fetched: wp; words still buffered: cnt bits in buf + 32 in nextw
    const uint8_t *consumed_to = (const uint8_t *)wp - 4 - (cnt >> 3);
    return consumed_to > (const uint8_t *)end;
  }
  __device__ __forceinline__ void align_byte() { consume(cnt & 7); }
  __device__ __forceinline__ const uint8_t *byte_ptr() const { return (const uint8_t *)wp - 4 - (cnt >> 3); }
};

__device__ __forceinline__ uint32_t bitrev(uint32_t code, int len) { return __brev(code) >> (32 - len); }

// Slow path: canonical decode bit by bit (RFC 1951 ordering), for codes longer than the primary LUT width.
__device__ __forceinline__ int canon_decode(BitReader &br, const uint16_t *count, const uint16_t *sorted, int &sym_out) {
  uint32_t bits = (uint32_t)br.buf;
  int code = 0, first = 0, index = 0;
  for (int len = 1; len <= 15; len++) {
    code |= (int)(bits & 1);
    bits >>= 1;
    int c = count[len];
    if (code - c < first) {
      sym_out = sorted[index + (code - first)];
      br.consume(len);
      return len;
    }
    index += c;
    first += c;
    first <<= 1;
    code <<= 1;
  }
  return 0;
}

// Build LUT + canonical tables for `n` symbols whose lengths are in S.lens[base .. base+n).
// kind 0: literal/length alphabet, 1: distance alphabet.  Returns false on an over-subscribed code.
template <int G>
__device__ bool build_tables(const Tile<G> &tile, TileSmem &S, int base, int n, int kind) {
  const int lane = tile.lane;
  uint16_t *count = kind == 0 ? S.lit_canon.count : S.dist_canon.count;
  uint16_t *sorted = kind == 0 ? S.lit_canon.sorted : S.dist_canon.sorted;
  uint32_t *lut = kind == 0 ? S.lit_lut : S.dist_lut;
  const int bits = kind == 0 ? LIT_BITS : DIST_BITS;
  const uint32_t invalid = kind == 0 ? (KIND_LONG | 0u) : (1u << 4);
  for (int i = lane; i < (1 << bits); i += G) lut[i] = invalid;  // length 0 => invalid (slow path rejects too)
  bool ok = true;
  if (lane == 0) {
    for (int l = 0; l < 16; l++) count[l] = 0;
    for (int s = 0; s < n; s++) count[S.lens[base + s]]++;
    count[0] = 0;
    // over-subscription check + first code / offset of each length
    uint16_t next_code[16], offs[16];
    int left = 1;
    uint32_t code = 0;
    uint16_t off = 0;
    for (int l = 1; l <= 15; l++) {
      left <<= 1;
      left -= count[l];
      if (left < 0) ok = false;
      code = (code + count[l - 1]) << 1;
      next_code[l] = (uint16_t)code;
      offs[l] = off;
      off += count[l];
    }
    if (ok)
      for (int s = 0; s < n; s++) {
        int l = S.lens[base + s];
        if (l) {
          S.codes[base + s] = next_code[l]++;
          sorted[offs[l]++] = (uint16_t)s;
        }
      }
  }
  ok = tile.shfl((int)ok, 0);
  tile.sync();
  if (!ok) return false;
  for (int s = lane; s < n; s += G) {
    int l = S.lens[base + s];
    if (l == 0) continue;
    uint32_t rev = bitrev(S.codes[base + s], l);
    if (l <= bits) {
      uint32_t e = invalid;
      if (kind == 0) {
        if (s < 256) e = KIND_LIT | ((uint32_t)s << 16) | (uint32_t)l;
        else if (s == 256) e = KIND_EOB | (uint32_t)l;
        else if (s < 286) e = KIND_LEN | ((uint32_t)c_len_extra[s - 257] << 8) | ((uint32_t)c_len_base[s - 257] << 16) | (uint32_t)l;
      } else {
        if (s < 30) e = (uint32_t)l | ((uint32_t)c_dist_extra[s] << 8) | ((uint32_t)c_dist_base[s] << 16);
      }
      for (uint32_t i = rev; i < (1u << bits); i += (1u << l)) lut[i] = e;
    } else {
      lut[rev & ((1u << bits) - 1)] = invalid;  // defer to the canonical walk
    }
  }
  tile.sync();
  return true;
}

// CRC-32 of out[0..n) computed by the tile; every lane returns the value.
template <int G>
__device__ uint32_t tile_crc32(const Tile<G> &tile, const uint8_t *out, uint32_t n, const CrcTables *T) {
  const int lane = tile.lane;
  uint32_t c = 0xffffffffu;
  // leading bytes up to 4-byte alignment, serial (every lane computes the same thing; cheap)
  uint32_t lead = (uint32_t)((4 - ((uintptr_t)out & 3)) & 3);
  if (lead > n) lead = n;
  for (uint32_t i = 0; i < lead; i++) c = T->t[(c ^ out[i]) & 0xff] ^ (c >> 8);
  const uint32_t *w = (const uint32_t *)(out + lead);
  uint32_t nwords = (n - lead) >> 2;
  uint32_t rounds = nwords / G;  // full rounds of G words
  uint32_t acc = 0;
  if (rounds) {
    for (uint32_t r = 0; r < rounds; r++) {
      uint32_t v = w[r * G + lane];
      if (r == 0 && lane == 0) v ^= c;
      acc = T->z[0][acc & 0xff] ^ T->z[1][(acc >> 8) & 0xff] ^ T->z[2][(acc >> 16) & 0xff] ^ T->z[3][acc >> 24] ^ v;
    }
    // lane l still has to advance by (G - l) words of zeros
    for (int k = 0; k < G - lane; k++) {
#pragma unroll
      for (int b = 0; b < 4; b++) acc = T->t[acc & 0xff] ^ (acc >> 8);
    }
#pragma unroll
    for (int d = G / 2; d >= 1; d >>= 1) acc ^= tile.shfl_xor(acc, d);
    c = acc;
  }
  // tail: remaining words and bytes, serial
  const uint8_t *tail = (const uint8_t *)(w + rounds * G);
  uint32_t ntail = n - lead - rounds * G * 4;
  for (uint32_t i = 0; i < ntail; i++) c = T->t[(c ^ tail[i]) & 0xff] ^ (c >> 8);
  return ~c;
}

// Inflate one BGZF block with a tile of G lanes.  Returns a status code (same on every lane).
template <int G>
__device__ int inflate_block(const Tile<G> &tile, TileSmem &S, const uint8_t *blk, uint32_t blk_size,
                             uint8_t *out, uint32_t isize_expected, const CrcTables *crcT, int verify_crc) {
  const int lane = tile.lane;
  if (blk_size < 26) return INF_BAD_HEADER;
  uint32_t xlen = blk[10] | ((uint32_t)blk[11] << 8);
  if (12 + xlen + 8 > blk_size) return INF_BAD_HEADER;
  const uint8_t *payload = blk + 12 + xlen;
  const uint8_t *pend = blk + blk_size - 8;
  uint32_t crc_expected = pend[0] | ((uint32_t)pend[1] << 8) | ((uint32_t)pend[2] << 16) | ((uint32_t)pend[3] << 24);
  uint32_t isize = pend[4] | ((uint32_t)pend[5] << 8) | ((uint32_t)pend[6] << 16) | ((uint32_t)pend[7] << 24);
  if (isize != isize_expected) return INF_SIZE_MISMATCH;

  BitReader br;
  br.init(payload, pend);  // every lane initialises; only lane 0 advances it
  uint32_t outpos = 0;
  int status = INF_OK;
  bool final_block = false;

  while (!final_block && status == INF_OK) {
    // ---- block header (lane 0), broadcast ----
    int btype = 0;
    int hlit = 0, hdist = 0;
    if (lane == 0) {
      br.ensure();
      final_block = br.take(1);
      btype = (int)br.take(2);
    }
    btype = tile.shfl(btype, 0);
    final_block = tile.shfl((int)final_block, 0);
    if (btype == 3) {
      status = INF_BAD_BTYPE;
      break;
    }
    if (btype == 0) {
      // stored block: LEN/NLEN then raw bytes
      uint32_t len = 0;
      const uint8_t *src = nullptr;
      int st = INF_OK;
      if (lane == 0) {
        br.align_byte();
        br.ensure();
        uint32_t l = br.take(16);
        br.ensure();
        uint32_t nl = br.take(16);
        if ((l ^ nl) != 0xffffu) st = INF_BAD_STORED;
        len = l;
        src = br.byte_ptr();
        if (src + len > pend) st = INF_INPUT_OVERRUN;
        if (outpos + len > isize) st = INF_OUTPUT_OVERRUN;
      }
      st = tile.shfl(st, 0);
      if (st != INF_OK) {
        status = st;
        break;
      }
      len = tile.shfl(len, 0);
      src = (const uint8_t *)tile.shfl((unsigned long long)(uintptr_t)src, 0);
      for (uint32_t i = lane; i < len; i += G) out[outpos + i] = src[i];
      outpos += len;
      if (lane == 0) br.init(src + len, pend);
      tile.sync();
      continue;
    }
    if (btype == 1) {
      // fixed Huffman code lengths
      for (int s = lane; s < 288; s += G) S.lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
      for (int s = lane; s < 30; s += G) S.lens[288 + s] = 5;
      hlit = 288;
      hdist = 30;
      tile.sync();
    } else {
      // dynamic: read code-length code, then HLIT + HDIST lengths (lane 0)
      int st = INF_OK;
      if (lane == 0) {
        br.ensure();
        hlit = (int)br.take(5) + 257;
        hdist = (int)br.take(5) + 1;
        int hclen = (int)br.take(4) + 4;
        if (hlit > 286 || hdist > 30) st = INF_BAD_CODELEN;
        uint8_t cl[19];
#pragma unroll
        for (int i = 0; i < 19; i++) cl[i] = 0;
        for (int i = 0; i < hclen; i++) {
          br.ensure();
          cl[c_cl_order[i]] = (uint8_t)br.take(3);
        }
        // 7-bit LUT for the code-length code, kept in S.tok/S.codes scratch: use S.codes as 128 x u16
        uint16_t *cl_lut = S.codes;  // reused before codes are assigned
        for (int i = 0; i < 128; i++) cl_lut[i] = 0;
        {
          int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          for (int i = 0; i < 19; i++) cnt[cl[i]]++;
          cnt[0] = 0;
          int left = 1;
          uint32_t code = 0, next[8];
          next[0] = 0;
          for (int l = 1; l <= 7; l++) {
            left <<= 1;
            left -= cnt[l];
            if (left < 0) st = INF_BAD_CODELEN;
            code = (code + cnt[l - 1]) << 1;
            next[l] = code;
          }
          if (st == INF_OK)
            for (int s = 0; s < 19; s++) {
              int l = cl[s];
              if (!l) continue;
              uint32_t rev = bitrev(next[l]++, l);
              for (uint32_t i = rev; i < 128; i += (1u << l)) cl_lut[i] = (uint16_t)((s << 4) | l);
            }
        }
        int n = 0, total = hlit + hdist;
        int prev = 0;
        while (st == INF_OK && n < total) {
          br.ensure();
          uint32_t e = cl_lut[br.peek(7)];
          int l = e & 15, s = e >> 4;
          if (l == 0) {
            st = INF_BAD_CODELEN;
            break;
          }
          br.consume(l);
          if (s < 16) {
            S.lens[n++] = (uint8_t)s;
            prev = s;
          } else {
            int rep, val;
            if (s == 16) {
              if (n == 0) {
                st = INF_BAD_CODELEN;
                break;
              }
              rep = 3 + (int)br.take(2);
              val = prev;
            } else if (s == 17) {
              rep = 3 + (int)br.take(3);
              val = 0;
              prev = 0;
            } else {
              rep = 11 + (int)br.take(7);
              val = 0;
              prev = 0;
            }
            if (n + rep > total) {
              st = INF_BAD_CODELEN;
              break;
            }
            for (int k = 0; k < rep; k++) S.lens[n++] = (uint8_t)val;
          }
        }
        if (st == INF_OK && S.lens[256] == 0) st = INF_BAD_CODELEN;  // no end-of-block code
        if (st == INF_OK && hlit < 288) {
          // move the distance lengths to their fixed slot at 288
          for (int k = hdist - 1; k >= 0; k--) S.lens[288 + k] = S.lens[hlit + k];
          for (int k = hlit; k < 288; k++) S.lens[k] = 0;
          for (int k = hdist; k < 30; k++) S.lens[288 + k] = 0;
        }
      }
      st = tile.shfl(st, 0);
      if (st != INF_OK) {
        status = st;
        break;
      }
      hlit = tile.shfl(hlit, 0);
      hdist = tile.shfl(hdist, 0);
      tile.sync();
    }
    if (!build_tables<G>(tile, S, 0, 288, 0) || !build_tables<G>(tile, S, 288, 30, 1)) {
      status = INF_BAD_CODELEN;
      break;
    }

    // ---- symbol loop: lane 0 decodes a batch, the tile executes it ----
    bool eob = false;
    while (!eob && status == INF_OK) {
      int ntok = 0;
      int st = INF_OK;
      if (lane == 0) {
        while (ntok < NTOK) {
          br.ensure();
          uint32_t e = S.lit_lut[br.peek(LIT_BITS)];
          uint32_t kind = e & (3u << 4);
          if (kind == KIND_LONG) {
            int sym;
            if (!canon_decode(br, S.lit_canon.count, S.lit_canon.sorted, sym)) {
              st = INF_BAD_SYMBOL;
              break;
            }
            if (sym < 256) e = KIND_LIT | ((uint32_t)sym << 16);
            else if (sym == 256) e = KIND_EOB;
            else if (sym < 286) e = KIND_LEN | ((uint32_t)c_len_extra[sym - 257] << 8) | ((uint32_t)c_len_base[sym - 257] << 16);
            else {
              st = INF_BAD_SYMBOL;
              break;
            }
            kind = e & (3u << 4);
          } else {
            br.consume((int)(e & 15));
          }
          if (kind == KIND_LIT) {
            S.tok[ntok++] = e >> 16;
            continue;
          }
          if (kind == KIND_EOB) {
            eob = true;
            break;
          }
          uint32_t length = (e >> 16) + br.take((int)((e >> 8) & 15));
          br.ensure();
          uint32_t d = S.dist_lut[br.peek(DIST_BITS)];
          if (d & (1u << 4)) {
            int sym;
            if (!canon_decode(br, S.dist_canon.count, S.dist_canon.sorted, sym) || sym >= 30) {
              st = INF_BAD_DISTANCE;
              break;
            }
            d = ((uint32_t)c_dist_extra[sym] << 8) | ((uint32_t)c_dist_base[sym] << 16);
          } else {
            br.consume((int)(d & 15));
          }
          uint32_t dist = (d >> 16) + br.take((int)((d >> 8) & 15));
          S.tok[ntok++] = 0x80000000u | (length << 16) | (dist - 1);  // dist-1 fits 15 bits
        }
        if (st == INF_OK && br.overrun()) st = INF_INPUT_OVERRUN;
      }
      st = tile.shfl(st, 0);
      if (st != INF_OK) {
        status = st;
        break;
      }
      ntok = tile.shfl(ntok, 0);
      eob = tile.shfl((int)eob, 0);
      tile.sync();
      // ---- execute the batch ----
      for (int tbase = 0; tbase < ntok; tbase += G) {
        int ti = tbase + lane;
        uint32_t t = ti < ntok ? S.tok[ti] : 0;
        bool is_match = (t >> 31) != 0;
        uint32_t olen = ti < ntok ? (is_match ? ((t >> 16) & 0x1ff) : 1u) : 0u;
        uint32_t incl = olen;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
          uint32_t v = tile.shfl_up(incl, d);
          if (lane >= d) incl += v;
        }
        uint32_t total = tile.shfl(incl, G - 1);
        if (outpos + total > isize) {
          status = INF_OUTPUT_OVERRUN;
          break;
        }
        uint32_t mypos = outpos + incl - olen;
        if (ti < ntok && !is_match) out[mypos] = (uint8_t)t;
        uint32_t mmask = tile.ballot(ti < ntok && is_match);
        tile.sync();
        while (mmask) {
          int src_lane = __ffs((int)mmask) - 1;
          mmask &= mmask - 1;
          uint32_t mt = tile.shfl(t, src_lane);
          uint32_t mp = tile.shfl(mypos, src_lane);
          uint32_t len = (mt >> 16) & 0x1ff, dist = (mt & 0x7fff) + 1;
          if (dist > mp) {
            status = INF_BAD_DISTANCE;
            break;
          }
          const uint8_t *src = out + mp - dist;
          uint8_t *dst = out + mp;
          if (dist >= len) {
            // disjoint source and destination
            for (uint32_t i = lane; i < len; i += G) dst[i] = src[i];
          } else if (dist >= (uint32_t)G) {
            // overlapping with a period of at least one pass: a pass only reads bytes of earlier passes
            for (uint32_t b = 0; b < len; b += G) {
              uint32_t i = b + lane;
              if (i < len) dst[i] = src[i];
              tile.sync();
            }
          } else {
            // short period: every output byte i equals src[i % dist]
            uint32_t r = (uint32_t)lane % dist, step = (uint32_t)G % dist;
            for (uint32_t i = lane; i < len; i += G) {
              dst[i] = src[r];
              r += step;
              if (r >= dist) r -= dist;
            }
          }
          tile.sync();
        }
        if (status != INF_OK) break;
        outpos += total;
      }
    }
  }
  if (status == INF_OK && outpos != isize) status = INF_SIZE_MISMATCH;
  tile.sync();
  if (status == INF_OK && verify_crc && isize > 0) {
    uint32_t c = tile_crc32<G>(tile, out, isize, crcT);
    if (c != crc_expected) status = INF_CRC_MISMATCH;
  }
  return status;
}

template <int G, int THREADS>
__global__ void __launch_bounds__(THREADS) bgzf_inflate_kernel(InflateArgs a) {
  BND_DYN_SMEM(smem_raw);
  constexpr int GROUPS = THREADS / G;
  TileSmem *tiles = reinterpret_cast<TileSmem *>(smem_raw);
  CrcTables *crc_s = reinterpret_cast<CrcTables *>(smem_raw + sizeof(TileSmem) * GROUPS);
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(a.crc);
    uint32_t *dst = reinterpret_cast<uint32_t *>(crc_s);
    for (int i = threadIdx.x; i < (int)(sizeof(CrcTables) / 4); i += THREADS) dst[i] = src[i];
  }
  __syncthreads();
  Tile<G> tile;
  TileSmem &S = tiles[threadIdx.x / G];
  for (;;) {
    uint32_t b = 0;
    if (tile.lane == 0) b = atomicAdd(a.work_counter, 1u);
    b = tile.shfl(b, 0);
    if (b >= a.n_blocks) break;
    uint64_t c0 = a.blk_coff[b], c1 = a.blk_coff[b + 1];
    uint64_t u0 = a.blk_uoff[b], u1 = a.blk_uoff[b + 1];
    int st = inflate_block<G>(tile, S, a.comp + c0, (uint32_t)(c1 - c0), a.out + u0, (uint32_t)(u1 - u0), crc_s, a.verify_crc);
    if (tile.lane == 0) {
      if (a.status) a.status[b] = st;
      if (st != INF_OK) atomicCAS(a.error_flag, 0, st | ((int)(b & 0xffffff) << 8));
    }
    tile.sync();
  }
}

constexpr size_t inflate_smem_bytes(int groups) { return sizeof(TileSmem) * (size_t)groups + sizeof(CrcTables); }

}  // namespace bnd
