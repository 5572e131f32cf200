// inflate.cuh — building blocks of K1 (per-BGZF-block DEFLATE inflate + CRC32 check), hand-written for sm_100a.
//
// K1 replaces noodles-bgzf block decoding (the reference reaches it through bam::io::indexed_reader,
// bamnado/src/coverage_analysis.rs:456-466).  This header holds what the kernel in inflate_spec.cuh is made of: the
// canonical-code description, the LUT entry formats, the lane-0 bit reader used for block headers, the dynamic
// code-length reader (RFC 1951 3.2.7) and the lane-interleaved streaming CRC-32.  No GEMM shape, no tensor cores.
#pragma once
#include "common.cuh"
#include "kernel_args.h"

namespace bnd {

// Canonical description of one Huffman code, used to build the LUT and to decode codes longer than the LUT width.
struct HuffCanon {
  uint16_t first[16];   // canonical (MSB-first) value of the first code of each length
  uint16_t count[16];   // number of codes of each length
  uint16_t offset[16];  // index into sorted[] of the first symbol of each length
};

// LUT entries are u16: bits 0-4 = bits to consume (code length + extra bits; 0 = code longer than the LUT width, or
// invalid), symbol from bit 7 (literal/length) or bit 8 (distance).

// length / distance base values and extra-bit counts (RFC 1951 3.2.5): base | extra_bits << 16, one copy per CTA
struct SymTables {
  uint32_t len_tab[32];
  uint32_t dist_tab[32];
};

__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// ---- bit reader (used by lane 0 only): a 64-bit window held as two 32-bit words and a bit position ---------------
struct BitReader {
  const uint32_t *wp;   // next aligned word to fetch
  uint32_t lo, hi, nxt; // current window words and the prefetched one
  uint32_t pos;         // bits of `lo` already consumed, 0..31
  const uint8_t *end;   // first byte past the payload
  __device__ __forceinline__ void init(const uint8_t *p, const uint8_t *pend) {
    uintptr_t a = (uintptr_t)p;
    wp = (const uint32_t *)(a & ~(uintptr_t)3);
    pos = (uint32_t)(a & 3) * 8;
    lo = __ldg(wp);
    hi = __ldg(wp + 1);
    nxt = __ldg(wp + 2);
    wp += 3;
    end = pend;
  }
  // start at bit `bit` of the word array `base` (the speculative decoder addresses the payload by bit offset)
  __device__ __forceinline__ void init_bits(const uint32_t *base, uint32_t bit, const uint8_t *pend) {
    wp = base + (bit >> 5);
    pos = bit & 31u;
    lo = __ldg(wp);
    hi = __ldg(wp + 1);
    nxt = __ldg(wp + 2);
    wp += 3;
    end = pend;
  }
  __device__ __forceinline__ uint32_t bit_pos(const uint32_t *base) const { return (uint32_t)((wp - 3 - base) << 5) + pos; }
  __device__ __forceinline__ uint32_t window() const { return __funnelshift_r(lo, hi, pos); }  // next 32 bits
  __device__ __forceinline__ void consume(uint32_t n) {  // n <= 32
    pos += n;
    if (pos >= 32) {
      lo = hi;
      hi = nxt;
      nxt = __ldg(wp);
      wp++;
      pos -= 32;
    }
  }
  __device__ __forceinline__ uint32_t take(uint32_t n) {  // n <= 16
    uint32_t v = window() & ((1u << n) - 1u);
    consume(n);
    return v;
  }
  __device__ __forceinline__ const uint8_t *byte_ptr() const { return (const uint8_t *)(wp - 3) + (pos >> 3); }
  __device__ __forceinline__ void align_byte() { consume((8u - (pos & 7u)) & 7u); }
  __device__ __forceinline__ bool overrun() const { return byte_ptr() > end; }
};

// Decode a code longer than `root` bits (or report an invalid one) from the bit window by canonical ranges.
__device__ __forceinline__ bool canon_decode(uint32_t w, int root, const HuffCanon &C, const uint16_t *sorted, uint32_t &sym, uint32_t &len) {
  const uint32_t r = __brev(w);  // code bits MSB-first in the top bits
#pragma unroll 1
  for (int l = root + 1; l <= 15; l++) {
    uint32_t c = (r >> (32 - l)) - C.first[l];
    if (c < C.count[l]) {
      sym = sorted[C.offset[l] + c];
      len = (uint32_t)l;
      return true;
    }
  }
  return false;
}

__device__ __forceinline__ uint32_t lit_entry(uint32_t sym, uint32_t len) {
  const uint32_t s = sym - 257;
  const uint32_t eb = sym < 265 || sym >= 285 ? 0u : (s - 4) >> 2;
  return (sym << 7) | (len + eb);
}
__device__ __forceinline__ uint32_t dist_entry(uint32_t sym, uint32_t len) {
  const uint32_t eb = sym < 4 || sym >= 30 ? 0u : (sym >> 1) - 1;
  return (sym << 8) | (len + eb);
}

// The same CRC-32 computed while the block is being produced: `advance` folds every complete group of 32 words that has
// become final since the last call (still in L1/L2: a pass at the end of the block re-read the whole output from DRAM,
// profiles/r1_inflate_v15_spec_shipped.txt), `finish` adds the last words and bytes.  Lane l keeps the register for words
// l, l + 32, ...; the registers are only combined in `finish`.
struct CrcStream {
  uint32_t c = 0xffffffffu;  // serial register: leading bytes up to 4-byte alignment
  uint32_t acc = 0;          // this lane's interleaved register
  uint32_t words = 0;        // aligned words folded so far (multiple of 32)
  uint32_t lead = 0;
  bool started = false;
  __device__ __forceinline__ void advance(int lane, const uint8_t *out, uint32_t upto, const CrcTables *T) {
    if (!started) {
      lead = (uint32_t)((4 - ((uintptr_t)out & 3)) & 3);
      if (upto < lead) return;
      for (uint32_t i = 0; i < lead; i++) c = T->t[(c ^ out[i]) & 0xff] ^ (c >> 8);
      started = true;
    }
    const uint32_t *w = (const uint32_t *)(out + lead);
    const uint32_t avail = (upto - lead) >> 2;
    while (words + 128 <= avail) {  // four loads in flight ahead of the table chain
      uint32_t v[4];
#pragma unroll
      for (int k = 0; k < 4; k++) v[k] = w[words + 32 * k + lane];
      if (words == 0 && lane == 0) v[0] ^= c;
#pragma unroll
      for (int k = 0; k < 4; k++)
        acc = T->z[0][acc & 0xff] ^ T->z[1][(acc >> 8) & 0xff] ^ T->z[2][(acc >> 16) & 0xff] ^ T->z[3][acc >> 24] ^ v[k];
      words += 128;
    }
    while (words + 32 <= avail) {
      uint32_t v = w[words + lane];
      if (words == 0 && lane == 0) v ^= c;
      acc = T->z[0][acc & 0xff] ^ T->z[1][(acc >> 8) & 0xff] ^ T->z[2][(acc >> 16) & 0xff] ^ T->z[3][acc >> 24] ^ v;
      words += 32;
    }
  }
  // all n bytes of the block are final; every lane returns the CRC
  __device__ __forceinline__ uint32_t finish(const Tile<32> &tile, const uint8_t *out, uint32_t n, const CrcTables *T) {
    const int lane = tile.lane;
    if (!started && n > 0) {
      lead = (uint32_t)((4 - ((uintptr_t)out & 3)) & 3);
      if (lead > n) lead = n;
      for (uint32_t i = 0; i < lead; i++) c = T->t[(c ^ out[i]) & 0xff] ^ (c >> 8);
      started = true;
    }
    if (n >= lead) advance(lane, out, n, T);
    uint32_t r = c;
    if (words) {
      // lane l still has to advance by (32 - l) words of zeros
      for (int k = 0; k < 32 - lane; k++) {
#pragma unroll
        for (int b = 0; b < 4; b++) acc = T->t[acc & 0xff] ^ (acc >> 8);
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) acc ^= tile.shfl_xor(acc, d);
      r = acc;
    }
    const uint32_t done = lead + words * 4;
    for (uint32_t i = done; i < n; i++) r = T->t[(r ^ out[i]) & 0xff] ^ (r >> 8);  // < 128 bytes, serial
    return ~r;
  }
};

// Dynamic block header (RFC 1951 3.2.7), read by ONE lane: HLIT/HDIST/HCLEN, the code-length code, then the literal/
// length and distance code lengths into lens[0..288) and lens[288..320).  Returns an INF_* status.
__device__ inline int read_dynamic_lengths(BitReader &br, uint16_t *cl_lut, uint8_t *lens) {
  int st = INF_OK;
  int hlit = (int)br.take(5) + 257;
  int hdist = (int)br.take(5) + 1;
  int hclen = (int)br.take(4) + 4;
  if (hlit > 286 || hdist > 30) st = INF_BAD_CODELEN;
  uint8_t cl[19];
#pragma unroll
  for (int i = 0; i < 19; i++) cl[i] = 0;
  for (int i = 0; i < hclen; i++) cl[c_cl_order[i]] = (uint8_t)br.take(3);
  // cl_lut: 7-bit LUT for the code-length code (128 x u16 of scratch not yet in use for this block)
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
        uint32_t rev = __brev(next[l]++) >> (32 - l);
        for (uint32_t i = rev; i < 128; i += (1u << l)) cl_lut[i] = (uint16_t)((s << 4) | l);
      }
  }
  int n = 0, total = hlit + hdist;
  int prev = 0;
  while (st == INF_OK && n < total) {
    uint32_t w = br.window();
    uint32_t e = cl_lut[w & 127];
    int l = e & 15, s = e >> 4;
    if (l == 0) {
      st = INF_BAD_CODELEN;
      break;
    }
    if (s < 16) {
      br.consume(l);
      lens[n++] = (uint8_t)s;
      prev = s;
    } else {
      int rep, val;
      if (s == 16) {
        if (n == 0) {
          st = INF_BAD_CODELEN;
          break;
        }
        rep = 3 + (int)((w >> l) & 3);
        br.consume(l + 2);
        val = prev;
      } else if (s == 17) {
        rep = 3 + (int)((w >> l) & 7);
        br.consume(l + 3);
        val = 0;
        prev = 0;
      } else {
        rep = 11 + (int)((w >> l) & 127);
        br.consume(l + 7);
        val = 0;
        prev = 0;
      }
      if (n + rep > total) {
        st = INF_BAD_CODELEN;
        break;
      }
      for (int k = 0; k < rep; k++) lens[n++] = (uint8_t)val;
    }
  }
  if (st == INF_OK && lens[256] == 0) st = INF_BAD_CODELEN;  // no end-of-block code
  if (st == INF_OK) {
    // move the distance lengths to their fixed slot at 288 (hlit <= 286 < 288: moving up, copy from the top)
    for (int k = hdist - 1; k >= 0; k--) lens[288 + k] = lens[hlit + k];
    for (int k = hlit; k < 288; k++) lens[k] = 0;
    for (int k = hdist; k < 32; k++) lens[288 + k] = 0;
  }
  return st;
}

}  // namespace bnd
