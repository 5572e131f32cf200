// inflate.cuh — K1: per-BGZF-block DEFLATE inflate + CRC32 check, hand-written for sm_100a.
//
// Replaces noodles-bgzf block decoding (the reference calls it through
// bam::io::indexed_reader, bamnado/src/coverage_analysis.rs:456-466).  One warp owns one BGZF block (<= 64 KiB in,
// <= 64 KiB out).  The Huffman walk is inherently serial, and ncu shows the kernel issue-bound on it (profiles/), so
// the serial part is kept to the bare dependency chain: lane 0 looks symbols up (funnel-shift bit window, u16
// shared-memory LUTs whose entries already carry the extra-bit counts), advances the bit position and stores the RAW
// token (LUT entries + the bit windows holding the extra bits).  Everything else happens once per batch with one
// token per lane: length/distance reconstruction, output positions by a warp scan, bounds checks, literal stores and
// LZ77 copies (copies that do not depend on the batch's own matches run concurrently, one match per lane).
// At the end the warp verifies ISIZE and the CRC32 with a lane-interleaved table CRC.  No GEMM shape, no tensor cores.
#pragma once
#include "common.cuh"
#include "kernel_args.h"

namespace bnd {

constexpr int NTOK = 32;       // tokens decoded per batch: one per lane in the parallel half

// Canonical description of one Huffman code, used to build the LUT and to decode codes longer than the LUT width.
struct HuffCanon {
  uint16_t first[16];   // canonical (MSB-first) value of the first code of each length
  uint16_t count[16];   // number of codes of each length
  uint16_t offset[16];  // index into sorted[] of the first symbol of each length
};

// LUT entries are u16: bits 0-4 = bits to consume (code length + extra bits; 0 = code longer than the LUT width, or
// invalid), symbol from bit 7 (literal/length) or bit 8 (distance).  The serial walk only needs that sum; code length
// and extra-bit count are separated again in the parallel half (the extra-bit count is a function of the symbol).
// LB / DB: primary lookup widths for literal/length and distance codes.
template <int LB, int DB>
struct alignas(16) TileSmemT {
  static constexpr int LIT_BITS = LB, DIST_BITS = DB;
  uint16_t lit_lut[1 << LB];
  uint16_t dist_lut[1 << DB];
  HuffCanon lit_canon, dist_canon;
  uint16_t lit_sorted[288];   // symbols ordered by (length, symbol)
  uint16_t dist_sorted[32];
  uint16_t codes[320];        // canonical code of each symbol (also scratch for the code-length code)
  uint8_t lens[320];          // code lengths of the block being set up (litlen, then distance at 288)
  // raw tokens of the batch: [0, NTOK) literal/length entry | distance entry << 16; [NTOK, 2 NTOK) bit window at the
  // length code (holds its extra bits); [2 NTOK, 3 NTOK) bit window at the distance code
  uint32_t raw[3 * NTOK];
};

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

// Build LUT + canonical tables for `n` symbols whose lengths are in S.lens[base .. base+n).
// Returns false on an over-subscribed code.
template <int G, class TileSmem>
__device__ bool build_tables(const Tile<G> &tile, TileSmem &S, int base, int n, int kind) {
  constexpr int LIT_BITS = TileSmem::LIT_BITS, DIST_BITS = TileSmem::DIST_BITS;
  const int lane = tile.lane;
  HuffCanon &C = kind == 0 ? S.lit_canon : S.dist_canon;
  uint16_t *sorted = kind == 0 ? S.lit_sorted : S.dist_sorted;
  uint16_t *lut = kind == 0 ? S.lit_lut : S.dist_lut;
  const int bits = kind == 0 ? LIT_BITS : DIST_BITS;
  for (int i = lane; i < (1 << bits); i += G) lut[i] = 0;
  bool ok = true;
  if (lane == 0) {
    for (int l = 0; l < 16; l++) C.count[l] = 0;
    for (int s = 0; s < n; s++) C.count[S.lens[base + s]]++;
    C.count[0] = 0;
    uint16_t next_code[16], offs[16];
    int left = 1;
    uint32_t code = 0;
    uint16_t off = 0;
    C.first[0] = 0;
    C.offset[0] = 0;
    for (int l = 1; l <= 15; l++) {
      left <<= 1;
      left -= C.count[l];
      if (left < 0) ok = false;
      code = (code + C.count[l - 1]) << 1;
      next_code[l] = (uint16_t)code;
      C.first[l] = (uint16_t)code;
      offs[l] = off;
      C.offset[l] = off;
      off += C.count[l];
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
  ok = tile.shfl((int)ok, 0) != 0;
  tile.sync();
  if (!ok) return false;
  for (int s = lane; s < n; s += G) {
    int l = S.lens[base + s];
    if (l == 0 || l > bits) continue;
    uint32_t rev = __brev((uint32_t)S.codes[base + s]) >> (32 - l);
    uint16_t e = (uint16_t)(kind == 0 ? lit_entry((uint32_t)s, (uint32_t)l) : dist_entry((uint32_t)s, (uint32_t)l));
    for (uint32_t i = rev; i < (1u << bits); i += (1u << l)) lut[i] = e;
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
    uint32_t r = 0;
    for (; r + 4 <= rounds; r += 4) {  // four loads in flight ahead of the table chain
      uint32_t v[4];
#pragma unroll
      for (int k = 0; k < 4; k++) v[k] = w[(r + k) * G + lane];
      if (r == 0 && lane == 0) v[0] ^= c;
#pragma unroll
      for (int k = 0; k < 4; k++)
        acc = T->z[0][acc & 0xff] ^ T->z[1][(acc >> 8) & 0xff] ^ T->z[2][(acc >> 16) & 0xff] ^ T->z[3][acc >> 24] ^ v[k];
    }
    for (; r < rounds; r++) {
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

// The same CRC-32 computed while the block is being produced: `advance` folds every complete group of 32 words that has
// become final since the last call (still in L1/L2: the block-end pass of tile_crc32 re-read the whole output from DRAM,
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

// The serial half of a batch: walk up to NTOK symbols, store their raw tokens in S.raw, advance the bit reader.
// Kept out of line on purpose: inside inflate_block the register allocator (capped by the occupancy target) kept
// re-deriving the shared-memory base addresses in every iteration (S2R + LEA + IMAD chains in the SASS); as a separate
// function the loop owns its ~16 live registers.  Returns ntok | status << 8 | eob << 16.
template <class TileSmem>
__device__ __noinline__ uint32_t walk_batch(BitReader &br_ref, TileSmem &S) {
  constexpr int LIT_BITS = TileSmem::LIT_BITS, DIST_BITS = TileSmem::DIST_BITS;
  BitReader br = br_ref;
  const saddr_t raw0 = smem_addr(S.raw), lit0 = smem_addr(S.lit_lut), dist0 = smem_addr(S.dist_lut);
  saddr_t tp = raw0;
  const saddr_t tend = raw0 + 4 * NTOK;
  uint32_t st = INF_OK, eob = 0;
#pragma unroll 1
  do {
    uint32_t w = br.window();
    uint32_t e = lds_u16(lit0 + 2 * (w & ((1u << LIT_BITS) - 1u)));
    if ((e & 31u) == 0) {  // rare: code longer than the lookup width
      uint32_t sym, len;
      if (!canon_decode(w, LIT_BITS, S.lit_canon, S.lit_sorted, sym, len)) {
        st = INF_BAD_SYMBOL;
        break;
      }
      e = lit_entry(sym, len);
    }
    br.consume(e & 31u);
    if (e >= (256u << 7)) {
      if (e < (257u << 7)) {
        eob = 1;
        break;
      }
      sts_u32(tp + 4 * NTOK, w);
      w = br.window();
      uint32_t d = lds_u16(dist0 + 2 * (w & ((1u << DIST_BITS) - 1u)));
      if ((d & 31u) == 0) {  // rare
        uint32_t ds, dl;
        if (!canon_decode(w, DIST_BITS, S.dist_canon, S.dist_sorted, ds, dl)) {
          st = INF_BAD_DISTANCE;
          break;
        }
        d = dist_entry(ds, dl);
      }
      sts_u32(tp + 8 * NTOK, w);
      br.consume(d & 31u);
      e |= d << 16;
    }
    sts_u32(tp, e);
    tp += 4;
  } while (tp != tend);
  if (st == INF_OK && br.overrun()) st = INF_INPUT_OVERRUN;
  br_ref = br;
  return (uint32_t)((tp - raw0) >> 2) | (st << 8) | (eob << 16);
}

// Inflate one BGZF block with a tile of G lanes.  Returns a status code (same on every lane).
template <int G, class TileSmem>
__device__ int inflate_block(const Tile<G> &tile, TileSmem &S, const SymTables &T, const uint8_t *blk, uint32_t blk_size,
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
    if (lane == 0) {
      final_block = br.take(1) != 0;
      btype = (int)br.take(2);
    }
    btype = tile.shfl(btype, 0);
    final_block = tile.shfl((int)final_block, 0) != 0;
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
        uint32_t l = br.take(16);
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
      for (int s = lane; s < 32; s += G) S.lens[288 + s] = s < 30 ? 5 : 0;
      tile.sync();
    } else {
      // dynamic: read code-length code, then HLIT + HDIST lengths (lane 0)
      int st = INF_OK;
      if (lane == 0) {
        st = read_dynamic_lengths(br, S.codes, S.lens);
      }
      st = tile.shfl(st, 0);
      if (st != INF_OK) {
        status = st;
        break;
      }
      tile.sync();
    }
    if (!build_tables<G, TileSmem>(tile, S, 0, 288, 0) || !build_tables<G, TileSmem>(tile, S, 288, 32, 1)) {
      status = INF_BAD_CODELEN;
      break;
    }

    // ---- symbol loop: lane 0 walks the stream and stores raw tokens; the warp finishes the batch, one token per lane ----
    bool eob = false;
    while (!eob && status == INF_OK) {
      int ntok = 0;
      int st = INF_OK;
      if (lane == 0) {
        const uint32_t r = walk_batch<TileSmem>(br, S);
        ntok = (int)(r & 0xffu);
        st = (int)((r >> 8) & 0xffu);
        eob = (r >> 16) != 0;
      }
      st = tile.shfl(st, 0);
      if (st != INF_OK) {
        status = st;
        break;
      }
      ntok = tile.shfl(ntok, 0);
      eob = tile.shfl((int)eob, 0) != 0;
      tile.sync();
      // ---- parallel half: lane t finishes token t ----
      for (int tbase = 0; tbase < ntok; tbase += G) {
        const int ti = tbase + lane;
        const bool mine = ti < ntok;
        const uint32_t tk = mine ? S.raw[ti] : 0u;
        const uint32_t sym = (tk >> 7) & 0x1ffu;
        const bool is_match = mine && sym > 256;
        uint32_t length = 0, dist = 1;
        bool bad = false;
        if (is_match) {
          const uint32_t lt = T.len_tab[(sym - 257) & 31u];   // 0 marks the invalid symbols 286, 287
          const uint32_t dt = T.dist_tab[(tk >> 24) & 31u];   // 0 marks the invalid symbols 30, 31
          const uint32_t eb = lt >> 16, deb = dt >> 16;
          const uint32_t len_bits = (tk & 31u) - eb, dist_bits = ((tk >> 16) & 31u) - deb;  // code lengths
          length = (lt & 0xffffu) + ((S.raw[NTOK + ti] >> len_bits) & ~(0xffffffffu << eb));
          dist = (dt & 0xffffu) + ((S.raw[2 * NTOK + ti] >> dist_bits) & ~(0xffffffffu << deb));
          bad = lt == 0 || dt == 0;
        }
        const uint32_t olen = is_match ? length : (mine ? 1u : 0u);
        uint32_t incl = olen;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
          uint32_t v = tile.shfl_up(incl, d);
          if (lane >= d) incl += v;
        }
        const uint32_t total = tile.shfl(incl, G - 1);
        const uint32_t mypos = outpos + incl - olen;
        bad = bad || (is_match && dist > mypos);
        const uint32_t badmask = tile.ballot(bad);
        if (badmask) {
          const uint32_t bsym = tile.shfl(sym, __ffs((int)badmask) - 1);
          status = bsym > 285 ? INF_BAD_SYMBOL : INF_BAD_DISTANCE;
          break;
        }
        if (outpos + total > isize) {
          status = INF_OUTPUT_OVERRUN;
          break;
        }
        if (mine && !is_match) out[mypos] = (uint8_t)sym;
        const uint32_t mmask = tile.ballot(is_match);
        if (mmask) {
          // ~96 % of the matches of a batch read only bytes produced before the batch's first match, so they do not
          // depend on each other: those (<= 16 bytes) are copied concurrently, one per lane, all loads before the
          // stores, and their L2/DRAM round trips overlap.  The others follow in stream order, the whole tile per match.
          const uint32_t first = tile.shfl(mypos, __ffs((int)mmask) - 1);
          const bool par = is_match && length <= 16 && mypos - dist + length <= first;
          tile.sync();  // this round's literals are visible
          if (par) {
            const uint8_t *src = out + mypos - dist;
            uint8_t *dst = out + mypos;
            uint8_t buf[16];
#pragma unroll
            for (int i = 0; i < 16; i++)
              if ((uint32_t)i < length) buf[i] = src[i];
#pragma unroll
            for (int i = 0; i < 16; i++)
              if ((uint32_t)i < length) dst[i] = buf[i];
          }
          uint32_t later = tile.ballot(is_match && !par);
          tile.sync();
#pragma unroll 1
          while (later) {
            const int sl = __ffs((int)later) - 1;
            later &= later - 1;
            const uint32_t len = tile.shfl(length, sl), dd = tile.shfl(dist, sl), mp = tile.shfl(mypos, sl);
            const uint8_t *src = out + mp - dd;
            uint8_t *dst = out + mp;
            if (dd >= len) {
              // disjoint source and destination
              for (uint32_t i = lane; i < len; i += G) dst[i] = src[i];
            } else if (dd >= (uint32_t)G) {
              // overlapping with a period of at least one pass: a pass only reads bytes of earlier passes
              for (uint32_t b = 0; b < len; b += G) {
                uint32_t i = b + lane;
                if (i < len) dst[i] = src[i];
                tile.sync();
              }
            } else {
              // short period: every output byte i equals src[i % dist]
              uint32_t r = (uint32_t)lane % dd, step = (uint32_t)G % dd;
              for (uint32_t i = lane; i < len; i += G) {
                dst[i] = src[r];
                r += step;
                if (r >= dd) r -= dd;
              }
            }
            tile.sync();
          }
        } else {
          tile.sync();
        }
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

template <int G, int THREADS, int MIN_CTAS, int LB, int DB>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) bgzf_inflate_kernel(InflateArgs a) {
  using TileSmem = TileSmemT<LB, DB>;
  BND_DYN_SMEM(smem_raw);
  constexpr int GROUPS = THREADS / G;
  TileSmem *tiles = reinterpret_cast<TileSmem *>(smem_raw);
  CrcTables *crc_s = reinterpret_cast<CrcTables *>(smem_raw + sizeof(TileSmem) * GROUPS);
  SymTables *sym_s = reinterpret_cast<SymTables *>(smem_raw + sizeof(TileSmem) * GROUPS + sizeof(CrcTables));
  if (threadIdx.x < 32) {
    const uint32_t s = threadIdx.x;
    // length symbols 257 + s: 8 codes without extra bits, then groups of 4 per extra-bit count, 285 -> 258
    uint32_t eb = s < 8 ? 0u : s == 28 ? 0u : (s - 4) >> 2;
    uint32_t base = s < 8 ? s + 3 : s == 28 ? 258u : ((4 + (s & 3)) << eb) + 3;
    sym_s->len_tab[s] = s < 29 ? (base | (eb << 16)) : 0u;
    uint32_t deb = s < 4 ? 0u : (s >> 1) - 1;
    uint32_t dbase = s < 4 ? s + 1 : ((2 + (s & 1)) << deb) + 1;
    sym_s->dist_tab[s] = s < 30 ? (dbase | (deb << 16)) : 0u;
  }
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
    int st = inflate_block<G, TileSmem>(tile, S, *sym_s, a.comp + c0, (uint32_t)(c1 - c0), a.out + u0, (uint32_t)(u1 - u0), crc_s, a.verify_crc);
    if (tile.lane == 0) {
      if (a.status) a.status[b] = st;
      if (st != INF_OK) atomicCAS(a.error_flag, 0, st | ((int)(b & 0xffffff) << 8));
    }
    tile.sync();
  }
}

template <int LB, int DB>
constexpr size_t inflate_smem_bytes(int groups) {
  return sizeof(TileSmemT<LB, DB>) * (size_t)groups + sizeof(CrcTables) + sizeof(SymTables);
}

}  // namespace bnd
