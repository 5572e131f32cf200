// inflate_spec.cuh — K1, second generation: the Huffman walk itself runs 32 lanes wide.
//
// profiles/r1_inflate_v10_outofline_walk.txt: with one lane walking the stream, 68 % of the warp instructions of K1 are
// executed with ONE active thread (≈ 42 instructions per token).  A prefix code re-synchronises by itself: a decoder
// started at a wrong bit offset falls onto true code boundaries after a few codes and stays on them.  So a round cuts
// the next 32 * SPAN bits of the stream into 32 spans, one per lane:
//   pass 1   every lane decodes its span from the span's first bit (lane 0 from the true position) and keeps only where
//            it left the span (`exit`: the first code boundary at or beyond the span's end);
//   fix-up   lane i restarts from the exit of lane i-1 whenever that differs from where it started; repeated until
//            nothing changes.  Lane 0 is right by construction and every iteration makes at least one more lane right,
//            so the loop ends after <= 32 iterations (2.6 on average on BAM data: 90 % of the lanes synchronise inside
//            their own span);
//   store    every lane now knows how many bytes and matches the lanes before it produce (a warp scan), so a last pass
//            writes the literals straight to their place in the output and appends the matches (position, length,
//            distance) to one list in shared memory.
// The warp then resolves the round's matches 32 at a time: those whose source lies before the first unfinished match
// are copied concurrently (their L2 round trips overlap), the rest follow wave by wave.  Results are bit-identical
// to a serial inflate by construction: only spans that start on a verified code boundary contribute tokens.
#pragma once
#include "inflate.cuh"

// Register cap of the kernel: 960 threads x 56 registers leave room on every SM for one block of the small kernels of the
// previous segment (K2, K3/K4), which otherwise wait for the whole launch to drain (profiles/r1_sweeps.md).
#ifndef BND_K1_MAXREG
#define BND_K1_MAXREG 56
#endif
namespace bnd {

// Per-warp shared memory.  The code-length scratch (codes/lens) is only live while a block's tables are being built
// and shares the match list.
template <int LB, int DB, int NT, int SPAN_BITS>
struct alignas(16) SpecSmemT {
  static constexpr int LIT_BITS = LB, DIST_BITS = DB, MAX_MATCHES = NT, SPAN = SPAN_BITS;
  static constexpr int LIT_SUB = 344;              // second-level entries: zlib's bound for 286 symbols is 308 (root 10) / 340 (root 9), + 2 reserved
  static constexpr int STAGE_WORDS = SPAN_BITS + 8;  // one round of input: 32 spans + the longest token + the reader's look-ahead
  // literal/length code: root table of 2^LB entries, then the second-level tables of the codes longer than LB bits.
  // entry: bits 0-4 = bits to consume (code + extra bits), symbol from bit 7; bits 0-4 == 0: second-level pointer with
  // bits 5-7 = index width, bits 8-15 = (offset / 2) into the second level (0 = the reserved all-zero entry = no code)
  uint16_t lit_lut[(1 << LB) + LIT_SUB];
  uint16_t dist_lut[1 << DB];
  HuffCanon dist_canon;
  uint16_t dist_sorted[32];
  uint32_t stage[STAGE_WORDS];  // this round's compressed words
  // the round's matches in stream order: output position | (length - 3) << 16, and the distance (<= 32768) beside it.  Six bytes
  // per match instead of eight: the list of a 384-bit round (448 entries) fits where 384 entries of eight bytes were
  // (profiles/r2_sweeps.md: longer spans pay, a list shorter than ~1.17 matches per span bit does not)
  uint32_t m_pl[NT];
  uint16_t m_dist[NT];
  __device__ __forceinline__ uint16_t *codes() { return reinterpret_cast<uint16_t *>(m_pl); }     // 320 x u16
  __device__ __forceinline__ uint8_t *lens() { return reinterpret_cast<uint8_t *>(m_pl) + 640; }  // 320 x u8
  __device__ __forceinline__ uint16_t *table_scratch() { return reinterpret_cast<uint16_t *>(m_pl) + 480; }  // 48 x u16 at byte 960
  static_assert(NT * 6 >= 1056 && NT % 8 == 0, "match list (m_pl and m_dist are adjacent) must hold the table-building scratch");
};

enum : uint32_t { SPAN_OK = 0, SPAN_EOB = 1, SPAN_BAD_SYMBOL = 2, SPAN_BAD_DISTANCE = 3, SPAN_PAST_END = 4 };

// Literal/length tables from lens[0..288): canonical codes, root LUT, second-level tables.  False on an over-subscribed
// code or one whose second level does not fit (impossible for a complete code).
template <class Smem>
__device__ bool build_lit_two_level(const Tile<32> &tile, Smem &S) {
  constexpr int ROOT = Smem::LIT_BITS, SUB = Smem::LIT_SUB;
  const int lane = tile.lane;
  uint16_t *lut = S.lit_lut;
  uint16_t *codes = S.codes();
  const uint8_t *lens = S.lens();
  // count / first code / next code per length: 48 x u16 behind the code-length scratch (lane 0 kept them in local memory before:
  // 370 LDL / STL in the SASS of round 2, and 2 x 288 iterations executed by one lane)
  uint16_t *count = S.table_scratch(), *first = count + 16, *next_code = count + 32;
  for (int i = lane; i < (1 << ROOT) + SUB; i += 32) lut[i] = 0;
  if (lane < 16) count[lane] = 0;
  tile.sync();
  // histogram of the 288 code lengths, 32 symbols at a time: the lanes holding the same length elect one of them to add their number
#pragma unroll 1
  for (int c = 0; c < 9; c++) {
    const uint32_t l = lens[c * 32 + lane];
    const unsigned peers = __match_any_sync(0xffffffffu, l);
    if (l && lane == __ffs((int)peers) - 1) count[l] = (uint16_t)(count[l] + __popc(peers));
    tile.sync();
  }
  bool ok = true;
  if (lane == 0) {
    int left = 1;
    uint32_t code = 0;
    first[0] = 0;
    for (int l = 1; l <= 15; l++) {
      left <<= 1;
      left -= count[l];
      if (left < 0) ok = false;
      code = (code + count[l - 1]) << 1;
      next_code[l] = (uint16_t)code;
      first[l] = (uint16_t)code;
    }
  }
  ok = tile.shfl((int)ok, 0) != 0;
  tile.sync();
  if (!ok) return false;
  // canonical codes: symbols of one length take consecutive codes in symbol order
#pragma unroll 1
  for (int c = 0; c < 9; c++) {
    const int s = c * 32 + lane;
    const uint32_t l = lens[s];
    const unsigned peers = __match_any_sync(0xffffffffu, l);
    const uint32_t base = next_code[l & 15u];
    tile.sync();
    if (l) {
      codes[s] = (uint16_t)(base + (uint32_t)__popc(peers & ((1u << lane) - 1u)));
      if (lane == __ffs((int)peers) - 1) next_code[l] = (uint16_t)(base + (uint32_t)__popc(peers));
    }
    tile.sync();
  }
  if (lane == 0) {
    {
      // second-level tables, longest codes first: the first code length that reaches a root prefix is its longest
      uint32_t off = 2;  // entries 0, 1 stay zero: "no code"
      for (int l = 15; l > ROOT && ok; l--) {
        if (!count[l]) continue;
        const uint32_t plo = (uint32_t)first[l] >> (l - ROOT), phi = ((uint32_t)first[l] + count[l] - 1u) >> (l - ROOT);
        for (uint32_t p = plo; p <= phi; p++) {
          const uint32_t rp = __brev(p) >> (32 - ROOT);
          if (lut[rp] != 0) continue;
          const uint32_t sb = (uint32_t)(l - ROOT);
          if (off + (1u << sb) > (uint32_t)SUB) {
            ok = false;
            break;
          }
          lut[rp] = (uint16_t)((sb << 5) | ((off >> 1) << 8));
          off += 1u << sb;
        }
      }
    }
  }
  ok = tile.shfl((int)ok, 0) != 0;
  tile.sync();
  if (!ok) return false;
  for (int s = lane; s < 288; s += 32) {
    const int l = lens[s];
    if (l == 0) continue;
    const uint32_t code = codes[s];
    const uint16_t e = (uint16_t)lit_entry((uint32_t)s, (uint32_t)l);
    if (l <= ROOT) {
      const uint32_t rev = __brev(code) >> (32 - l);
      for (uint32_t i = rev; i < (1u << ROOT); i += (1u << l)) lut[i] = e;
    } else {
      const int k = l - ROOT;
      const uint32_t pe = lut[__brev(code >> k) >> (32 - ROOT)];
      const uint32_t sb = (pe >> 5) & 7u, sub0 = (1u << ROOT) + ((pe >> 8) << 1);
      const uint32_t rlow = __brev(code & ((1u << k) - 1u)) >> (32 - k);
      for (uint32_t i = rlow; i < (1u << sb); i += (1u << k)) lut[sub0 + i] = e;
    }
  }
  tile.sync();
  return true;
}

// Distance tables from lens[288..320): LUT of 2^DB entries + canonical ranges for the (rare) longer codes.
template <class Smem>
__device__ bool build_dist_tables(const Tile<32> &tile, Smem &S) {
  constexpr int BITS = Smem::DIST_BITS;
  const int lane = tile.lane;
  uint16_t *codes = S.codes() + 288;
  const uint8_t *lens = S.lens() + 288;
  HuffCanon &C = S.dist_canon;
  for (int i = lane; i < (1 << BITS); i += 32) S.dist_lut[i] = 0;
  bool ok = true;
  if (lane == 0) {
    for (int l = 0; l < 16; l++) C.count[l] = 0;
    for (int s = 0; s < 32; s++) C.count[lens[s]]++;
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
      for (int s = 0; s < 32; s++) {
        const int l = lens[s];
        if (l) {
          codes[s] = next_code[l]++;
          S.dist_sorted[offs[l]++] = (uint16_t)s;
        }
      }
  }
  ok = tile.shfl((int)ok, 0) != 0;
  tile.sync();
  if (!ok) return false;
  {
    const int l = lens[lane];
    if (l != 0 && l <= BITS) {
      const uint32_t rev = __brev((uint32_t)codes[lane]) >> (32 - l);
      const uint16_t e = (uint16_t)dist_entry((uint32_t)lane, (uint32_t)l);
      for (uint32_t i = rev; i < (1u << BITS); i += (1u << l)) S.dist_lut[i] = e;
    }
  }
  tile.sync();
  return true;
}


// Decode tokens from bit `start` until the position reaches `limit` (> start), an end-of-block code or an undecodable
// code.  Bits are addressed from the first bit of the staged words.  Returns {bit position after the last token |
// flag << 24, matches | output bytes << 12}.  STORE: the span is known to be a true one and `opos` is the output position
// of its first byte: literals go straight to out[], matches are appended to the list at `ml`.
// The loop body has one shape for literals and matches (the distance lookup runs for every token and counts only for
// matches): lanes on different token kinds do not diverge, so an iteration costs one pass through the body.
template <bool STORE, class Smem>
__device__ __noinline__ uint2 span_decode(uint32_t start, uint32_t limit, Smem &S, const SymTables &T, uint8_t *out, uint32_t opos, saddr_t ml) {
  constexpr int LIT_BITS = Smem::LIT_BITS, DIST_BITS = Smem::DIST_BITS;
  const saddr_t lit0 = smem_addr(S.lit_lut), dist0 = smem_addr(S.dist_lut);
  saddr_t wa = smem_addr(S.stage) + 4 * (start >> 5);
  uint32_t pos = start & 31u;
  uint32_t lo = lds_u32(wa), hi = lds_u32(wa + 4), nxt = lds_u32(wa + 8);
  wa += 12;
  int32_t remain = (int32_t)(limit - start);
  uint32_t nm = 0, bytes = 0, flag = SPAN_OK;
  const saddr_t dadj = STORE ? smem_addr(S.m_dist) - (smem_addr(S.m_pl) >> 1) : 0;  // distance of entry i lives at (address of m_pl[i]) / 2 + dadj
#pragma unroll 1
  do {
    const uint32_t w = __funnelshift_r(lo, hi, pos);
    uint32_t e = lds_u16(lit0 + 2 * (w & ((1u << LIT_BITS) - 1u)));
    if ((e & 31u) == 0)  // second level (or the reserved zero entry)
      e = lds_u16(lit0 + 2 * ((1u << LIT_BITS) + ((e >> 8) << 1) + ((w >> LIT_BITS) & ~(0xffffffffu << ((e >> 5) & 7u)))));
    const uint32_t nb = e & 31u;
    if (nb == 0) {
      flag = SPAN_BAD_SYMBOL;
      break;
    }
    pos += nb;
    if (pos >= 32) {
      lo = hi;
      hi = nxt;
      nxt = lds_u32(wa);
      wa += 4;
      pos -= 32;
    }
    const uint32_t sym = e >> 7;
    if (sym == 256) {
      remain -= (int32_t)nb;
      flag = SPAN_EOB;
      break;
    }
    const bool is_match = sym > 256;
    // length base | extra bits << 16 (0 marks the invalid symbols 286, 287); a literal counts as base 1 without extra bits
    const uint32_t lt = is_match ? T.len_tab[(sym - 257u) & 31u] : 1u;
    const uint32_t w2 = __funnelshift_r(lo, hi, pos);
    uint32_t d = lds_u16(dist0 + 2 * (w2 & ((1u << DIST_BITS) - 1u)));
    if (is_match && ((d & 31u) == 0 || lt == 0)) {  // rare: distance code longer than the lookup width, or a bad symbol
      uint32_t ds, dl;
      if (lt == 0) {
        flag = SPAN_BAD_SYMBOL;
        break;
      }
      if (!canon_decode(w2, DIST_BITS, S.dist_canon, S.dist_sorted, ds, dl)) {
        flag = SPAN_BAD_DISTANCE;
        break;
      }
      d = dist_entry(ds, dl);
    }
    const uint32_t db = is_match ? (d & 31u) : 0u;
    pos += db;
    if (pos >= 32) {
      lo = hi;
      hi = nxt;
      nxt = lds_u32(wa);
      wa += 4;
      pos -= 32;
    }
    remain -= (int32_t)(nb + db);
    const uint32_t eb = lt >> 16;
    const uint32_t length = (lt & 0xffffu) + ((w >> (nb - eb)) & ~(0xffffffffu << eb));
    if (STORE) {
      if (is_match) {
        const uint32_t dt = T.dist_tab[(d >> 8) & 31u];  // 0 marks the invalid symbols 30, 31
        const uint32_t deb = dt >> 16;
        const uint32_t dist = (dt & 0xffffu) + ((w2 >> (db - deb)) & ~(0xffffffffu << deb));
        if (dt == 0 || dist > opos + bytes) {
          flag = SPAN_BAD_DISTANCE;
          break;
        }
        sts_u32(ml, (opos + bytes) | ((length - 3u) << 16));
        sts_u16((ml >> 1) + dadj, dist);
        ml += 4;
      } else {
        out[opos + bytes] = (uint8_t)sym;
      }
    }
    bytes += length;
    nm += is_match ? 1u : 0u;
  } while (remain > 0);
  return make_uint2((uint32_t)((int32_t)limit - remain) | (flag << 24), nm | (bytes << 12));
}


// Inflate one BGZF block with one warp.  Returns a status code (same on every lane).
// ZLIB: the same walk over a zlib stream (RFC 1950: 2-byte header, DEFLATE, big-endian Adler-32) as BigWig sections are stored
// (bigwig_io.cuh); `isize_expected` is then the capacity of `out`, the inflated size is returned through `out_len` and the
// Adler-32 is checked instead of the CRC-32.
template <class Smem, bool ZLIB = false>
__device__ int inflate_block_spec(const Tile<32> &tile, Smem &S, const SymTables &T, const uint8_t *blk, uint32_t blk_size, uint8_t *out,
                                  uint32_t isize_expected, const CrcTables *crcT, int verify_crc, uint32_t *out_len = nullptr) {
  constexpr int NM = Smem::MAX_MATCHES, SPAN = Smem::SPAN, STAGE_WORDS = Smem::STAGE_WORDS;
  static_assert(SPAN >= 64 && (SPAN + 64) / 2 <= NM, "a span's matches (>= 2 bits each) must fit the match list");
  const int lane = tile.lane;
  const uint8_t *payload, *pend;
  uint32_t crc_expected, isize;
  if (ZLIB) {
    if (blk_size < 6) return INF_BAD_HEADER;
    const uint32_t cmf = blk[0], flg = blk[1];
    if ((cmf & 0x0fu) != 8u || ((cmf << 8) | flg) % 31u != 0u || (flg & 0x20u)) return INF_BAD_HEADER;  // deflate, valid check bits, no preset dictionary
    payload = blk + 2;
    pend = blk + blk_size - 4;
    crc_expected = ((uint32_t)pend[0] << 24) | ((uint32_t)pend[1] << 16) | ((uint32_t)pend[2] << 8) | pend[3];  // Adler-32, big-endian
    isize = isize_expected;  // capacity
  } else {
    if (blk_size < 26) return INF_BAD_HEADER;
    const uint32_t xlen = blk[10] | ((uint32_t)blk[11] << 8);
    if (12 + xlen + 8 > blk_size) return INF_BAD_HEADER;
    payload = blk + 12 + xlen;
    pend = blk + blk_size - 8;
    crc_expected = pend[0] | ((uint32_t)pend[1] << 8) | ((uint32_t)pend[2] << 16) | ((uint32_t)pend[3] << 24);
    isize = pend[4] | ((uint32_t)pend[5] << 8) | ((uint32_t)pend[6] << 16) | ((uint32_t)pend[7] << 24);
    if (isize != isize_expected) return INF_SIZE_MISMATCH;
  }
  if (isize > 65536u) return INF_BAD_HEADER;  // BGZF blocks hold at most 64 KiB (SAM spec 4.1); match positions are 16-bit here

  // the payload as a bit-addressed array of aligned words
  const uint32_t *base = (const uint32_t *)((uintptr_t)payload & ~(uintptr_t)3);
  uint32_t p0 = (uint32_t)((uintptr_t)payload & 3) * 8;
  const uint32_t end_bit = p0 + (uint32_t)(pend - payload) * 8;
  const saddr_t ml0 = smem_addr(S.m_pl), stage0 = smem_addr(S.stage);
  const uint32_t word_limit = ((end_bit + 31u) >> 5) + 2u;  // words holding payload bits, + 2 inside the block's 8-byte trailer
  // asynchronous copy of one round of input, starting at the word holding bit p: global -> shared, zeros past the payload.
  // (One cp.async.bulk per round with an mbarrier - UBLKCP / SYNCS.PHASECHK in the SASS - was built and measured in round 2:
  // bit-exact, but the pass went 115.5 -> 123.1 ms at 40 M pairs: the engine's operands live in uniform registers and the round
  // loop, which is at the 56-register cap, picked up 43 spill instructions inside the match resolution.  Eleven 4-byte LDGSTS per
  // lane are 0.2 % of a round's instructions, so there was nothing to win on the other side; profiles/r2_sweeps.md.)
  auto stage_issue = [&](uint32_t p) {
    const uint32_t w0 = p >> 5;
    for (uint32_t k = (uint32_t)lane; k < (uint32_t)STAGE_WORDS; k += 32) {
      if (w0 + k < word_limit)
        cp_async_u32(stage0 + 4 * k, base + w0 + k);
      else
        sts_u32(stage0 + 4 * k, 0u);
    }
  };
  uint32_t outpos = 0;
  int status = INF_OK;
  bool final_block = false;
  CrcStream crc;

  while (!final_block && status == INF_OK) {
    // ---- block header (lane 0), broadcast ----
    int btype = 0, st = INF_OK;
    uint32_t np = p0, stored_len = 0;
    if (lane == 0) {
      if (p0 + 3 > end_bit) {
        st = INF_INPUT_OVERRUN;
      } else {
        BitReader br;
        br.init_bits(base, p0, pend);
        final_block = br.take(1) != 0;
        btype = (int)br.take(2);
        if (btype == 0) {
          br.align_byte();
          const uint32_t l = br.take(16), nl = br.take(16);
          if ((l ^ nl) != 0xffffu) st = INF_BAD_STORED;
          stored_len = l;
        } else if (btype == 2) {
          st = read_dynamic_lengths(br, S.codes(), S.lens());
        } else if (btype == 3) {
          st = INF_BAD_BTYPE;
        }
        np = br.bit_pos(base);
      }
    }
    st = tile.shfl(st, 0);
    if (st != INF_OK) {
      status = st;
      break;
    }
    btype = tile.shfl(btype, 0);
    final_block = tile.shfl((int)final_block, 0) != 0;
    p0 = tile.shfl(np, 0);
    if (btype == 0) {
      stored_len = tile.shfl(stored_len, 0);
      const uint8_t *src = (const uint8_t *)base + (p0 >> 3);
      if (p0 + stored_len * 8 > end_bit) {
        status = INF_INPUT_OVERRUN;
        break;
      }
      if (outpos + stored_len > isize) {
        status = INF_OUTPUT_OVERRUN;
        break;
      }
      for (uint32_t i = lane; i < stored_len; i += 32) out[outpos + i] = src[i];
      outpos += stored_len;
      p0 += stored_len * 8;
      tile.sync();
      if (!ZLIB && verify_crc) crc.advance(lane, out, outpos, crcT);
      continue;
    }
    if (btype == 1) {
      uint8_t *lens = S.lens();
      for (int s = lane; s < 288; s += 32) lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
      lens[288 + lane] = lane < 30 ? 5 : 0;
    }
    tile.sync();
    if (!build_lit_two_level<Smem>(tile, S) || !build_dist_tables<Smem>(tile, S)) {
      status = INF_BAD_CODELEN;
      break;
    }

    bool eob = false, staged = false;
    while (!eob && status == INF_OK) {
      if (p0 >= end_bit) {
        status = INF_INPUT_OVERRUN;
        break;
      }
      if (!staged) stage_issue(p0);
      cp_async_wait_all();
      tile.sync();
      const uint32_t sbit0 = p0 & ~31u;  // bit address of the first staged word
      // ---- pass 1: every lane decodes its span from the span's first bit ----
      uint32_t start = p0 + (uint32_t)lane * SPAN;
      const uint32_t limit = tmin(p0 + (uint32_t)(lane + 1) * SPAN, end_bit);
      uint32_t exitb = start, flag = SPAN_PAST_END, cnt = 0;  // cnt: matches | bytes << 12
      auto scan = [&]() {
        if (start < end_bit) {
          const uint2 r = span_decode<false>(start - sbit0, limit - sbit0, S, T, nullptr, 0u, 0);
          exitb = (r.x & 0xffffffu) + sbit0;
          flag = r.x >> 24;
          cnt = r.y;
        } else {
          exitb = start;
          flag = SPAN_PAST_END;
          cnt = 0;
        }
      };
      scan();
      // ---- fix-up: restart from the predecessor's exit until every live lane starts where its predecessor stopped ----
      int first_term;
      for (;;) {
        uint32_t want = tile.shfl_up(exitb, 1);
        if (lane == 0) want = p0;
        const uint32_t term = tile.ballot(flag != SPAN_OK);
        first_term = term ? __ffs((int)term) - 1 : 32;
        const bool need = lane <= first_term && want != start;
        if (tile.ballot(need) == 0) break;
        if (need) {
          start = want;
          scan();
        }
      }
      // ---- place: lanes 0..first_term hold true tokens; take as many whole lanes as the match list holds ----
      const int nvalid = first_term < 32 ? first_term + 1 : 32;
      const uint32_t nm = lane < nvalid ? (cnt & 0xfffu) : 0u, nbytes = lane < nvalid ? (cnt >> 12) : 0u;
      uint32_t minc = nm, binc = nbytes;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = tile.shfl_up(minc, d), u = tile.shfl_up(binc, d);
        if (lane >= d) {
          minc += v;
          binc += u;
        }
      }
      const bool taken = lane < nvalid && minc <= (uint32_t)NM;
      const int m = __popc(tile.ballot(taken));  // >= 1: one span never holds more than NM matches
      const uint32_t total = tile.shfl(binc, m - 1), nmatch = tile.shfl(minc, m - 1);
      if (outpos + total > isize) {
        status = INF_OUTPUT_OVERRUN;
        break;
      }
      // ---- store pass: literals to out[], matches to the list ----
      bool bad = false;
      if (taken && start < end_bit) {
        const uint2 r = span_decode<true>(start - sbit0, limit - sbit0, S, T, out, outpos + binc - nbytes, ml0 + 4 * (minc - nm));
        bad = (r.x & 0xffffffu) + sbit0 != exitb || r.y != cnt;  // only an invalid distance ends the store pass early
        if (bad) flag = r.x >> 24;
      }
      const uint32_t badmask = tile.ballot(taken && (bad || flag >= SPAN_BAD_SYMBOL));
      if (badmask) {
        const uint32_t f = tile.shfl(flag, __ffs((int)badmask) - 1);
        status = f == SPAN_BAD_SYMBOL ? INF_BAD_SYMBOL : f == SPAN_BAD_DISTANCE ? INF_BAD_DISTANCE : INF_INPUT_OVERRUN;
        break;
      }
      p0 = tile.shfl(exitb, m - 1);
      eob = tile.shfl(flag, m - 1) == SPAN_EOB;
      if (eob && p0 > end_bit) {
        status = INF_INPUT_OVERRUN;
        break;
      }
      tile.sync();  // literals and the match list are visible to the whole warp; nobody reads the staged words any more
      staged = !eob && p0 < end_bit;
      if (staged) stage_issue(p0);  // the next round's input arrives while this round's matches are resolved
      // ---- resolve the matches, 32 at a time: lane t owns match mb + t ----
      for (uint32_t mb = 0; mb < nmatch; mb += 32) {
        const bool mine = mb + (uint32_t)lane < nmatch;
        const uint2 me = mine ? make_uint2(S.m_pl[mb + lane], S.m_dist[mb + lane]) : make_uint2(0u, 1u);
        const uint32_t mypos = me.x & 0xffffu, length = (me.x >> 16) + 3u, dist = me.y;
        const uint32_t src_end = mypos - dist + length;
        uint32_t pend = tile.ballot(mine);
        // Everything before the first unfinished match is final (literals were stored by the store pass), so every
        // match whose source ends there can be copied now, all of them at once; the first unfinished one itself always
        // can (its source precedes it, or it repeats its own first bytes).  Usually one or two waves empty the batch.
#pragma unroll 1
        while (pend) {
          const int fl = __ffs((int)pend) - 1;
          const uint32_t frontier = tile.shfl(mypos, fl);
          const bool ready = ((pend >> lane) & 1u) != 0 && src_end <= frontier;
          const uint32_t rmask = tile.ballot(ready);
          if (rmask == 0) {
            // the first unfinished match overlaps its own output: period `dist` < length, copied by the whole warp
            const uint32_t len = tile.shfl(length, fl), dd = tile.shfl(dist, fl);
            const uint8_t *src = out + frontier - dd;
            uint8_t *dst = out + frontier;
            if (dd >= 32u) {
              // a pass of 32 bytes only reads bytes of earlier passes
              for (uint32_t b = 0; b < len; b += 32) {
                const uint32_t i = b + lane;
                if (i < len) dst[i] = src[i];
                tile.sync();
              }
            } else {
              // short period: output byte i equals src[i % dist]
              uint32_t r = (uint32_t)lane % dd;
              const uint32_t step = 32u % dd;
              for (uint32_t i = lane; i < len; i += 32) {
                dst[i] = src[r];
                r += step;
                if (r >= dd) r -= dd;
              }
            }
            pend &= ~(1u << fl);
            tile.sync();
            continue;
          }
          // Many ready matches: the short ones (<= 16 bytes) one per lane, the others by the whole warp, up to three at
          // a time (<= 64 bytes each); all loads are issued before the first store.  Up to three ready matches (the usual
          // case after a batch's first wave): all of them by the whole warp, which costs a fraction of the per-lane path.
          const bool few = __popc(rmask) <= 3;
          const bool par = !few && ready && length <= 16;
          uint32_t wide = rmask & ~tile.ballot(par);
          const uint32_t wide0 = wide;
          uint8_t buf[16];
          if (par) {
            const uint8_t *src = out + mypos - dist;
#pragma unroll
            for (int i = 0; i < 16; i++)
              if ((uint32_t)i < length) buf[i] = src[i];
          }
          uint32_t wlen[3], wpos[3];
          uint8_t wa[3], wb[3];
#pragma unroll
          for (int k = 0; k < 3; k++) {
            wlen[k] = 0;
            wpos[k] = 0;
            wa[k] = wb[k] = 0;
            if (wide) {  // uniform
              const int sl = __ffs((int)wide) - 1;
              if (tile.shfl(length, sl) <= 64u) {
                wide &= wide - 1;
                wlen[k] = tile.shfl(length, sl);
                wpos[k] = tile.shfl(mypos, sl);
                const uint8_t *src = out + wpos[k] - tile.shfl(dist, sl);
                if ((uint32_t)lane < wlen[k]) wa[k] = src[lane];
                if ((uint32_t)lane + 32u < wlen[k]) wb[k] = src[lane + 32];
              }
            }
          }
          if (par) {
            uint8_t *dst = out + mypos;
#pragma unroll
            for (int i = 0; i < 16; i++)
              if ((uint32_t)i < length) dst[i] = buf[i];
          }
          if (wide0) {  // uniform
#pragma unroll
            for (int k = 0; k < 3; k++) {
              uint8_t *dst = out + wpos[k];
              if ((uint32_t)lane < wlen[k]) dst[lane] = wa[k];
              if ((uint32_t)lane + 32u < wlen[k]) dst[lane + 32] = wb[k];
            }
          }
#pragma unroll 1
          while (wide) {  // ready, disjoint from their source, not yet copied: more than three or longer than 64 bytes
            const int sl = __ffs((int)wide) - 1;
            wide &= wide - 1;
            const uint32_t len = tile.shfl(length, sl), mp = tile.shfl(mypos, sl);
            const uint8_t *src = out + mp - tile.shfl(dist, sl);
            uint8_t *dst = out + mp;
            for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i];
          }
          pend &= ~rmask;
          tile.sync();
        }
      }
      outpos += total;
      if (!ZLIB && verify_crc) crc.advance(lane, out, outpos, crcT);  // the round's bytes are final and still cached
    }
  }
  if (ZLIB) {
    tile.sync();
    if (status == INF_OK) {
      // Adler-32 of the inflated bytes: a = 1 + sum b_i, b = n + sum (n - i) b_i, both mod 65521
      uint64_t sa = 0, sb = 0;
      for (uint32_t i = (uint32_t)lane; i < outpos; i += 32) {
        const uint64_t x = out[i];
        sa += x;
        sb += (uint64_t)(outpos - i) * x;
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        sa += tile.shfl_xor(sa, d);
        sb += tile.shfl_xor(sb, d);
      }
      const uint32_t adler = ((uint32_t)((outpos + sb) % 65521u) << 16) | (uint32_t)((1 + sa) % 65521u);
      if (verify_crc && adler != crc_expected) status = INF_CRC_MISMATCH;
      if (out_len && lane == 0) *out_len = outpos;
    }
    return status;
  }
  if (status == INF_OK && outpos != isize) status = INF_SIZE_MISMATCH;
  tile.sync();
  if (status == INF_OK && verify_crc && isize > 0) {
    const uint32_t c = crc.finish(tile, out, isize, crcT);
    if (c != crc_expected) status = INF_CRC_MISMATCH;
  }
  return status;
}

// CRC_SMEM: the CRC tables (5 KB) are copied into the CTA's shared memory; one-warp CTAs (30 per SM) cannot afford the copy and
// read them through L1 from global memory instead.  That shape was tried to let a finished warp hand its slot to the next
// segment's launch without waiting for its nine CTA neighbours: measured 8 % slower at 40 M pairs (124.5 vs 115.5 ms per pass;
// two-warp CTAs 121.7 ms), so blocks are not waiting for slots - the ten-warp shape stays (profiles/r2_sweeps.md).
template <int THREADS, int MIN_CTAS, int LB, int DB, int SPAN, int NT, bool CRC_SMEM, bool ZLIB = false>
__global__ void __maxnreg__(BND_K1_MAXREG) bgzf_inflate_spec_kernel(InflateArgs a) {
  using Smem = SpecSmemT<LB, DB, NT, SPAN>;
  BND_DYN_SMEM(smem_raw);
  constexpr int WARPS = THREADS / 32;
  Smem *warps = reinterpret_cast<Smem *>(smem_raw);
  SymTables *sym_s = reinterpret_cast<SymTables *>(smem_raw + sizeof(Smem) * WARPS);
  CrcTables *crc_s = reinterpret_cast<CrcTables *>(smem_raw + sizeof(Smem) * WARPS + sizeof(SymTables));
  if (threadIdx.x < 32) {
    const uint32_t s = threadIdx.x;
    const uint32_t eb = s < 8 ? 0u : s == 28 ? 0u : (s - 4) >> 2;
    const uint32_t lbase = s < 8 ? s + 3 : s == 28 ? 258u : ((4 + (s & 3)) << eb) + 3;
    sym_s->len_tab[s] = s < 29 ? (lbase | (eb << 16)) : 0u;
    const uint32_t deb = s < 4 ? 0u : (s >> 1) - 1;
    const uint32_t dbase = s < 4 ? s + 1 : ((2 + (s & 1)) << deb) + 1;
    sym_s->dist_tab[s] = s < 30 ? (dbase | (deb << 16)) : 0u;
  }
  if (CRC_SMEM && !ZLIB) {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(a.crc);
    uint32_t *dst = reinterpret_cast<uint32_t *>(crc_s);
    for (int i = threadIdx.x; i < (int)(sizeof(CrcTables) / 4); i += THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const CrcTables *crcT = CRC_SMEM ? crc_s : a.crc;
  Tile<32> tile;
  Smem &S = warps[threadIdx.x / 32];
  for (;;) {
    uint32_t b = 0;
    if (tile.lane == 0) b = atomicAdd(a.work_counter, 1u);
    b = tile.shfl(b, 0);
    if (b >= a.n_blocks) break;
    const uint32_t cs = ZLIB && a.coff_stride == 2 ? 2u : 1u;
    const uint64_t c0 = a.blk_coff[(uint64_t)b * cs], c1 = a.blk_coff[(uint64_t)b * cs + 1];
    const uint64_t u0 = a.blk_uoff[b], u1 = a.blk_uoff[b + 1];
    const int st = inflate_block_spec<Smem, ZLIB>(tile, S, *sym_s, a.comp + c0, (uint32_t)(c1 - c0), a.out + u0, (uint32_t)(u1 - u0), crcT, a.verify_crc,
                                                  ZLIB && a.out_len ? a.out_len + b : nullptr);
    if (tile.lane == 0) {
      if (a.status) a.status[b] = st;
      if (st != INF_OK) atomicCAS(a.error_flag, 0, st | ((int)(b & 0xffffff) << 8));
    }
    tile.sync();
  }
}

template <int LB, int DB, int NT, int SPAN>
constexpr size_t inflate_spec_smem_bytes(int warps, bool crc_smem) {
  return sizeof(SpecSmemT<LB, DB, NT, SPAN>) * (size_t)warps + sizeof(SymTables) + (crc_smem ? sizeof(CrcTables) : 0);
}

}  // namespace bnd
