// synth_bam.cpp — seeded synthetic coordinate-sorted BAM + BAI generator (tests and bench input).
//
// Produces the input shapes SURVEY.md §8(d) names for configs 2-5: paired-end (or single-end) reads over an
// hg38-like header, insert-size / flag / MAPQ mixtures, optional ATAC insert mixture, optional CB barcodes, and
// optional "odd" records (MAPQ 255, zero-span CIGARs, TLEN 0, long N-skips crossing 5 Mb chunk boundaries,
// unmapped-but-placed mates).  Output is deterministic for a given (seed, parameters) regardless of thread
// count: the genome is cut into fixed tiles, every pair has its own RNG stream, records are sorted per tile.
// BGZF blocks are cut every 65280 bytes of the record stream like htslib and compressed with zlib.
//
// This is original code; it reads nothing from the reference.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Contig {
  std::string name;
  uint64_t len;
};

const Contig HG38[] = {
    {"chr1", 248956422}, {"chr2", 242193529}, {"chr3", 198295559}, {"chr4", 190214555}, {"chr5", 181538259},
    {"chr6", 170805979}, {"chr7", 159345973}, {"chr8", 145138636}, {"chr9", 138394717}, {"chr10", 133797422},
    {"chr11", 135086622}, {"chr12", 133275309}, {"chr13", 114364328}, {"chr14", 107043718}, {"chr15", 101991189},
    {"chr16", 90338345}, {"chr17", 83257441}, {"chr18", 80373285}, {"chr19", 58617616}, {"chr20", 64444167},
    {"chr21", 46709983}, {"chr22", 50818468}, {"chrX", 156040895}, {"chrY", 57227415}, {"chrM", 16569},
    {"chr1_KI270706v1_random", 175055}, {"chr14_GL000009v2_random", 201709}, {"chrUn_KI270742v1", 186739},
    {"chrUn_GL000195v1", 182896}, {"chr6_GL000250v2_alt", 4672374}, {"chr19_KI270938v1_alt", 1066800},
};

struct Rng {  // xoshiro256** seeded by splitmix64
  uint64_t s[4];
  static uint64_t splitmix(uint64_t &x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) {
    for (auto &v : s) v = splitmix(seed);
  }
  static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return r;
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint64_t below(uint64_t n) { return (uint64_t)(uni() * (double)n); }
  double gauss() {
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};

uint64_t mix(uint64_t a, uint64_t b) {
  uint64_t x = a ^ (b + 0x9e3779b97f4a7c15ull + (a << 6) + (a >> 2));
  return Rng::splitmix(x);
}

struct Params {
  std::string out;
  uint64_t pairs = 100000;
  uint64_t seed = 20260219;
  std::string mode = "pe";  // pe | atac | se
  bool barcodes = false;    // CB:Z tags (config 4)
  bool odd = false;         // adversarial records
  bool rg = false;          // RG:Z tags (two read groups) + VP:Z tag
  int threads = 0;
  int level = 6;
  uint32_t read_len = 100;
  uint64_t n_peaks = 50000;
  uint64_t n_barcodes = 50000;
  uint64_t n_unplaced = 0;
  uint64_t tile = 1000000;
  std::vector<Contig> contigs;
};

struct RecMeta {
  int32_t ref, pos, end;  // 0-based [pos,end) used for binning (end = pos+1 when span is 0)
  uint32_t size;          // bytes including block_size
  uint16_t flag;
};

int reg2bin(int64_t beg, int64_t end) {
  --end;
  if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

void put16(std::vector<uint8_t> &b, uint32_t v) {
  b.push_back((uint8_t)v);
  b.push_back((uint8_t)(v >> 8));
}
void put32(std::vector<uint8_t> &b, uint32_t v) {
  for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i)));
}
void put64(std::vector<uint8_t> &b, uint64_t v) {
  for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i)));
}

struct Read {  // one alignment record before serialisation
  int32_t ref, pos;  // 0-based
  uint8_t mapq;
  uint16_t flag;
  int32_t next_ref, next_pos, tlen;
  std::vector<uint32_t> cigar;
  uint64_t pair_id;
  uint64_t rseed;  // seed for seq/qual/aux
  uint32_t l_seq;
  int64_t barcode = -1;  // barcode index or -1
  int rg = -1;
};

struct Genome {
  std::vector<Contig> contigs;
  std::vector<uint64_t> cum;  // cumulative start of each contig in the concatenated genome
  uint64_t total = 0;
  // tiles
  struct Tile {
    int ref;
    uint64_t start, end;  // 0-based [start,end) within contig
    uint64_t gstart;      // start in concatenated coordinates
  };
  std::vector<Tile> tiles;
  std::vector<size_t> first_tile;  // per contig
  void build(uint64_t tile) {
    cum.clear();
    total = 0;
    for (auto &c : contigs) {
      cum.push_back(total);
      total += c.len;
    }
    tiles.clear();
    first_tile.clear();
    for (size_t r = 0; r < contigs.size(); r++) {
      first_tile.push_back(tiles.size());
      for (uint64_t s = 0; s < contigs[r].len; s += tile)
        tiles.push_back(Tile{(int)r, s, std::min(s + tile, contigs[r].len), cum[r] + s});
    }
  }
};

struct Gen {
  Params P;
  Genome G;
  std::vector<uint64_t> peak_pos;        // concatenated coordinate of each peak centre, sorted
  std::vector<size_t> peak_first;        // per tile: first peak index
  std::vector<uint64_t> uni_first;       // per tile: first uniform pair index (prefix)
  uint64_t n_uniform = 0, n_peak_pairs = 0, pairs_per_peak = 0, peak_extra = 0;

  void init() {
    G.contigs = P.contigs;
    G.build(P.tile);
    uint64_t n_peaks = std::min<uint64_t>(P.n_peaks, std::max<uint64_t>(P.pairs / 20, 1));
    n_peak_pairs = (uint64_t)((double)P.pairs * 0.15);
    if (n_peaks == 0) n_peak_pairs = 0;
    n_uniform = P.pairs - n_peak_pairs;
    pairs_per_peak = n_peaks ? n_peak_pairs / n_peaks : 0;
    peak_extra = n_peaks ? n_peak_pairs % n_peaks : 0;
    Rng r(mix(P.seed, 0x7065616b73ull));
    peak_pos.resize(n_peaks);
    for (auto &p : peak_pos) p = r.below(G.total);
    std::sort(peak_pos.begin(), peak_pos.end());
    peak_first.assign(G.tiles.size() + 1, 0);
    size_t pi = 0;
    for (size_t t = 0; t < G.tiles.size(); t++) {
      peak_first[t] = pi;
      uint64_t gend = G.tiles[t].gstart + (G.tiles[t].end - G.tiles[t].start);
      while (pi < peak_pos.size() && peak_pos[pi] < gend) pi++;
    }
    peak_first[G.tiles.size()] = peak_pos.size();
    // uniform pairs per tile proportional to tile length (deterministic floor of the cumulative share)
    uni_first.assign(G.tiles.size() + 1, 0);
    for (size_t t = 0; t <= G.tiles.size(); t++) {
      uint64_t g = t < G.tiles.size() ? G.tiles[t].gstart : G.total;
      uni_first[t] = (uint64_t)((long double)n_uniform * (long double)g / (long double)G.total);
    }
    uni_first[G.tiles.size()] = n_uniform;
  }

  // pair ids: uniform pairs first [0, n_uniform), then peak pairs
  uint64_t peak_pair_base(size_t peak) const { return n_uniform + peak * pairs_per_peak + std::min<uint64_t>(peak, peak_extra); }
  uint64_t peak_pair_count(size_t peak) const { return pairs_per_peak + (peak < peak_extra ? 1 : 0); }

  // Generate the (up to two) reads of one pair whose anchor is `anchor` (0-based contig coordinate, may be <0).
  void make_pair(uint64_t pair_id, int ref, int64_t anchor, std::vector<Read> &out) const {
    Rng r(mix(P.seed ^ 0x70616972ull, pair_id));
    const uint64_t clen = G.contigs[ref].len;
    const uint32_t L = P.read_len;
    bool single = P.mode == "se";
    // insert size
    int64_t isize;
    if (P.mode == "atac") {
      double u = r.uni();
      double v = u < 0.45 ? 80 + 20 * r.gauss() : u < 0.80 ? 190 + 25 * r.gauss() : 380 + 40 * r.gauss();
      isize = (int64_t)std::llround(v);
      isize = std::max<int64_t>(30, std::min<int64_t>(1000, isize));
    } else {
      isize = (int64_t)std::llround(300 + 60 * r.gauss());
      isize = std::max<int64_t>(100, std::min<int64_t>(1000, isize));
    }
    uint32_t l1 = L, l2 = L;
    if (P.mode == "atac" && isize < (int64_t)L) l1 = l2 = (uint32_t)isize;  // short fragments: reads trimmed to the insert
    int64_t pos1 = anchor;
    if (pos1 < 0) pos1 = 0;
    if ((uint64_t)pos1 + (uint64_t)std::max<int64_t>(isize, L) + 2 >= clen) {
      if (clen < (uint64_t)std::max<int64_t>(isize, L) + 4) return;  // contig too small for a pair
      pos1 = (int64_t)clen - std::max<int64_t>(isize, L) - 3;
    }
    int64_t pos2 = pos1 + isize - (int64_t)l2;
    // category
    double cu = r.uni();
    enum { PROPER, IMPROPER, MATE_UNMAPPED, DUP, SECONDARY, SUPPL } cat =
        cu < 0.92 ? PROPER : cu < 0.97 ? IMPROPER : cu < 0.98 ? MATE_UNMAPPED : cu < 0.99 ? DUP : cu < 0.995 ? SECONDARY : SUPPL;
    bool first_fwd = r.uni() < 0.5;  // 99/147 vs 83/163 orientation of the leftmost read
    auto draw_mapq = [&]() -> uint8_t {
      double u = r.uni();
      if (u < 0.70) return 60;
      if (u < 0.85) return (uint8_t)(30 + r.below(30));
      if (u < 0.95) return (uint8_t)(1 + r.below(29));
      if (u < 0.999) return 0;
      return 255;
    };
    auto draw_cigar = [&](uint32_t l, std::vector<uint32_t> &cg) {
      double u = r.uni();
      cg.clear();
      if (u < 0.95 || l < 30) {
        cg.push_back(l << 4 | 0);
        return;
      }
      uint32_t a = 10 + (uint32_t)r.below(l - 25), k = 1 + (uint32_t)r.below(5);
      double v = r.uni();
      if (v < 0.34) {  // insertion: aM kI (l-a-k)M
        cg = {a << 4 | 0, k << 4 | 1, (l - a - k) << 4 | 0};
      } else if (v < 0.67) {  // deletion: aM kD (l-a)M
        cg = {a << 4 | 0, k << 4 | 2, (l - a) << 4 | 0};
      } else {  // soft clip: kS (l-k)M  (two ops + trailing) -> kS aM (l-k-a)M merged is unusual; use kS (l-2k)M kS
        cg = {k << 4 | 4, (l - 2 * k) << 4 | 0, k << 4 | 4};
      }
    };
    int64_t bc = -1;
    if (P.barcodes && r.uni() < 0.97) {  // Zipf(1.1)-like rank draw over n_barcodes
      double u = r.uni();
      double n = (double)P.n_barcodes;
      double x = std::pow(1.0 + u * (std::pow(n, -0.1) - 1.0), -10.0);  // inverse CDF of a continuous power law s=1.1
      bc = (int64_t)std::min<double>(n - 1, std::floor(x - 1.0 + 1e-9));
      if (bc < 0) bc = 0;
    }
    int rg = P.rg ? (int)r.below(3) - 1 : -1;  // -1: no RG tag, 0/1: two groups
    if (!P.rg) rg = -1;

    Read a, b;
    a.ref = b.ref = ref;
    a.pair_id = b.pair_id = pair_id;
    a.barcode = b.barcode = bc;
    a.rg = b.rg = rg;
    a.l_seq = l1;
    b.l_seq = l2;
    a.pos = (int32_t)pos1;
    b.pos = (int32_t)pos2;
    a.mapq = draw_mapq();
    b.mapq = draw_mapq();
    draw_cigar(l1, a.cigar);
    draw_cigar(l2, b.cigar);
    a.rseed = r.next();
    b.rseed = r.next();
    if (single) {
      a.flag = first_fwd ? 0 : 16;
      if (cat == DUP) a.flag |= 0x400;
      if (cat == SECONDARY) a.flag |= 0x100;
      if (cat == SUPPL) a.flag |= 0x800;
      a.next_ref = -1;
      a.next_pos = -1;
      a.tlen = 0;
      if (P.odd) oddify(r, a, nullptr, clen);
      out.push_back(a);
      return;
    }
    // paired: leftmost read a (forward if first_fwd: flags 99 / 147, else 163 / 83)
    a.flag = first_fwd ? 99 : 163;
    b.flag = first_fwd ? 147 : 83;
    a.next_ref = b.next_ref = ref;
    a.next_pos = b.pos;
    b.next_pos = a.pos;
    a.tlen = (int32_t)isize;
    b.tlen = -(int32_t)isize;
    switch (cat) {
      case PROPER: break;
      case IMPROPER:
        a.flag &= ~0x2;
        b.flag &= ~0x2;
        break;
      case MATE_UNMAPPED:  // b unmapped, placed at a's coordinates
        a.flag = (uint16_t)((a.flag & ~0x2 & ~0x20) | 0x8);
        b.flag = (uint16_t)(0x1 | 0x4 | (b.flag & 0xC0) | (a.flag & 0x10 ? 0x20 : 0));
        b.pos = a.pos;
        b.mapq = 0;
        b.cigar.clear();
        a.next_pos = a.pos;
        b.next_pos = a.pos;
        a.tlen = b.tlen = 0;
        break;
      case DUP:
        a.flag |= 0x400;
        b.flag |= 0x400;
        break;
      case SECONDARY: a.flag |= 0x100; break;
      case SUPPL: b.flag |= 0x800; break;
    }
    if (P.odd) oddify(r, a, &b, clen);
    out.push_back(a);
    out.push_back(b);
  }

  // adversarial edits (small probabilities)
  void oddify(Rng &r, Read &a, Read *b, uint64_t clen) const {
    double u = r.uni();
    if (u < 0.01) {  // zero-span CIGAR on a mapped read (all insertion)
      a.cigar = {a.l_seq << 4 | 1};
    } else if (u < 0.02 && b) {  // TLEN 0 on a proper pair
      a.tlen = 0;
      b->tlen = 0;
    } else if (u < 0.03) {  // long N skip (spans up to ~2 chunk lengths)
      uint32_t skip = 1000 + (uint32_t)r.below(7000000);
      if ((uint64_t)a.pos + a.l_seq + skip + 10 < clen && a.l_seq >= 20) {
        a.cigar = {10u << 4 | 0, skip << 4 | 3, (a.l_seq - 10) << 4 | 0};
      }
    } else if (u < 0.04) {
      a.mapq = 255;
    } else if (u < 0.05) {  // = / X ops and hard clip + padding
      if (a.l_seq >= 20) a.cigar = {5u << 4 | 5, 10u << 4 | 7, 2u << 4 | 8, (a.l_seq - 12) << 4 | 7, 3u << 4 | 6};
    } else if (u < 0.06 && b) {  // mate on another contig (next_ref differs), tlen 0
      b->next_ref = a.next_ref = (a.ref + 1) % (int)G.contigs.size();
      a.tlen = b->tlen = 0;
      a.flag &= ~0x2;
      b->flag &= ~0x2;
    } else if (u < 0.065) {
      a.l_seq = 0;  // no sequence stored ('*'), CIGAR kept
    }
  }

  void serialise(const Read &rd, std::vector<uint8_t> &buf, std::vector<RecMeta> &meta) const {
    Rng r(rd.rseed);
    char name[32];
    int ln = snprintf(name, sizeof name, "r%09llu", (unsigned long long)rd.pair_id) + 1;
    uint64_t span = 0;
    for (uint32_t c : rd.cigar) {
      uint32_t op = c & 0xf;
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += c >> 4;
    }
    int64_t end = rd.pos + (int64_t)(span ? span : 1);
    size_t start_off = buf.size();
    put32(buf, 0);  // block_size placeholder
    put32(buf, (uint32_t)rd.ref);
    put32(buf, (uint32_t)rd.pos);
    buf.push_back((uint8_t)ln);
    buf.push_back(rd.mapq);
    put16(buf, (uint32_t)reg2bin(rd.pos, end));
    put16(buf, (uint32_t)rd.cigar.size());
    put16(buf, rd.flag);
    put32(buf, rd.l_seq);
    put32(buf, (uint32_t)rd.next_ref);
    put32(buf, (uint32_t)rd.next_pos);
    put32(buf, (uint32_t)rd.tlen);
    buf.insert(buf.end(), name, name + ln);
    for (uint32_t c : rd.cigar) put32(buf, c);
    // sequence: random bases, 4-bit packed
    static const uint8_t code[4] = {1, 2, 4, 8};
    for (uint32_t i = 0; i < (rd.l_seq + 1) / 2; i++) {
      uint64_t x = r.next();
      uint8_t hi = code[x & 3], lo = (2 * i + 1 < rd.l_seq) ? code[(x >> 2) & 3] : 0;
      buf.push_back((uint8_t)(hi << 4 | lo));
    }
    // qualities: 4-level binned alphabet with long runs
    static const uint8_t ql[4] = {37, 23, 12, 2};
    uint32_t i = 0;
    while (i < rd.l_seq) {
      uint64_t x = r.next();
      uint8_t q = ql[(x & 15) < 11 ? 0 : (x & 15) < 14 ? 1 : (x & 15) < 15 ? 2 : 3];
      uint32_t run = 1 + (uint32_t)((x >> 8) % 40);
      for (uint32_t k = 0; k < run && i < rd.l_seq; k++, i++) buf.push_back(q);
    }
    // aux: NM:C AS:C MD:Z [CB:Z] [RG:Z VP:Z]
    uint64_t x = r.next();
    buf.push_back('N'); buf.push_back('M'); buf.push_back('C'); buf.push_back((uint8_t)(x % 4));
    buf.push_back('A'); buf.push_back('S'); buf.push_back('C'); buf.push_back((uint8_t)(rd.l_seq > 255 ? 255 : rd.l_seq - (x >> 3) % 8));
    char md[16];
    int lm = snprintf(md, sizeof md, "%u", rd.l_seq) + 1;
    buf.push_back('M'); buf.push_back('D'); buf.push_back('Z');
    buf.insert(buf.end(), md, md + lm);
    if (rd.barcode >= 0) {
      static const char B[4] = {'A', 'C', 'G', 'T'};
      uint64_t h = mix(P.seed ^ 0x6263ull, (uint64_t)rd.barcode);
      buf.push_back('C'); buf.push_back('B'); buf.push_back('Z');
      for (int k = 0; k < 16; k++) buf.push_back((uint8_t)B[(h >> (2 * k)) & 3]);
      buf.push_back('-'); buf.push_back('1'); buf.push_back(0);
    }
    if (P.rg) {
      // an integer-typed tag before RG so that the aux walk has to skip mixed types
      buf.push_back('X'); buf.push_back('S'); buf.push_back('i'); put32(buf, (uint32_t)(x >> 20));
      buf.push_back('X'); buf.push_back('B'); buf.push_back('B'); buf.push_back('S'); put32(buf, 2); put16(buf, 7); put16(buf, 9);
      if (rd.rg >= 0) {
        buf.push_back('R'); buf.push_back('G'); buf.push_back('Z');
        const char *g = rd.rg == 0 ? "grpA" : "grpB";
        buf.insert(buf.end(), g, g + 5);
      }
      const char *vp = (x >> 40) % 3 == 0 ? "BCL2" : (x >> 40) % 3 == 1 ? "MYC" : nullptr;
      if (vp) {
        buf.push_back('V'); buf.push_back('P'); buf.push_back('Z');
        buf.insert(buf.end(), vp, vp + strlen(vp) + 1);
      } else if ((x >> 44) & 1) {  // VP present but as an integer (type mismatch case)
        buf.push_back('V'); buf.push_back('P'); buf.push_back('C'); buf.push_back(5);
      }
    }
    uint32_t bs = (uint32_t)(buf.size() - start_off - 4);
    memcpy(&buf[start_off], &bs, 4);
    meta.push_back(RecMeta{rd.ref, rd.pos, (int32_t)end, bs + 4, rd.flag});
  }

  // all records whose pos falls in tile t, sorted by (pos, pair_id, flag)
  void make_tile(size_t t, std::vector<uint8_t> &buf, std::vector<RecMeta> &meta) const {
    const auto &T = G.tiles[t];
    std::vector<Read> reads, cand;
    for (int d = -1; d <= 1; d++) {
      int64_t tt = (int64_t)t + d;
      if (tt < 0 || tt >= (int64_t)G.tiles.size()) continue;
      const auto &S = G.tiles[(size_t)tt];
      if (S.ref != T.ref) continue;
      uint64_t slen = S.end - S.start;
      // uniform pairs anchored in tile S
      for (uint64_t id = uni_first[tt]; id < uni_first[tt + 1]; id++) {
        Rng r(mix(P.seed ^ 0x616e63ull, id));
        int64_t anchor = (int64_t)(S.start + r.below(slen));
        if (anchor + 1200 < (int64_t)T.start || anchor > (int64_t)T.end + 1200) continue;
        cand.clear();
        make_pair(id, S.ref, anchor, cand);
        for (auto &rd : cand)
          if ((uint64_t)rd.pos >= T.start && (uint64_t)rd.pos < T.end) reads.push_back(rd);
      }
      // peak pairs anchored at peaks of tile S
      for (size_t pk = peak_first[tt]; pk < peak_first[tt + 1]; pk++) {
        int64_t centre = (int64_t)(peak_pos[pk] - G.cum[S.ref]);
        uint64_t base = peak_pair_base(pk), cnt = peak_pair_count(pk);
        for (uint64_t k = 0; k < cnt; k++) {
          uint64_t id = base + k;
          Rng r(mix(P.seed ^ 0x616e63ull, id));
          int64_t anchor = centre + (int64_t)std::llround(150.0 * r.gauss());
          cand.clear();
          make_pair(id, S.ref, anchor, cand);
          for (auto &rd : cand)
            if ((uint64_t)rd.pos >= T.start && (uint64_t)rd.pos < T.end) reads.push_back(rd);
        }
      }
    }
    std::sort(reads.begin(), reads.end(), [](const Read &a, const Read &b) {
      if (a.pos != b.pos) return a.pos < b.pos;
      if (a.pair_id != b.pair_id) return a.pair_id < b.pair_id;
      return a.flag < b.flag;
    });
    buf.clear();
    meta.clear();
    buf.reserve(reads.size() * 230);
    for (auto &rd : reads) serialise(rd, buf, meta);
  }
};

// ---- BGZF ------------------------------------------------------------------------------------------------
const uint32_t BLOCK_DATA = 0xff00;

void bgzf_compress(const uint8_t *src, uint32_t n, int level, std::vector<uint8_t> &out) {
  out.resize(18 + compressBound(n) + 8 + 64);
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) abort();
  zs.next_in = const_cast<uint8_t *>(src);
  zs.avail_in = n;
  zs.next_out = out.data() + 18;
  zs.avail_out = (uInt)(out.size() - 18 - 8);
  if (deflate(&zs, Z_FINISH) != Z_STREAM_END) abort();
  uint32_t clen = (uint32_t)zs.total_out;
  deflateEnd(&zs);
  uint32_t total = 18 + clen + 8;
  if (total > 65536) {  // incompressible: store
    if (deflateInit2(&zs, 0, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) abort();
    zs.next_in = const_cast<uint8_t *>(src);
    zs.avail_in = n;
    zs.next_out = out.data() + 18;
    zs.avail_out = (uInt)(out.size() - 18 - 8);
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) abort();
    clen = (uint32_t)zs.total_out;
    deflateEnd(&zs);
    total = 18 + clen + 8;
  }
  static const uint8_t H[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
  memcpy(out.data(), H, 12);
  out[12] = 'B';
  out[13] = 'C';
  out[14] = 2;
  out[15] = 0;
  out[16] = (uint8_t)((total - 1) & 0xff);
  out[17] = (uint8_t)((total - 1) >> 8);
  uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, n);
  memcpy(out.data() + 18 + clen, &crc, 4);
  memcpy(out.data() + 18 + clen + 4, &n, 4);
  out.resize(total);
}

struct BaiBuilder {
  struct RefIdx {
    std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
    std::vector<uint64_t> linear;
    uint64_t off_beg = UINT64_MAX, off_end = 0, n_mapped = 0, n_unmapped = 0;
  };
  std::vector<RefIdx> refs;
  uint64_t n_no_coor = 0;
  int last_ref = -2;
  uint32_t last_bin = UINT32_MAX;
  explicit BaiBuilder(size_t n) : refs(n) {}
  void add(const RecMeta &m, uint64_t vbeg, uint64_t vend) {
    if (m.ref < 0) {
      n_no_coor++;
      return;
    }
    RefIdx &R = refs[m.ref];
    uint32_t bin = (uint32_t)reg2bin(m.pos, m.end);
    if (m.ref != last_ref || bin != last_bin) {
      R.bins[bin].push_back({vbeg, vend});
      last_ref = m.ref;
      last_bin = bin;
    } else {
      R.bins[bin].back().second = vend;
    }
    size_t w0 = (size_t)(m.pos >> 14), w1 = (size_t)((m.end - 1) >> 14);
    if (R.linear.size() <= w1) R.linear.resize(w1 + 1, UINT64_MAX);
    for (size_t w = w0; w <= w1; w++)
      if (R.linear[w] == UINT64_MAX) R.linear[w] = vbeg;
    R.off_beg = std::min(R.off_beg, vbeg);
    R.off_end = std::max(R.off_end, vend);
    if (m.flag & 4)
      R.n_unmapped++;
    else
      R.n_mapped++;
  }
  void write(const std::string &path) {
    std::vector<uint8_t> b;
    b.insert(b.end(), {'B', 'A', 'I', 1});
    put32(b, (uint32_t)refs.size());
    for (auto &R : refs) {
      bool any = R.off_beg != UINT64_MAX;
      put32(b, (uint32_t)(R.bins.size() + (any ? 1 : 0)));
      for (auto &kv : R.bins) {
        put32(b, kv.first);
        put32(b, (uint32_t)kv.second.size());
        for (auto &c : kv.second) {
          put64(b, c.first);
          put64(b, c.second);
        }
      }
      if (any) {
        put32(b, 37450);
        put32(b, 2);
        put64(b, R.off_beg);
        put64(b, R.off_end);
        put64(b, R.n_mapped);
        put64(b, R.n_unmapped);
      }
      for (size_t i = R.linear.size(); i-- > 1;)  // back-fill empty windows with the next valid offset
        if (R.linear[i - 1] == UINT64_MAX) R.linear[i - 1] = R.linear[i];
      put32(b, (uint32_t)R.linear.size());
      for (uint64_t v : R.linear) put64(b, v);
    }
    put64(b, n_no_coor);
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) {
      perror("bai");
      exit(1);
    }
    fwrite(b.data(), 1, b.size(), f);
    fclose(f);
  }
};

template <class Fn>
void parallel_for(size_t n, int threads, Fn fn) {
  std::atomic<size_t> next{0};
  auto work = [&] {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= n) break;
      fn(i);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < threads; t++) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
}

int run(Params P) {
  if (P.threads <= 0) P.threads = (int)std::thread::hardware_concurrency();
  if (P.threads < 1) P.threads = 1;
  Gen gen;
  gen.P = P;
  gen.init();
  const Genome &G = gen.G;

  FILE *fo = fopen(P.out.c_str(), "wb");
  if (!fo) {
    perror("open output");
    return 1;
  }
  uint64_t coff = 0;
  // header
  {
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (auto &c : G.contigs) text += "@SQ\tSN:" + c.name + "\tLN:" + std::to_string(c.len) + "\n";
    text += "@PG\tID:synth_bam\tPN:synth_bam\tVN:1\n";
    std::vector<uint8_t> h;
    h.insert(h.end(), {'B', 'A', 'M', 1});
    put32(h, (uint32_t)text.size());
    h.insert(h.end(), text.begin(), text.end());
    put32(h, (uint32_t)G.contigs.size());
    for (auto &c : G.contigs) {
      put32(h, (uint32_t)c.name.size() + 1);
      h.insert(h.end(), c.name.begin(), c.name.end());
      h.push_back(0);
      put32(h, (uint32_t)c.len);
    }
    std::vector<uint8_t> blk;
    for (size_t o = 0; o < h.size(); o += BLOCK_DATA) {
      bgzf_compress(h.data() + o, (uint32_t)std::min<size_t>(BLOCK_DATA, h.size() - o), P.level, blk);
      fwrite(blk.data(), 1, blk.size(), fo);
      coff += blk.size();
    }
  }

  BaiBuilder bai(G.contigs.size());
  std::vector<uint8_t> stream;  // pending uncompressed record bytes; stream[0] is at a block boundary
  uint64_t n_records = 0, n_ubytes = 0, n_blocks = 0;
  size_t n_tiles = G.tiles.size();
  size_t batch = (size_t)P.threads * 2;

  // Records are indexed with absolute stream offsets: a table of block coffsets for the whole file (8 B per
  // 64 KiB block, 6 MB for a 12 GB BAM) resolves them to virtual offsets once the holding blocks are written.
  std::vector<uint64_t> block_coff;  // coffset of record-stream block i
  uint64_t stream_abs = 0;           // absolute offset of stream[0]
  struct PendRec {
    RecMeta m;
    uint64_t abs;
  };
  std::vector<PendRec> pend;
  size_t pend_head = 0;
  uint64_t final_coff_end = 0;
  auto resolve = [&](uint64_t abs_off) -> uint64_t {
    uint64_t bi = abs_off / BLOCK_DATA, bo = abs_off % BLOCK_DATA;
    if (bi >= block_coff.size()) return final_coff_end << 16;
    return (block_coff[bi] << 16) | bo;
  };
  auto flush2 = [&](bool final) {
    size_t nfull = stream.size() / BLOCK_DATA;
    size_t nblk = nfull + ((final && stream.size() % BLOCK_DATA) ? 1 : 0);
    if (nblk) {
      std::vector<std::vector<uint8_t>> comp(nblk);
      parallel_for(nblk, P.threads, [&](size_t i) {
        size_t o = i * BLOCK_DATA;
        bgzf_compress(stream.data() + o, (uint32_t)std::min<size_t>(BLOCK_DATA, stream.size() - o), P.level, comp[i]);
      });
      for (size_t i = 0; i < nblk; i++) {
        block_coff.push_back(coff);
        fwrite(comp[i].data(), 1, comp[i].size(), fo);
        coff += comp[i].size();
      }
      n_blocks += nblk;
      size_t consumed = std::min(stream.size(), nblk * (size_t)BLOCK_DATA);
      stream.erase(stream.begin(), stream.begin() + (ptrdiff_t)consumed);
      stream_abs += consumed;
    }
    final_coff_end = coff;
    // index every pending record that ends within the compressed region (or all of them when final)
    while (pend_head < pend.size()) {
      const PendRec &pr = pend[pend_head];
      uint64_t e = pr.abs + pr.m.size;
      if (!final && e >= stream_abs) break;  // need the block after the end to exist unless final
      bai.add(pr.m, resolve(pr.abs), resolve(e));
      pend_head++;
    }
    if (pend_head > (1u << 20)) {
      pend.erase(pend.begin(), pend.begin() + (ptrdiff_t)pend_head);
      pend_head = 0;
    }
  };

  std::vector<std::vector<uint8_t>> tbuf(batch);
  std::vector<std::vector<RecMeta>> tmeta(batch);
  for (size_t t0 = 0; t0 < n_tiles; t0 += batch) {
    size_t nb = std::min(batch, n_tiles - t0);
    parallel_for(nb, P.threads, [&](size_t i) { gen.make_tile(t0 + i, tbuf[i], tmeta[i]); });
    for (size_t i = 0; i < nb; i++) {
      uint64_t abs = stream_abs + stream.size();
      for (auto &m : tmeta[i]) {
        pend.push_back(PendRec{m, abs});
        abs += m.size;
      }
      n_records += tmeta[i].size();
      n_ubytes += tbuf[i].size();
      stream.insert(stream.end(), tbuf[i].begin(), tbuf[i].end());
    }
    flush2(false);
  }
  // unplaced unmapped reads at the end
  if (P.n_unplaced) {
    std::vector<uint8_t> buf;
    std::vector<RecMeta> meta;
    for (uint64_t i = 0; i < P.n_unplaced; i++) {
      Read rd;
      rd.ref = -1;
      rd.pos = -1;
      rd.mapq = 0;
      rd.flag = 4;
      rd.next_ref = -1;
      rd.next_pos = -1;
      rd.tlen = 0;
      rd.pair_id = P.pairs + i;
      rd.rseed = mix(P.seed, P.pairs + i);
      rd.l_seq = P.read_len;
      gen.serialise(rd, buf, meta);
    }
    uint64_t abs = stream_abs + stream.size();
    for (auto &m : meta) {
      pend.push_back(PendRec{m, abs});
      abs += m.size;
    }
    n_records += meta.size();
    n_ubytes += buf.size();
    stream.insert(stream.end(), buf.begin(), buf.end());
  }
  flush2(true);
  // EOF block
  static const uint8_t EOFB[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  fwrite(EOFB, 1, 28, fo);
  coff += 28;
  fclose(fo);
  bai.write(P.out + ".bai");
  printf("{\"records\": %llu, \"inflated_bytes\": %llu, \"file_bytes\": %llu, \"blocks\": %llu, \"contigs\": %zu}\n",
         (unsigned long long)n_records, (unsigned long long)n_ubytes, (unsigned long long)coff,
         (unsigned long long)n_blocks, G.contigs.size());
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  Params P;
  std::string genome = "hg38";
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&]() -> std::string {
      if (i + 1 >= argc) {
        fprintf(stderr, "missing value for %s\n", a.c_str());
        exit(2);
      }
      return argv[++i];
    };
    if (a == "--out") P.out = val();
    else if (a == "--pairs") P.pairs = strtoull(val().c_str(), nullptr, 10);
    else if (a == "--seed") P.seed = strtoull(val().c_str(), nullptr, 10);
    else if (a == "--mode") P.mode = val();
    else if (a == "--barcodes") P.barcodes = true;
    else if (a == "--odd") P.odd = true;
    else if (a == "--rg") P.rg = true;
    else if (a == "--threads") P.threads = atoi(val().c_str());
    else if (a == "--level") P.level = atoi(val().c_str());
    else if (a == "--read-len") P.read_len = (uint32_t)atoi(val().c_str());
    else if (a == "--peaks") P.n_peaks = strtoull(val().c_str(), nullptr, 10);
    else if (a == "--n-barcodes") P.n_barcodes = strtoull(val().c_str(), nullptr, 10);
    else if (a == "--unplaced") P.n_unplaced = strtoull(val().c_str(), nullptr, 10);
    else if (a == "--tile") P.tile = strtoull(val().c_str(), nullptr, 10);
    else if (a == "--genome") genome = val();
    else if (a == "--contigs") {  // name:len,name:len
      std::string s = val();
      size_t p = 0;
      while (p < s.size()) {
        size_t c = s.find(',', p);
        if (c == std::string::npos) c = s.size();
        std::string item = s.substr(p, c - p);
        size_t k = item.find(':');
        P.contigs.push_back(Contig{item.substr(0, k), strtoull(item.c_str() + k + 1, nullptr, 10)});
        p = c + 1;
      }
    } else {
      fprintf(stderr, "unknown argument %s\n", a.c_str());
      return 2;
    }
  }
  if (P.out.empty()) {
    fprintf(stderr, "usage: synth_bam --out X.bam [--pairs N] [--seed S] [--mode pe|atac|se] [--barcodes] [--odd] [--rg]\n"
                    "                 [--genome hg38|small] [--contigs name:len,...] [--threads T] [--level Z]\n");
    return 2;
  }
  if (P.contigs.empty()) {
    if (genome == "hg38") {
      for (auto &c : HG38) P.contigs.push_back(c);
    } else if (genome == "small") {  // chr20-22 + chrM + two scaffolds (the reduced genome of config 5)
      for (auto &c : HG38)
        if (c.name == "chr20" || c.name == "chr21" || c.name == "chr22" || c.name == "chrM" ||
            c.name == "chrUn_KI270742v1" || c.name == "chr1_KI270706v1_random")
          P.contigs.push_back(c);
    } else {
      fprintf(stderr, "unknown genome %s\n", genome.c_str());
      return 2;
    }
  }
  return run(P);
}
