// bamio.cpp — see bamio.h.
#include "bamio.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace bnd {

namespace {
inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }
}  // namespace

Bai read_bai(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) fail("BAM index file does not exist: " + path);
  std::vector<uint8_t> d;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  d.resize((size_t)std::max<long>(sz, 0));
  bool ok = d.empty() || fread(d.data(), 1, d.size(), f) == d.size();
  fclose(f);
  if (!ok) fail("short read of " + path);
  size_t p = 0;
  auto need = [&](size_t n) {
    if (p + n > d.size()) fail("truncated BAM index: " + path);
  };
  need(8);
  if (memcmp(d.data(), "BAI\1", 4) != 0) fail("not a BAI index: " + path);
  const uint32_t n_ref = rd32(&d[4]);
  p = 8;
  Bai bai;
  bai.refs.resize(n_ref);
  for (uint32_t r = 0; r < n_ref; r++) {
    BaiRef &R = bai.refs[r];
    need(4);
    const uint32_t n_bin = rd32(&d[p]);
    p += 4;
    for (uint32_t i = 0; i < n_bin; i++) {
      need(8);
      const uint32_t bin = rd32(&d[p]), n_chunk = rd32(&d[p + 4]);
      p += 8;
      need((size_t)n_chunk * 16);
      if (bin == 37450) {  // metadata pseudo-bin: chunk 1 holds (mapped, unmapped)
        if (n_chunk == 2) {
          R.has_meta = true;
          R.mapped = rd64(&d[p + 16]);
          R.unmapped = rd64(&d[p + 24]);
        }
      } else {
        for (uint32_t c = 0; c < n_chunk; c++) {
          BaiChunk ch{rd64(&d[p + 16 * (size_t)c]), rd64(&d[p + 16 * (size_t)c + 8])};
          R.chunks.push_back(ch);
          R.min_voff = std::min(R.min_voff, ch.beg);
          R.max_voff = std::max(R.max_voff, ch.end);
        }
      }
      p += (size_t)n_chunk * 16;
    }
    need(4);
    const uint32_t n_intv = rd32(&d[p]);
    p += 4;
    need((size_t)n_intv * 8);
    R.linear.resize(n_intv);
    for (uint32_t i = 0; i < n_intv; i++) R.linear[i] = rd64(&d[p + 8 * (size_t)i]);
    p += (size_t)n_intv * 8;
  }
  if (p + 8 <= d.size()) {
    bai.has_no_coor = true;
    bai.n_no_coor = rd64(&d[p]);
  }
  return bai;
}

void scan_bgzf_blocks(const uint8_t *buf, uint64_t n, BlockTable &out, uint64_t max_blocks) {
  out.coff.clear();
  out.uoff.clear();
  uint64_t c = 0, u = 0;
  out.coff.push_back(0);
  out.uoff.push_back(0);
  while (c + 18 <= n && out.coff.size() - 1 < max_blocks) {
    const uint8_t *h = buf + c;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) fail("invalid BGZF block header");
    const uint32_t xlen = rd16(h + 10);
    uint32_t bsize = 0;
    bool found = false;
    if (xlen == 6 && h[12] == 'B' && h[13] == 'C' && rd16(h + 14) == 2) {
      bsize = rd16(h + 16);
      found = true;
    } else {
      if (c + 12 + xlen > n) break;
      for (uint32_t i = 0; i + 4 <= xlen;) {
        const uint8_t *x = h + 12 + i;
        const uint32_t slen = rd16(x + 2);
        if (x[0] == 'B' && x[1] == 'C' && slen == 2 && i + 6 <= xlen) {
          bsize = rd16(x + 4);
          found = true;
          break;
        }
        i += 4 + slen;
      }
    }
    if (!found) fail("BGZF block without a BC subfield");
    const uint64_t total = (uint64_t)bsize + 1;
    if (total < 12ull + xlen + 8) fail("invalid BGZF block size");
    if (c + total > n) break;
    const uint32_t isize = rd32(h + total - 4);
    if (isize > 65536) fail("BGZF block ISIZE exceeds 64 KiB");
    c += total;
    u += isize;
    out.coff.push_back(c);
    out.uoff.push_back(u);
  }
}

bool parse_bam_header(const uint8_t *d, uint64_t n, BamHeader &out, uint64_t &consumed) {
  uint64_t p = 0;
  auto have = [&](uint64_t k) { return p + k <= n; };
  if (!have(8)) return false;
  if (memcmp(d, "BAM\1", 4) != 0) fail("not a BAM file (bad magic)");
  const uint32_t l_text = rd32(d + 4);
  p = 8;
  if (!have((uint64_t)l_text + 4)) return false;
  p += l_text;
  const uint32_t n_ref = rd32(d + p);
  p += 4;
  BamHeader h;
  h.text.assign((const char *)d + 8, l_text);
  h.names.reserve(n_ref);
  h.lens.reserve(n_ref);
  for (uint32_t i = 0; i < n_ref; i++) {
    if (!have(4)) return false;
    const uint32_t l_name = rd32(d + p);
    p += 4;
    if (!have((uint64_t)l_name + 4)) return false;
    std::string name((const char *)d + p, l_name ? l_name - 1 : 0);
    p += l_name;
    h.lens.push_back(rd32(d + p));
    p += 4;
    h.name_to_id[name] = (int)h.names.size();
    h.names.push_back(std::move(name));
  }
  consumed = p;
  out = std::move(h);
  return true;
}

uint64_t Geometry::first_bin_of_chunk(uint64_t c) const {
  if (c >= total_chunks()) return total_bins();
  int r = ref_of_chunk(c);
  return ref_first_bin[r] + (c - ref_first_chunk[r]) * nbf;
}
int Geometry::ref_of_chunk(uint64_t c) const {
  // last r with ref_first_chunk[r] <= c and chunks_of(r) > 0
  size_t lo = 0, hi = ref_first_chunk.size() - 1;  // r in [0, n_ref)
  while (hi - lo > 1) {
    size_t mid = (lo + hi) / 2;
    if (ref_first_chunk[mid] <= c) lo = mid;
    else hi = mid;
  }
  return (int)lo;
}

Geometry make_geometry(const BamHeader &h, const Bai &bai, uint64_t bin, bool drop_scaffold_refs) {
  if (bin == 0) fail("bin size must be greater than 0", ERR_INVALID);
  Geometry g;
  const size_t n_ref = h.names.size();
  g.bin = bin;
  g.in_stats.assign(n_ref, 0);
  // chrom_stats is keyed by name over zip(header refs, index refs): a duplicated name keeps its last entry
  const size_t n = std::min(n_ref, bai.refs.size());
  std::unordered_map<std::string, size_t> by_name;
  for (size_t i = 0; i < n; i++) by_name[h.names[i]] = i;
  uint64_t mapped = 0, unmapped = bai.has_no_coor ? bai.n_no_coor : 0;
  for (auto &kv : by_name) {
    const size_t i = kv.second;
    g.in_stats[i] = 1;
    g.genome_len += h.lens[i];
    mapped += bai.refs[i].mapped;
    unmapped += bai.refs[i].unmapped;
  }
  g.n_mapped = mapped;
  g.n_unmapped = unmapped;
  // estimate_genome_chunk_length (bam_utils.rs:384-402), f64 arithmetic then truncation
  if (g.genome_len == 0 || g.n_mapped == 0) {
    g.L = bin;
  } else {
    double per_bp = (double)g.n_mapped / (double)g.genome_len;
    double L = std::min(5e6, 2e6 / per_bp);
    L -= std::fmod(L, (double)bin);
    g.L = (uint64_t)L;
  }
  if (g.L == 0) fail("genome chunk length is 0 (bin size larger than the estimated chunk length)");
  g.nbf = (g.L - 1 + bin - 1) / bin;
  g.ref_first_bin.assign(n_ref + 1, 0);
  g.ref_first_chunk.assign(n_ref + 1, 0);
  for (size_t r = 0; r < n_ref; r++) {
    uint64_t bins = 0, chunks = 0;
    if (g.in_stats[r] && h.lens[r] > 0 && !(drop_scaffold_refs && is_scaffold(h.names[r]))) {
      const uint64_t len = h.lens[r];
      chunks = (len + g.L - 1) / g.L;
      const uint64_t cs_last = (chunks - 1) * g.L + 1;  // last chunk [cs_last, len]: bins tile [cs_last, len)
      bins = (chunks - 1) * g.nbf + (len - cs_last + bin - 1) / bin;
    }
    g.ref_first_bin[r + 1] = g.ref_first_bin[r] + bins;
    g.ref_first_chunk[r + 1] = g.ref_first_chunk[r] + chunks;
  }
  return g;
}

bool is_scaffold(const std::string &n) {
  auto ends = [&](const char *s) {
    size_t l = strlen(s);
    return n.size() >= l && n.compare(n.size() - l, l, s) == 0;
  };
  return n.compare(0, 5, "chrUn") == 0 || ends("_alt") || ends("_random");
}

uint64_t voff_lower_bound(const Bai &bai, int ref, uint64_t pos1) {
  const int n_ref = (int)bai.refs.size();
  if (ref < 0) ref = 0, pos1 = 1;
  if (ref < n_ref) {
    const BaiRef &R = bai.refs[ref];
    const uint64_t w = (pos1 > 0 ? pos1 - 1 : 0) >> 14;
    if (w < R.linear.size() && R.min_voff != ~0ull) {
      for (uint64_t k = w + 1; k-- > 0;)
        if (R.linear[k] != 0) return std::max(R.linear[k], R.min_voff);
      return R.min_voff;
    }
    // beyond the last indexed window: nothing of this reference can overlap pos1 or later
  }
  for (int r = ref + 1; r < n_ref; r++)
    if (bai.refs[r].min_voff != ~0ull) return bai.refs[r].min_voff;
  return ~0ull;
}

std::string format_f64(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  std::string out;
  if (std::signbit(v)) {
    out.push_back('-');
    v = -v;
  }
  if (v == 0.0) return out + "0.0";
  char buf[64];
  auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);  // shortest round-trip digits
  std::string sci(buf, res.ptr);
  const size_t epos = sci.find('e');
  std::string digits;
  for (size_t i = 0; i < epos; i++)
    if (sci[i] != '.') digits.push_back(sci[i]);
  const int e10 = atoi(sci.c_str() + epos + 1);
  const int len = (int)digits.size(), kk = e10 + 1;  // value = 0.digits * 10^kk
  if (len <= kk && kk <= 16) {                        // integer: digits, zeros, ".0"
    out += digits;
    out.append((size_t)(kk - len), '0');
    out += ".0";
  } else if (0 < kk && kk <= 16) {                    // dd.ddd
    out += digits.substr(0, (size_t)kk);
    out.push_back('.');
    out += digits.substr((size_t)kk);
  } else if (-5 < kk && kk <= 0) {                    // 0.000ddd
    out += "0.";
    out.append((size_t)(-kk), '0');
    out += digits;
  } else {                                            // d[.ddd]e±x
    out.push_back(digits[0]);
    if (len > 1) {
      out.push_back('.');
      out += digits.substr(1);
    }
    out.push_back('e');
    out += std::to_string(kk - 1);
  }
  return out;
}

}  // namespace bnd
