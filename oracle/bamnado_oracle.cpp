// bamnado_oracle.cpp — CPU ORACLE for the BamNado coverage hot path.
//
// TEST INFRASTRUCTURE ONLY (see bamnado_oracle.h).  A from-scratch C++17 + zlib restatement of what the
// reference computes, following its work decomposition (one task per genomic chunk, BAI query, per-read
// filter, interval sort + two binary searches per bin) so that it can double as the timed CPU baseline.
// Citations are into /root/reference/bamnado/src/.  Third-party crate behaviour that is not on disk
// (noodles 0.103 BGZF/BAM/BAI/BED, rust-lapper 1.1 count/merge_overlaps, polars 0.51 CSV float text via ryu)
// is restated from the published algorithms; SURVEY.md §9.11 lists the assumptions.
#include "bamnado_oracle.h"

#include <zlib.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <sstream>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

thread_local std::string g_err;

struct OrcError : std::runtime_error {
  int code;
  OrcError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] void fail(const std::string &m, int code = 1) { throw OrcError(code, m); }

inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline int32_t rdi32(const uint8_t *p) { return (int32_t)rd32(p); }
inline uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

// ---------------------------------------------------------------------------------------------------------
// BGZF reader (noodles-bgzf semantics: virtual position = (coffset << 16) | uoffset, and an exhausted block
// reports the next block's coffset with uoffset 0).
// ---------------------------------------------------------------------------------------------------------
class Bgzf {
 public:
  explicit Bgzf(const std::string &path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) fail("File not found: " + path);
    struct stat st;
    fstat(fd_, &st);
    fsize_ = (uint64_t)st.st_size;
    memset(&zs_, 0, sizeof zs_);
    if (inflateInit2(&zs_, -15) != Z_OK) fail("zlib init failed");
    ubuf_.resize(65536);
    cbuf_.resize(65536 + 32);
  }
  ~Bgzf() {
    inflateEnd(&zs_);
    if (fd_ >= 0) ::close(fd_);
  }
  Bgzf(const Bgzf &) = delete;

  void seek(uint64_t voff) {
    uint64_t co = voff >> 16;
    uint32_t uo = (uint32_t)(voff & 0xffff);
    if (!(have_block_ && co == block_coff_)) load_block(co);
    upos_ = uo;
  }
  uint64_t tell() const {
    if (!have_block_) return next_coff_ << 16;
    if (upos_ < ulen_) return (block_coff_ << 16) | upos_;
    return next_coff_ << 16;
  }
  // read exactly n bytes unless EOF; returns bytes read
  size_t read(uint8_t *dst, size_t n) {
    size_t got = 0;
    while (got < n) {
      if (!have_block_ || upos_ >= ulen_) {
        if (!load_block(next_coff_)) break;
        continue;
      }
      size_t take = std::min<size_t>(n - got, ulen_ - upos_);
      memcpy(dst + got, ubuf_.data() + upos_, take);
      upos_ += (uint32_t)take;
      got += take;
    }
    return got;
  }
  uint64_t n_blocks_loaded = 0, bytes_inflated = 0;

 private:
  bool load_block(uint64_t coff) {
    have_block_ = false;
    if (coff + 18 > fsize_) {
      next_coff_ = coff;
      return false;
    }
    uint8_t hdr[18];
    if (pread(fd_, hdr, 18, (off_t)coff) != 18) fail("BGZF: short read of block header");
    if (hdr[0] != 0x1f || hdr[1] != 0x8b || hdr[2] != 8 || !(hdr[3] & 4)) fail("BGZF: bad block magic");
    uint32_t xlen = rd16(hdr + 10);
    // locate the BC subfield (normally the only one)
    uint32_t bsize = 0;
    if (xlen == 6 && hdr[12] == 'B' && hdr[13] == 'C') {
      bsize = rd16(hdr + 16);
    } else {
      std::vector<uint8_t> x(xlen);
      if (pread(fd_, x.data(), xlen, (off_t)coff + 12) != (ssize_t)xlen) fail("BGZF: short extra field");
      bool found = false;
      for (uint32_t i = 0; i + 4 <= xlen;) {
        uint32_t slen = rd16(&x[i + 2]);
        if (x[i] == 'B' && x[i + 1] == 'C' && slen == 2) {
          bsize = rd16(&x[i + 4]);
          found = true;
          break;
        }
        i += 4 + slen;
      }
      if (!found) fail("BGZF: no BC subfield");
    }
    uint32_t total = bsize + 1;
    uint32_t hlen = 12 + xlen;
    if (total < hlen + 8) fail("BGZF: block too small");
    if (pread(fd_, cbuf_.data(), total, (off_t)coff) != (ssize_t)total) fail("BGZF: short block");
    uint32_t clen = total - hlen - 8;
    uint32_t crc = rd32(&cbuf_[total - 8]);
    uint32_t isize = rd32(&cbuf_[total - 4]);
    if (isize > 65536) fail("BGZF: isize too large");
    inflateReset(&zs_);
    zs_.next_in = cbuf_.data() + hlen;
    zs_.avail_in = clen;
    zs_.next_out = ubuf_.data();
    zs_.avail_out = 65536;
    int rc = inflate(&zs_, Z_FINISH);
    if (rc != Z_STREAM_END || zs_.total_out != isize) fail("BGZF: inflate failed");
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), ubuf_.data(), isize) != crc) fail("BGZF: CRC mismatch");
    block_coff_ = coff;
    next_coff_ = coff + total;
    ulen_ = isize;
    upos_ = 0;
    have_block_ = true;
    n_blocks_loaded++;
    bytes_inflated += isize;
    return true;
  }
  int fd_ = -1;
  uint64_t fsize_ = 0;
  z_stream zs_;
  std::vector<uint8_t> ubuf_, cbuf_;
  bool have_block_ = false;
  uint64_t block_coff_ = 0, next_coff_ = 0;
  uint32_t ulen_ = 0, upos_ = 0;
};

// ---------------------------------------------------------------------------------------------------------
// BAM header and BAI index
// ---------------------------------------------------------------------------------------------------------
struct Header {
  std::vector<std::string> names;
  std::vector<uint64_t> lens;
  std::unordered_map<std::string, int> name_to_id;
  uint64_t first_record_voff = 0;
};

Header read_header(Bgzf &bz) {
  Header h;
  uint8_t b[8];
  if (bz.read(b, 4) != 4 || memcmp(b, "BAM\1", 4) != 0) fail("not a BAM file");
  if (bz.read(b, 4) != 4) fail("truncated header");
  uint32_t l_text = rd32(b);
  std::vector<uint8_t> text(l_text);
  if (bz.read(text.data(), l_text) != l_text) fail("truncated header text");
  if (bz.read(b, 4) != 4) fail("truncated header");
  uint32_t n_ref = rd32(b);
  for (uint32_t i = 0; i < n_ref; i++) {
    if (bz.read(b, 4) != 4) fail("truncated header refs");
    uint32_t l_name = rd32(b);
    std::vector<uint8_t> nm(l_name);
    if (bz.read(nm.data(), l_name) != l_name) fail("truncated header refs");
    if (bz.read(b, 4) != 4) fail("truncated header refs");
    std::string name((const char *)nm.data(), l_name ? l_name - 1 : 0);
    h.name_to_id[name] = (int)h.names.size();
    h.names.push_back(name);
    h.lens.push_back(rd32(b));
  }
  h.first_record_voff = bz.tell();
  return h;
}

struct BaiChunk {
  uint64_t beg, end;
};
struct BaiRef {
  std::vector<std::pair<uint32_t, std::vector<BaiChunk>>> bins;  // file order
  std::vector<uint64_t> linear;
  bool has_meta = false;
  uint64_t mapped = 0, unmapped = 0;
};
struct Bai {
  std::vector<BaiRef> refs;
  bool has_no_coor = false;
  uint64_t n_no_coor = 0;
};

Bai read_bai(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) fail("BAM index file does not exist: " + path);
  std::vector<uint8_t> d;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  d.resize((size_t)sz);
  if (sz && fread(d.data(), 1, (size_t)sz, f) != (size_t)sz) {
    fclose(f);
    fail("short read of BAI");
  }
  fclose(f);
  size_t p = 0;
  auto need = [&](size_t n) {
    if (p + n > d.size()) fail("truncated BAI");
  };
  need(8);
  if (memcmp(d.data(), "BAI\1", 4) != 0) fail("bad BAI magic");
  uint32_t n_ref = rd32(&d[4]);
  p = 8;
  Bai bai;
  bai.refs.resize(n_ref);
  for (uint32_t r = 0; r < n_ref; r++) {
    need(4);
    uint32_t n_bin = rd32(&d[p]);
    p += 4;
    BaiRef &R = bai.refs[r];
    for (uint32_t i = 0; i < n_bin; i++) {
      need(8);
      uint32_t bin = rd32(&d[p]);
      uint32_t n_chunk = rd32(&d[p + 4]);
      p += 8;
      need((size_t)n_chunk * 16);
      std::vector<BaiChunk> ch(n_chunk);
      for (uint32_t c = 0; c < n_chunk; c++) {
        ch[c].beg = rd64(&d[p]);
        ch[c].end = rd64(&d[p + 8]);
        p += 16;
      }
      if (bin == 37450) {  // metadata pseudo-bin
        if (n_chunk == 2) {
          R.has_meta = true;
          R.mapped = ch[1].beg;
          R.unmapped = ch[1].end;
        }
      } else {
        R.bins.emplace_back(bin, std::move(ch));
      }
    }
    need(4);
    uint32_t n_intv = rd32(&d[p]);
    p += 4;
    need((size_t)n_intv * 8);
    R.linear.resize(n_intv);
    for (uint32_t i = 0; i < n_intv; i++) {
      R.linear[i] = rd64(&d[p]);
      p += 8;
    }
  }
  if (p + 8 <= d.size()) {
    bai.has_no_coor = true;
    bai.n_no_coor = rd64(&d[p]);
  }
  return bai;
}

// UCSC binning scheme, min_shift 14, depth 5: bins overlapping 0-based half-open [beg,end)
void reg2bins(uint64_t beg, uint64_t end, std::vector<uint32_t> &out) {
  out.clear();
  if (beg >= end) return;
  if (end > (1ull << 29)) end = 1ull << 29;
  --end;
  out.push_back(0);
  for (uint32_t k = 1 + (uint32_t)(beg >> 26); k <= 1 + (uint32_t)(end >> 26); ++k) out.push_back(k);
  for (uint32_t k = 9 + (uint32_t)(beg >> 23); k <= 9 + (uint32_t)(end >> 23); ++k) out.push_back(k);
  for (uint32_t k = 73 + (uint32_t)(beg >> 20); k <= 73 + (uint32_t)(end >> 20); ++k) out.push_back(k);
  for (uint32_t k = 585 + (uint32_t)(beg >> 17); k <= 585 + (uint32_t)(end >> 17); ++k) out.push_back(k);
  for (uint32_t k = 4681 + (uint32_t)(beg >> 14); k <= 4681 + (uint32_t)(end >> 14); ++k) out.push_back(k);
}

// noodles csi binning_index query: chunks of all overlapping bins, dropped when end <= linear min offset,
// sorted by start and merged when they touch/overlap.
std::vector<BaiChunk> bai_query(const BaiRef &R, uint64_t start1, uint64_t end1) {
  std::vector<uint32_t> bins;
  reg2bins(start1 - 1, end1, bins);
  std::vector<char> want(37451, 0);
  for (uint32_t b : bins)
    if (b < want.size()) want[b] = 1;
  std::vector<BaiChunk> chunks;
  for (auto &kv : R.bins)
    if (kv.first < want.size() && want[kv.first])
      for (auto &c : kv.second) chunks.push_back(c);
  uint64_t min_off = 0;
  size_t li = (size_t)((start1 - 1) >> 14);
  if (li < R.linear.size()) min_off = R.linear[li];
  std::vector<BaiChunk> kept;
  for (auto &c : chunks)
    if (c.end > min_off) kept.push_back(c);
  if (kept.empty()) return kept;
  std::sort(kept.begin(), kept.end(), [](const BaiChunk &a, const BaiChunk &b) { return a.beg < b.beg; });
  std::vector<BaiChunk> merged;
  merged.push_back(kept[0]);
  for (size_t i = 1; i < kept.size(); i++) {
    BaiChunk &a = merged.back();
    if (kept[i].beg > a.end) {
      merged.push_back(kept[i]);
      continue;
    }
    if (a.end < kept[i].end) a.end = kept[i].end;
  }
  return merged;
}

// ---------------------------------------------------------------------------------------------------------
// BamStats (bam_utils.rs:311-436)
// ---------------------------------------------------------------------------------------------------------
struct BamStats {
  Header header;
  Bai bai;
  uint64_t n_mapped = 0, n_unmapped = 0, genome_len = 0;
  std::vector<char> in_stats;  // refs that got a chrom_stats entry (zip of header refs and index refs)
};

std::string bai_path_for(const std::string &bam) { return bam + ".bai"; }

BamStats bam_stats_new(const std::string &bam) {
  BamStats s;
  {
    Bgzf bz(bam);
    s.header = read_header(bz);
  }
  s.bai = read_bai(bai_path_for(bam));
  size_t n = std::min(s.header.names.size(), s.bai.refs.size());  // zip()
  s.in_stats.assign(s.header.names.size(), 0);
  // chrom_stats is a HashMap keyed by name: duplicate names overwrite (bam_utils.rs:355) — mirror with a map
  std::unordered_map<std::string, size_t> by_name;
  for (size_t i = 0; i < n; i++) by_name[s.header.names[i]] = i;
  uint64_t unmapped = s.bai.has_no_coor ? s.bai.n_no_coor : 0, mapped = 0;
  for (auto &kv : by_name) {
    size_t i = kv.second;
    s.in_stats[i] = 1;
    s.genome_len += s.header.lens[i];
    mapped += s.bai.refs[i].mapped;  // metadata().unwrap_or_default()
    unmapped += s.bai.refs[i].unmapped;
  }
  s.n_mapped = mapped;
  s.n_unmapped = unmapped;
  return s;
}

// estimate_genome_chunk_length (bam_utils.rs:384-402)
uint64_t chunk_length(const BamStats &s, uint64_t bin) {
  if (s.genome_len == 0 || s.n_mapped == 0) return bin;
  double max_reads_per_bp = (double)s.n_mapped / (double)s.genome_len;
  double L = std::min(5e6, 2e6 / max_reads_per_bp);
  double corr = std::fmod(L, (double)bin);
  L -= corr;
  return (uint64_t)L;
}

struct Region {
  int ref_id;
  uint64_t start, end;  // 1-based inclusive
};

// genome_chunks / chromosome_chunks (bam_utils.rs:409-466)
void chunks_for_ref(const BamStats &s, int ref_id, uint64_t L, std::vector<Region> &out) {
  if (L == 0) fail("genome chunk length is 0 (bin size larger than the chunk length)");
  uint64_t cs = 1, ce_max = s.header.lens[ref_id];
  while (cs <= ce_max) {
    uint64_t ce = std::min(cs + L - 1, ce_max);
    out.push_back(Region{ref_id, cs, ce});
    cs = ce + 1;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Filter (read_filter.rs)
// ---------------------------------------------------------------------------------------------------------
struct Lapper {  // rust-lapper 1.1: only what the path uses (new / count / merge_overlaps)
  struct Iv {
    uint64_t start, stop;
  };
  std::vector<Iv> ivs;
  std::vector<uint64_t> starts, stops;
  void build(std::vector<Iv> v) {
    ivs = std::move(v);
    std::sort(ivs.begin(), ivs.end(), [](const Iv &a, const Iv &b) { return a.start != b.start ? a.start < b.start : a.stop < b.stop; });
    refresh();
  }
  void refresh() {
    starts.resize(ivs.size());
    stops.resize(ivs.size());
    for (size_t i = 0; i < ivs.size(); i++) {
      starts[i] = ivs[i].start;
      stops[i] = ivs[i].stop;
    }
    std::sort(starts.begin(), starts.end());
    std::sort(stops.begin(), stops.end());
  }
  // count(start, stop) = #{starts < stop} - #{stops < start+1}, in wrapping usize arithmetic (BITS)
  uint64_t count(uint64_t start, uint64_t stop) const {
    uint64_t len = ivs.size();
    uint64_t first = (uint64_t)(std::lower_bound(stops.begin(), stops.end(), start + 1) - stops.begin());
    uint64_t last = (uint64_t)(std::lower_bound(starts.begin(), starts.end(), stop) - starts.begin());
    uint64_t num_cant_after = len - last;
    return len - first - num_cant_after;
  }
  void merge_overlaps() {
    std::vector<Iv> st;
    for (auto &iv : ivs) {
      if (st.empty()) {
        st.push_back(iv);
        continue;
      }
      Iv &top = st.back();
      if (top.stop < iv.start)
        st.push_back(iv);
      else if (top.stop < iv.stop)
        top.stop = iv.stop;
    }
    ivs = std::move(st);
    refresh();
  }
};

struct Filter {
  int strand = 0;
  bool proper_pair = false;
  uint32_t min_mapq = 0, min_length = 0, max_length = 0xffffffffu;
  bool has_min_frag = false, has_max_frag = false;
  uint32_t min_frag = 0, max_frag = 0;
  bool ex_dup = false, ex_sec = false, ex_sup = false;
  bool has_rg = false, has_tag = false, has_tag_value = false;
  std::string rg, tag, tag_value;
  bool has_blacklist = false;
  std::unordered_map<int, Lapper> blacklist;
  bool has_barcodes = false;
  std::unordered_set<std::string> barcodes;
  mutable std::atomic<uint64_t> stats[ORC_N_COUNTERS];
  Filter() {
    for (auto &s : stats) s.store(0);
  }
  void bump(int i) const { stats[i].fetch_add(1, std::memory_order_relaxed); }
};

std::vector<std::string> split_ws(const std::string &l) {
  std::vector<std::string> out;
  size_t i = 0;
  while (i < l.size()) {
    while (i < l.size() && isspace((unsigned char)l[i])) i++;
    size_t j = i;
    while (j < l.size() && !isspace((unsigned char)l[j])) j++;
    if (j > i) out.emplace_back(l.substr(i, j - i));
    i = j;
  }
  return out;
}

bool read_lines(const std::string &path, std::vector<std::string> &lines) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  std::string cur;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) {
    for (size_t i = 0; i < n; i++) {
      if (buf[i] == '\n') {
        if (!cur.empty() && cur.back() == '\r') cur.pop_back();
        lines.push_back(cur);
        cur.clear();
      } else
        cur.push_back(buf[i]);
    }
  }
  if (!cur.empty()) lines.push_back(cur);
  fclose(f);
  return true;
}

// bed_to_lapper + convert_lapper_chrom_names_to_ids (bam_utils.rs:650-703).  noodles BED: feature_start =
// BED start + 1 (1-based Position), feature_end = BED end; stored as Lapper [start, stop).
void load_blacklist(Filter &F, const std::string &bed, const Header &h, bool merge) {
  std::vector<std::string> lines;
  if (!read_lines(bed, lines)) fail("Failed to read blacklisted locations: " + bed);
  std::map<std::string, std::vector<Lapper::Iv>> by_chrom;
  for (auto &l : lines) {
    if (l.empty() || l[0] == '#') continue;
    auto parts = split_ws(l);
    if (parts.empty() || parts[0] == "track" || parts[0] == "browser") continue;
    if (parts.size() < 4) fail("Failed to read blacklisted locations: BED record needs >= 4 fields");
    char *e1 = nullptr, *e2 = nullptr;
    unsigned long long s = strtoull(parts[1].c_str(), &e1, 10), e = strtoull(parts[2].c_str(), &e2, 10);
    if (*e1 || *e2) fail("Failed to read blacklisted locations: invalid BED coordinates");
    if (e == 0) fail("Failed to get feature end");
    by_chrom[parts[0]].push_back(Lapper::Iv{(uint64_t)s + 1, (uint64_t)e});
  }
  if (by_chrom.empty()) fail("Lapper is empty");
  for (auto &kv : by_chrom) {
    auto it = h.name_to_id.find(kv.first);
    if (it == h.name_to_id.end()) fail("Chromosome " + kv.first + " not found in BAM file");
    Lapper lp;
    lp.build(kv.second);
    if (merge) lp.merge_overlaps();
    F.blacklist[it->second] = std::move(lp);
  }
  F.has_blacklist = true;
}

std::vector<std::string> split_csv_line(const std::string &l) {
  std::vector<std::string> out;
  std::string cur;
  bool q = false;
  for (size_t i = 0; i < l.size(); i++) {
    char c = l[i];
    if (q) {
      if (c == '"') {
        if (i + 1 < l.size() && l[i + 1] == '"') {
          cur.push_back('"');
          i++;
        } else
          q = false;
      } else
        cur.push_back(c);
    } else if (c == '"')
      q = true;
    else if (c == ',') {
      out.push_back(cur);
      cur.clear();
    } else
      cur.push_back(c);
  }
  out.push_back(cur);
  return out;
}

// CellBarcodes::from_csv / CellBarcodesMulti::from_csv (bam_utils.rs:186-269)
void load_barcodes_csv(Filter &F, const std::string &csv, int column) {
  std::vector<std::string> lines;
  if (!read_lines(csv, lines) || lines.empty()) fail("Failed to read barcodes: " + csv);
  auto hdr = split_csv_line(lines[0]);
  int col = column;
  if (col < 0) {
    for (size_t i = 0; i < hdr.size(); i++)
      if (hdr[i] == "barcode") col = (int)i;
    if (col < 0) fail("Missing 'barcode' column in barcode CSV");
  } else if ((size_t)col >= hdr.size())
    fail("barcode CSV has too few columns");
  for (size_t i = 1; i < lines.size(); i++) {
    if (lines[i].empty()) continue;
    auto parts = split_csv_line(lines[i]);
    if ((size_t)col >= parts.size() || parts[col].empty()) {
      if (column < 0) fail("Null value in 'barcode' column");
      continue;  // CellBarcodesMulti drops nulls
    }
    F.barcodes.insert(parts[col]);
  }
  F.has_barcodes = true;
}

std::unique_ptr<Filter> make_filter(const orc_filter_cfg *c, const Header &h) {
  auto F = std::make_unique<Filter>();
  F->strand = c->strand;
  F->proper_pair = c->proper_pair != 0;
  F->min_mapq = c->min_mapq;
  F->min_length = c->min_length;
  F->max_length = c->max_length;
  F->has_min_frag = c->has_min_fragment_length != 0;
  F->min_frag = c->min_fragment_length;
  F->has_max_frag = c->has_max_fragment_length != 0;
  F->max_frag = c->max_fragment_length;
  F->ex_dup = c->exclude_duplicates != 0;
  F->ex_sec = c->exclude_secondary != 0;
  F->ex_sup = c->exclude_supplementary != 0;
  if (c->read_group) {
    F->has_rg = true;
    F->rg = c->read_group;
  }
  if (c->filter_tag) {
    F->has_tag = true;
    F->tag = c->filter_tag;
  }
  if (c->filter_tag_value) {
    F->has_tag_value = true;
    F->tag_value = c->filter_tag_value;
  }
  if (c->blacklist_bed) load_blacklist(*F, c->blacklist_bed, h, c->blacklist_merge_overlaps != 0);
  if (c->barcode_csv) load_barcodes_csv(*F, c->barcode_csv, c->barcode_csv_column);
  if (c->has_barcode_list) {
    F->has_barcodes = true;
    for (uint64_t i = 0; i < c->n_barcodes; i++) F->barcodes.insert(c->barcodes[i]);
  }
  return F;
}

// a decoded view of one BAM record (payload after block_size)
struct Rec {
  const uint8_t *p;
  uint32_t len;
  int32_t ref_id() const { return rdi32(p); }
  int32_t pos() const { return rdi32(p + 4); }
  uint32_t l_read_name() const { return p[8]; }
  uint32_t mapq() const { return p[9]; }
  uint32_t n_cigar() const { return rd16(p + 12); }
  uint32_t flag() const { return rd16(p + 14); }
  uint32_t l_seq() const { return rd32(p + 16); }
  int32_t next_pos() const { return rdi32(p + 24); }
  int32_t tlen() const { return rdi32(p + 28); }
  const uint8_t *cigar() const { return p + 32 + l_read_name(); }
  const uint8_t *aux() const { return cigar() + 4 * (size_t)n_cigar() + (l_seq() + 1) / 2 + l_seq(); }
  const uint8_t *end() const { return p + len; }
};

// alignment_span (genomic_intervals.rs:27-44): M, D, N, =, X consume the reference
uint64_t cigar_span(const Rec &r) {
  uint64_t span = 0;
  const uint8_t *c = r.cigar();
  for (uint32_t i = 0; i < r.n_cigar(); i++) {
    uint32_t v = rd32(c + 4 * i);
    uint32_t op = v & 0xf, n = v >> 4;
    if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += n;
  }
  return span;
}

// noodles aux `data.get(tag)`: first field with that tag.  Returns 0 not found, 1 found with Z string
// (value set), 2 found but not a Z string, -1 malformed before/at the field (an Err in the reference).
int aux_get_string(const Rec &r, char t0, char t1, std::string &value) {
  const uint8_t *a = r.aux(), *e = r.end();
  if (a > e) return -1;
  while (a < e) {
    if (e - a < 3) return -1;
    char a0 = (char)a[0], a1 = (char)a[1];
    uint8_t ty = a[2];
    a += 3;
    size_t sz = 0;
    bool is_z = false;
    switch (ty) {
      case 'A': case 'c': case 'C': sz = 1; break;
      case 's': case 'S': sz = 2; break;
      case 'i': case 'I': case 'f': sz = 4; break;
      case 'Z': case 'H': {
        const uint8_t *q = (const uint8_t *)memchr(a, 0, (size_t)(e - a));
        if (!q) return -1;
        sz = (size_t)(q - a) + 1;
        is_z = (ty == 'Z');
        break;
      }
      case 'B': {
        if (e - a < 5) return -1;
        uint8_t st = a[0];
        uint32_t n = rd32(a + 1);
        size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : (st == 'i' || st == 'I' || st == 'f') ? 4 : 0;
        if (!es) return -1;
        sz = 5 + (size_t)n * es;
        break;
      }
      default: return -1;
    }
    if ((size_t)(e - a) < sz) return -1;
    if (a0 == t0 && a1 == t1) {
      if (is_z) {
        value.assign((const char *)a, sz - 1);
        return 1;
      }
      return 2;
    }
    a += sz;
  }
  return 0;
}

// BamReadFilter::is_valid (read_filter.rs:570-803): 1 valid, 0 filtered (one counter bumped), -1 Err (dropped,
// no failure counter).
int is_valid(const Filter &F, const Rec &r) {
  F.bump(ORC_N_TOTAL);
  uint32_t flag = r.flag();
  if (F.ex_dup && (flag & 0x400)) return F.bump(ORC_FAIL_DUPLICATE), 0;
  if (F.ex_sec && (flag & 0x100)) return F.bump(ORC_FAIL_SECONDARY), 0;
  if (F.ex_sup && (flag & 0x800)) return F.bump(ORC_FAIL_SUPPLEMENTARY), 0;
  bool rev = flag & 0x10;
  if (rev && F.strand == 1) return F.bump(ORC_FAIL_STRAND), 0;
  if (!rev && F.strand == 2) return F.bump(ORC_FAIL_STRAND), 0;
  if (F.proper_pair && !(flag & 0x2)) return F.bump(ORC_FAIL_PROPER_PAIR), 0;
  if (flag & 0x4) return F.bump(ORC_FAIL_MAPQ), 0;
  uint32_t mq = r.mapq();
  if (mq != 255 && mq < F.min_mapq) return F.bump(ORC_FAIL_MAPQ), 0;  // 255 == missing == passes
  uint64_t l_seq = r.l_seq();
  if (l_seq < F.min_length || l_seq > F.max_length) return F.bump(ORC_FAIL_LENGTH), 0;
  if (F.has_min_frag || F.has_max_frag) {
    int32_t t = r.tlen();
    uint32_t at = t < 0 ? (uint32_t)(-(int64_t)t) : (uint32_t)t;
    if (F.has_min_frag && at < F.min_frag) return F.bump(ORC_FAIL_FRAGMENT_LENGTH), 0;
    if (F.has_max_frag && at > F.max_frag) return F.bump(ORC_FAIL_FRAGMENT_LENGTH), 0;
  }
  int32_t ref_id = r.ref_id();
  if (ref_id < 0) return -1;  // "Failed to get reference sequence ID"
  int32_t pos0 = r.pos();
  if (pos0 < 0) return F.bump(ORC_FAIL_MAPQ), 0;  // missing alignment start
  uint64_t start = (uint64_t)pos0 + 1, end = start + l_seq;
  if (F.has_blacklist) {
    auto it = F.blacklist.find(ref_id);
    if (it != F.blacklist.end() && it->second.count(start, end) > 0) return F.bump(ORC_FAIL_BLACKLIST), 0;
  }
  std::string v;
  if (F.has_barcodes) {
    int rc = aux_get_string(r, 'C', 'B', v);
    if (rc != 1 || !F.barcodes.count(v)) return F.bump(ORC_FAIL_BARCODE), 0;
  }
  if (F.has_rg) {
    int rc = aux_get_string(r, 'R', 'G', v);
    if (rc == 0) return F.bump(ORC_FAIL_READ_GROUP), 0;
    if (rc < 0) return -1;  // `Some(v) => v?`
    if (rc == 2 || v != F.rg) return F.bump(ORC_FAIL_READ_GROUP), 0;
  }
  if (F.has_tag && F.has_tag_value) {
    if (F.tag.size() != 2) return F.bump(ORC_FAIL_TAG), 0;
    int rc = aux_get_string(r, F.tag[0], F.tag[1], v);
    if (rc != 1 || v != F.tag_value) return F.bump(ORC_FAIL_TAG), 0;
  }
  return 1;
}

struct PileCfg {
  uint64_t bin;
  bool use_fragment, collapse;
  bool shift_on = false;
  int32_t shift[4] = {0, 0, 0, 0};
  bool has_trunc = false, t_has5 = false, t_has3 = false;
  uint64_t t5 = 0, t3 = 0;
};

inline uint64_t sat_sub(uint64_t a, uint64_t b) { return a > b ? a - b : 0; }
inline uint64_t sat_add(uint64_t a, uint64_t b) { return a + b < a ? UINT64_MAX : a + b; }

// IntervalMaker::coords (genomic_intervals.rs:355-377)
bool coords(const Filter &F, const PileCfg &P, const Header &h, const Rec &r, uint64_t &s_out, uint64_t &e_out) {
  if (is_valid(F, r) != 1) return false;
  uint64_t start, end;
  uint32_t flag = r.flag();
  if (P.use_fragment) {  // coords_fragment :206-239
    if (!(flag & 0x40)) return false;
    int32_t t = r.tlen();
    if (t > 0) {
      if (r.pos() < 0) return false;
      start = (uint64_t)r.pos() + 1;
      end = start + (uint64_t)t;
    } else if (t < 0) {
      if (r.next_pos() < 0) return false;
      start = (uint64_t)r.next_pos() + 1;
      end = start + (uint64_t)(-(int64_t)t);
    } else
      return false;
  } else {  // coords_read :241-256
    if (r.pos() < 0) return false;
    start = (uint64_t)r.pos() + 1;
    end = start + cigar_span(r);
  }
  if (P.shift_on) {  // coords_shifted :258-350
    bool rev = flag & 0x10, first = flag & 0x40;
    auto apply_start = [&](int32_t v) { start = v >= 0 ? sat_sub(start, (uint64_t)v) : sat_add(start, (uint64_t)(-(int64_t)v)); };
    auto apply_end = [&](int32_t v) { end = v >= 0 ? sat_add(end, (uint64_t)v) : sat_sub(end, (uint64_t)(-(int64_t)v)); };
    int32_t five = P.shift[0], three = P.shift[1], five_r = P.shift[2], three_r = P.shift[3];
    if (rev && first) {
      apply_start(three);
      apply_end(five_r);
    } else if (rev && !first) {
      apply_start(three_r);
      apply_end(five_r);
    } else if (!rev && first) {
      apply_start(five);
      apply_end(three);
    } else {
      apply_start(five);
      apply_end(three_r);
    }
    int32_t ref_id = r.ref_id();
    if (ref_id < 0 || (size_t)ref_id >= h.lens.size()) return false;
    start = std::max<uint64_t>(start, 1);
    end = std::min<uint64_t>(h.lens[ref_id], end);
  }
  if (P.has_trunc) {  // truncate :379-402
    if (P.t_has5) start = sat_add(start, P.t5);
    if (P.t_has3) end = sat_sub(end, P.t3);
  }
  s_out = start;
  e_out = end;
  return true;
}

struct Run {
  int32_t ref_id;
  uint32_t val;
  uint64_t start, stop;
};

// bin_intervals (bam_utils.rs:48-91)
void bin_intervals(const Lapper &lp, int ref_id, uint64_t region_start, uint64_t region_end, uint64_t bin, bool collapse,
                   std::vector<Run> &out) {
  size_t first = out.size();
  uint64_t start = region_start;
  while (start < region_end) {
    uint64_t end = std::min(start + bin, region_end);
    uint32_t count = (uint32_t)lp.count(start, end);
    if (collapse && out.size() > first && out.back().val == count)
      out.back().stop = end;
    else
      out.push_back(Run{ref_id, count, start, end});
    start = end;
  }
}

struct ChunkReader {  // one indexed reader (bam::io::indexed_reader::Builder::build_from_path)
  Bgzf bz;
  Bai bai;
  ChunkReader(const std::string &bam) : bz(bam), bai(read_bai(bai_path_for(bam))) {}
};

// the per-chunk closure of pileup_genome (coverage_analysis.rs:454-507)
void process_chunk(const std::string &bam, const Region &reg, const Header &h, const Filter &F, const PileCfg &P,
                   ChunkReader *shared_reader, std::vector<Run> &out, std::vector<uint8_t> &recbuf) {
  std::unique_ptr<ChunkReader> own;
  ChunkReader *rd = shared_reader;
  if (!rd) {
    own = std::make_unique<ChunkReader>(bam);
    rd = own.get();
  }
  std::vector<Lapper::Iv> ivs;
  if ((size_t)reg.ref_id < rd->bai.refs.size()) {
    auto chunks = bai_query(rd->bai.refs[reg.ref_id], reg.start, reg.end);
    for (auto &ch : chunks) {
      rd->bz.seek(ch.beg);
      while (rd->bz.tell() < ch.end) {
        uint8_t b4[4];
        size_t g = rd->bz.read(b4, 4);
        if (g == 0) break;
        if (g != 4) fail("truncated BAM record");
        uint32_t bs = rd32(b4);
        if (bs < 32) fail("invalid BAM record size");
        if (recbuf.size() < bs) recbuf.resize(bs);
        if (rd->bz.read(recbuf.data(), bs) != bs) fail("truncated BAM record");
        Rec r{recbuf.data(), bs};
        // noodles query `intersects`: same reference, has start and end (span > 0), intervals intersect
        if (r.ref_id() != reg.ref_id || r.pos() < 0) continue;
        if (32 + r.l_read_name() + 4 * (uint64_t)r.n_cigar() > bs) continue;  // undecodable: dropped by r.ok()
        uint64_t span = cigar_span(r);
        if (span == 0) continue;
        uint64_t a_start = (uint64_t)r.pos() + 1, a_end = a_start + span - 1;
        if (a_start > reg.end || a_end < reg.start) continue;
        uint64_t s, e;
        if (coords(F, P, h, r, s, e)) ivs.push_back(Lapper::Iv{s, e});
      }
    }
  }
  Lapper lp;
  lp.build(std::move(ivs));
  bin_intervals(lp, reg.ref_id, reg.start, reg.end, P.bin, P.collapse, out);
}

PileCfg make_pile_cfg(const orc_pileup_cfg *c) {
  PileCfg P;
  P.bin = c->bin_size;
  P.use_fragment = c->use_fragment != 0;
  P.collapse = c->collapse != 0;
  if (c->has_shift) {
    for (int i = 0; i < 4; i++) P.shift[i] = c->shift[i];
    P.shift_on = P.shift[0] || P.shift[1] || P.shift[2] || P.shift[3];
  }
  if (c->has_truncate) {
    P.has_trunc = true;
    P.t_has5 = c->trunc_has5 != 0;
    P.t5 = c->trunc5;
    P.t_has3 = c->trunc_has3 != 0;
    P.t3 = c->trunc3;
  }
  return P;
}

// rayon par_iter over chunks (coverage_analysis.rs:448-527) -> runs sorted by (ref_id, start)
std::vector<Run> run_chunks(const orc_pileup_cfg *c, const Filter &F, const BamStats &st, const std::vector<Region> &regions) {
  PileCfg P = make_pile_cfg(c);
  if (P.bin == 0) fail("bin size must be > 0");
  int nt = c->n_threads > 0 ? c->n_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  nt = (int)std::min<size_t>((size_t)nt, std::max<size_t>(regions.size(), 1));
  std::vector<std::vector<Run>> per_thread(nt);
  std::atomic<size_t> next{0};
  std::mutex err_mu;
  std::string err;
  std::string bam = c->bam_path;
  auto work = [&](int t) {
    try {
      std::unique_ptr<ChunkReader> shared;
      if (!c->faithful_io) shared = std::make_unique<ChunkReader>(bam);
      std::vector<uint8_t> recbuf(1 << 16);
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= regions.size()) break;
        process_chunk(bam, regions[i], st.header, F, P, shared.get(), per_thread[t], recbuf);
      }
    } catch (const std::exception &e) {
      std::lock_guard<std::mutex> g(err_mu);
      err = e.what();
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(work, t);
  work(0);
  for (auto &t : th) t.join();
  if (!err.empty()) fail(err);
  std::vector<Run> all;
  size_t n = 0;
  for (auto &v : per_thread) n += v.size();
  all.reserve(n);
  for (auto &v : per_thread) all.insert(all.end(), v.begin(), v.end());
  std::sort(all.begin(), all.end(), [](const Run &a, const Run &b) { return a.ref_id != b.ref_id ? a.ref_id < b.ref_id : a.start < b.start; });
  return all;
}

void snapshot(const Filter &F, uint64_t out[ORC_N_COUNTERS]) {
  for (int i = 0; i < ORC_N_COUNTERS; i++) out[i] = F.stats[i].load();
}
uint64_t n_after_filtering(const uint64_t s[ORC_N_COUNTERS]) {
  uint64_t f = 0;
  for (int i = 1; i < ORC_N_COUNTERS; i++) f += s[i];
  return s[0] - f;
}

double scale_factor(int method, float base_scale, uint64_t bin, uint64_t n) {
  double base = (double)base_scale;
  if (method == 0) return base;
  if (n == 0) return 0.0;
  if (method == 2) return base * (1000000.0 / (double)n);
  double rpm = 1000000.0 / (double)n;
  double per_kb = 1000.0 / (double)bin;
  return base * rpm * per_kb;
}

// ryu `Buffer::format` text for a finite f64 (polars CsvWriter float text)
std::string fmt_f64(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  std::string out;
  if (std::signbit(v)) {
    out.push_back('-');
    v = -v;
  }
  if (v == 0.0) return out + "0.0";
  char buf[64];
  auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
  std::string sci(buf, res.ptr);
  size_t epos = sci.find('e');
  std::string mant = sci.substr(0, epos);
  int e10 = atoi(sci.c_str() + epos + 1);
  std::string digits;
  for (char ch : mant)
    if (ch != '.') digits.push_back(ch);
  int length = (int)digits.size();
  int kk = e10 + 1;
  int k = kk - length;
  if (0 <= k && kk <= 16) {
    out += digits;
    out.append((size_t)(kk - length), '0');
    out += ".0";
  } else if (0 < kk && kk <= 16) {
    out += digits.substr(0, (size_t)kk);
    out.push_back('.');
    out += digits.substr((size_t)kk);
  } else if (-5 < kk && kk <= 0) {
    out += "0.";
    out.append((size_t)(-kk), '0');
    out += digits;
  } else if (length == 1) {
    out += digits;
    out.push_back('e');
    out += std::to_string(kk - 1);
  } else {
    out.push_back(digits[0]);
    out.push_back('.');
    out += digits.substr(1);
    out.push_back('e');
    out += std::to_string(kk - 1);
  }
  return out;
}

// is_scaffold (bam_utils.rs:38-43): ^chrUn|_alt$|_random$
bool is_scaffold(const std::string &n) {
  auto ends = [&](const char *s) {
    size_t l = strlen(s);
    return n.size() >= l && n.compare(n.size() - l, l, s) == 0;
  };
  return n.compare(0, 5, "chrUn") == 0 || ends("_alt") || ends("_random");
}

struct FileOut {
  FILE *f;
  std::vector<char> buf;
  explicit FileOut(const std::string &p) {
    f = fopen(p.c_str(), "wb");
    if (!f) fail("cannot create " + p);
    buf.reserve(1 << 22);
  }
  void put(const std::string &s) {
    buf.insert(buf.end(), s.begin(), s.end());
    if (buf.size() > (1 << 22) - 256) flush();
  }
  void flush() {
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), f);
    buf.clear();
  }
  ~FileOut() {
    flush();
    fclose(f);
  }
};

template <class Fn>
int guarded(Fn fn) {
  try {
    fn();
    return 0;
  } catch (const OrcError &e) {
    g_err = e.what();
    return e.code;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 1;
  }
}

std::vector<Run> pileup_genome(const orc_pileup_cfg *cfg, const orc_filter_cfg *fc, uint64_t stats[ORC_N_COUNTERS],
                               BamStats *st_out = nullptr) {
  BamStats st = bam_stats_new(cfg->bam_path);
  auto F = make_filter(fc, st.header);
  uint64_t L = chunk_length(st, cfg->bin_size);
  std::vector<Region> regions;
  for (size_t i = 0; i < st.header.names.size(); i++)
    if (st.in_stats[i]) chunks_for_ref(st, (int)i, L, regions);
  auto runs = run_chunks(cfg, *F, st, regions);
  snapshot(*F, stats);
  if (st_out) *st_out = std::move(st);
  return runs;
}

}  // namespace

extern "C" {

const char *orc_last_error(void) { return g_err.c_str(); }
void orc_free(void *p) { free(p); }

int orc_pileup_runs(const orc_pileup_cfg *cfg, const orc_filter_cfg *fc, orc_run **runs, uint64_t *n_runs,
                    uint64_t stats[ORC_N_COUNTERS], double *norm_factor) {
  return guarded([&] {
    auto r = pileup_genome(cfg, fc, stats);
    *n_runs = r.size();
    *runs = (orc_run *)malloc(std::max<size_t>(r.size(), 1) * sizeof(orc_run));
    for (size_t i = 0; i < r.size(); i++) (*runs)[i] = orc_run{r[i].ref_id, r[i].val, r[i].start, r[i].stop};
    if (norm_factor) *norm_factor = scale_factor(cfg->norm_method, cfg->scale_factor, cfg->bin_size, n_after_filtering(stats));
  });
}

int orc_pileup_chromosome(const orc_pileup_cfg *cfg, const orc_filter_cfg *fc, const char *chrom, orc_run **runs,
                          uint64_t *n_runs, uint64_t stats[ORC_N_COUNTERS]) {
  return guarded([&] {
    BamStats st = bam_stats_new(cfg->bam_path);
    auto it = st.header.name_to_id.find(chrom);
    if (it == st.header.name_to_id.end() || !st.in_stats[it->second]) fail(std::string("Chromosome ") + chrom + " not found", 2);
    auto F = make_filter(fc, st.header);
    uint64_t L = chunk_length(st, cfg->bin_size);
    std::vector<Region> regions;
    chunks_for_ref(st, it->second, L, regions);
    auto r = run_chunks(cfg, *F, st, regions);
    snapshot(*F, stats);
    *n_runs = r.size();
    *runs = (orc_run *)malloc(std::max<size_t>(r.size(), 1) * sizeof(orc_run));
    for (size_t i = 0; i < r.size(); i++) (*runs)[i] = orc_run{r[i].ref_id, r[i].val, r[i].start, r[i].stop};
  });
}

int orc_signal_for_chromosome(const orc_pileup_cfg *cfg, const orc_filter_cfg *fc, const char *chrom, float **out,
                              uint64_t *n) {
  return guarded([&] {
    BamStats st = bam_stats_new(cfg->bam_path);
    auto it = st.header.name_to_id.find(chrom);
    if (it == st.header.name_to_id.end()) fail(std::string("Chromosome ") + chrom + " not found in BAM file.", 2);
    uint64_t chrom_size = st.header.lens[it->second];
    orc_pileup_cfg c2 = *cfg;
    c2.norm_method = 0;
    c2.collapse = 1;
    c2.has_shift = 0;
    c2.has_truncate = 0;
    auto F = make_filter(fc, st.header);
    uint64_t L = chunk_length(st, c2.bin_size);
    std::vector<Region> regions;
    if (!st.in_stats[it->second]) fail(std::string("Chromosome ") + chrom + " not found");
    chunks_for_ref(st, it->second, L, regions);
    auto r = run_chunks(&c2, *F, st, regions);
    float *arr = (float *)calloc(std::max<uint64_t>(chrom_size, 1), sizeof(float));
    for (auto &iv : r) {
      float v = (float)iv.val * cfg->scale_factor;
      for (uint64_t i = iv.start; i < iv.stop && i < chrom_size; i++) arr[i] = v;
    }
    *out = arr;
    *n = chrom_size;
  });
}

// write_bedgraph (coverage_analysis.rs:43-75): optional scaffold drop, sort by (chrom string, start), one text row per
// run.  polars serialises CSV chunks on its thread pool, so the rows are formatted in parallel here as well and written
// in order.
static void write_bedgraph_rows(const orc_pileup_cfg *cfg, const BamStats &st, const std::vector<Run> &runs, double factor, const char *out_path) {
  size_t nref = st.header.names.size();
  std::vector<int> order(nref);
  for (size_t i = 0; i < nref; i++) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return st.header.names[a] < st.header.names[b]; });
  std::vector<size_t> first(nref + 1, 0);
  for (auto &r : runs) first[r.ref_id + 1]++;
  for (size_t i = 0; i < nref; i++) first[i + 1] += first[i];
  struct Piece {
    int rid;
    size_t lo, hi;
  };
  std::vector<Piece> pieces;
  const size_t CH = 1 << 18;
  for (int rid : order) {
    if (cfg->ignore_scaffold && is_scaffold(st.header.names[rid])) continue;
    for (size_t lo = first[rid]; lo < first[rid + 1]; lo += CH) pieces.push_back(Piece{rid, lo, std::min(first[rid + 1], lo + CH)});
  }
  std::vector<std::string> text(pieces.size());
  int nt = cfg->n_threads > 0 ? cfg->n_threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min<int>(nt, (int)std::max<size_t>(pieces.size(), 1)));
  std::atomic<size_t> next{0};
  auto work = [&] {
    for (;;) {
      size_t k = next.fetch_add(1);
      if (k >= pieces.size()) break;
      const Piece &pc = pieces[k];
      const std::string &name = st.header.names[pc.rid];
      std::string &out = text[k];
      out.reserve((pc.hi - pc.lo) * (name.size() + 32));
      for (size_t i = pc.lo; i < pc.hi; i++) {
        const Run &r = runs[i];
        out += name;
        out.push_back('\t');
        out += std::to_string(r.start);
        out.push_back('\t');
        out += std::to_string(r.stop);
        out.push_back('\t');
        out += fmt_f64((double)r.val * factor);
        out.push_back('\n');
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
  FileOut fo(out_path);
  for (auto &t : text) {
    fo.flush();
    if (!t.empty()) fwrite(t.data(), 1, t.size(), fo.f);
  }
}

int orc_pileup_to_bedgraph(const orc_pileup_cfg *cfg, const orc_filter_cfg *fc, const char *out_path,
                           uint64_t stats[ORC_N_COUNTERS], double *norm_factor) {
  return guarded([&] {
    BamStats st;
    auto runs = pileup_genome(cfg, fc, stats, &st);
    double factor = scale_factor(cfg->norm_method, cfg->scale_factor, cfg->bin_size, n_after_filtering(stats));
    if (norm_factor) *norm_factor = factor;
    write_bedgraph_rows(cfg, st, runs, factor, out_path);
  });
}

// bench helper: the same job restricted to a few chromosomes (a bounded sample of a large BAM): pileup of their chunks,
// normalisation from their counters, bedGraph written.
int orc_pileup_chroms_to_bedgraph(const orc_pileup_cfg *cfg, const orc_filter_cfg *fc, const char *const *chroms, int32_t n_chroms,
                                  const char *out_path, uint64_t stats[ORC_N_COUNTERS], double *norm_factor) {
  return guarded([&] {
    BamStats st = bam_stats_new(cfg->bam_path);
    auto F = make_filter(fc, st.header);
    uint64_t L = chunk_length(st, cfg->bin_size);
    std::vector<Region> regions;
    for (int i = 0; i < n_chroms; i++) {
      auto it = st.header.name_to_id.find(chroms[i]);
      if (it == st.header.name_to_id.end() || !st.in_stats[it->second]) fail(std::string("Chromosome ") + chroms[i] + " not found", 2);
      chunks_for_ref(st, it->second, L, regions);
    }
    auto runs = run_chunks(cfg, *F, st, regions);
    snapshot(*F, stats);
    double factor = scale_factor(cfg->norm_method, cfg->scale_factor, cfg->bin_size, n_after_filtering(stats));
    if (norm_factor) *norm_factor = factor;
    write_bedgraph_rows(cfg, st, runs, factor, out_path);
  });
}

uint64_t orc_collapse_equal_bins(int32_t *chrom, uint64_t *start, uint64_t *end, uint32_t *scores, uint64_t n_rows,
                                 uint32_t n_cols) {
  // sort by (chrom, start); a new group starts whenever chrom or any score differs from the previous row
  std::vector<uint64_t> idx(n_rows);
  for (uint64_t i = 0; i < n_rows; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return chrom[a] != chrom[b] ? chrom[a] < chrom[b] : start[a] < start[b]; });
  std::vector<int32_t> oc;
  std::vector<uint64_t> os, oe;
  std::vector<uint32_t> osc;
  for (uint64_t k = 0; k < n_rows; k++) {
    uint64_t i = idx[k];
    bool same = false;
    if (k > 0) {
      uint64_t j = idx[k - 1];
      same = chrom[i] == chrom[j] && memcmp(scores + i * n_cols, scores + j * n_cols, n_cols * sizeof(uint32_t)) == 0;
    }
    if (same) {
      os.back() = std::min(os.back(), start[i]);
      oe.back() = std::max(oe.back(), end[i]);
    } else {
      oc.push_back(chrom[i]);
      os.push_back(start[i]);
      oe.push_back(end[i]);
      osc.insert(osc.end(), scores + i * n_cols, scores + (i + 1) * n_cols);
    }
  }
  uint64_t m = oc.size();
  for (uint64_t i = 0; i < m; i++) {
    chrom[i] = oc[i];
    start[i] = os[i];
    end[i] = oe[i];
  }
  if (m) memcpy(scores, osc.data(), m * n_cols * sizeof(uint32_t));
  return m;
}

int orc_multi_to_tsv(const orc_pileup_cfg *cfgs, const orc_filter_cfg *filters, int32_t n_bams, const char *out_path) {
  return guarded([&] {
    if (n_bams < 1) fail("At least one BAM pileup is required");
    for (int i = 0; i < n_bams; i++) {
      if (cfgs[i].bin_size != cfgs[0].bin_size) fail("All BAM pileups must have the same bin size");
      if (cfgs[i].collapse) fail("All BAM pileups must have collapse set to false");
    }
    // sequential per-BAM raw pileups, full outer join on (chrom name, start, end): every BAM's rows are sorted by that key
    // (the order collapse_equal_bins sorts by), then merged; -1 == null (a key one BAM does not have)
    std::vector<std::vector<Run>> per_bam((size_t)n_bams);
    std::vector<std::vector<std::string>> names((size_t)n_bams);
    for (int b = 0; b < n_bams; b++) {
      uint64_t stats[ORC_N_COUNTERS];
      BamStats st;
      per_bam[b] = pileup_genome(&cfgs[b], &filters[b], stats, &st);
      names[b] = st.header.names;
      const std::vector<std::string> &nm = names[b];
      auto by_key = [&](const Run &x, const Run &y) {
        if (x.ref_id != y.ref_id) return nm[x.ref_id] < nm[y.ref_id];
        return x.start != y.start ? x.start < y.start : x.stop < y.stop;
      };
      // the rows arrive grouped by reference and ordered by start inside one: move whole references into name order, and only
      // fall back to a full sort if that did not give the key order
      {
        std::vector<size_t> first(nm.size() + 1, 0);
        for (const Run &r : per_bam[b]) first[(size_t)r.ref_id + 1]++;
        for (size_t r = 0; r < nm.size(); r++) first[r + 1] += first[r];
        std::vector<int> order(nm.size());
        for (size_t r = 0; r < nm.size(); r++) order[r] = (int)r;
        std::sort(order.begin(), order.end(), [&](int x, int y) { return nm[x] < nm[y]; });
        std::vector<size_t> dst0(nm.size(), 0);
        size_t off = 0;
        for (int r : order) {
          dst0[r] = off;
          off += first[r + 1] - first[r];
        }
        std::vector<Run> sorted(per_bam[b].size());
        std::vector<size_t> fill = dst0;
        for (const Run &r : per_bam[b]) sorted[fill[r.ref_id]++] = r;
        per_bam[b].swap(sorted);
      }
      if (!std::is_sorted(per_bam[b].begin(), per_bam[b].end(), by_key)) std::sort(per_bam[b].begin(), per_bam[b].end(), by_key);
    }
    // collapse_equal_bins over all sample columns (nulls compare as "different")
    FileOut fo(out_path);
    std::string line = "chrom\tstart\tend";
    for (int b = 0; b < n_bams; b++) {
      line.push_back('\t');
      line += cfgs[b].bam_path;
    }
    line.push_back('\n');
    fo.put(line);
    bool have = false;
    std::string gchrom;
    std::vector<int64_t> gv((size_t)n_bams), cur((size_t)n_bams);
    uint64_t gs = 0, ge = 0;
    auto emit = [&] {
      if (!have) return;
      line = gchrom + "\t" + std::to_string(gs) + "\t" + std::to_string(ge);
      for (int64_t x : gv) {
        line.push_back('\t');
        if (x >= 0) line += std::to_string(x);
      }
      line.push_back('\n');
      fo.put(line);
    };
    std::vector<size_t> at((size_t)n_bams, 0);
    auto key_less = [&](int x, int y) {  // row at[x] of BAM x < row at[y] of BAM y
      const Run &p = per_bam[x][at[x]], &q = per_bam[y][at[y]];
      const std::string &pn = names[x][p.ref_id], &qn = names[y][q.ref_id];
      if (pn != qn) return pn < qn;
      return p.start != q.start ? p.start < q.start : p.stop < q.stop;
    };
    for (;;) {
      int lead = -1;
      for (int b = 0; b < n_bams; b++)
        if (at[b] < per_bam[b].size() && (lead < 0 || key_less(b, lead))) lead = b;
      if (lead < 0) break;
      const Run k = per_bam[lead][at[lead]];
      const std::string &kchrom = names[lead][k.ref_id];
      for (int b = 0; b < n_bams; b++) {
        cur[b] = -1;
        if (at[b] < per_bam[b].size()) {
          const Run &q = per_bam[b][at[b]];
          if (q.start == k.start && q.stop == k.stop && names[b][q.ref_id] == kchrom) {
            cur[b] = q.val;
            at[b]++;
          }
        }
      }
      bool same = have && gchrom == kchrom;
      if (same)
        for (int b = 0; b < n_bams; b++)
          if (cur[b] < 0 || gv[b] < 0 || cur[b] != gv[b]) same = false;
      if (same) {
        gs = std::min(gs, k.start);
        ge = std::max(ge, k.stop);
      } else {
        emit();
        have = true;
        gchrom = kchrom;
        gv = cur;
        gs = k.start;
        ge = k.stop;
      }
    }
    emit();
  });
}

int orc_collapse_bedgraph(const char *in_path, const char *out_path) {
  return guarded([&] {
    std::vector<std::string> lines;
    if (!read_lines(in_path, lines)) fail(std::string("cannot open ") + in_path);
    struct Row {
      std::string chrom;
      int64_t start, end;
      double score;
    };
    std::vector<Row> rows;
    for (auto &l : lines) {
      auto parts = split_ws(l);
      if (parts.empty() || parts[0][0] == '#') continue;
      if (parts.size() < 4) continue;
      char *e1, *e2, *e3;
      errno = 0;
      long long s = strtoll(parts[1].c_str(), &e1, 10), e = strtoll(parts[2].c_str(), &e2, 10);
      double sc = strtod(parts[3].c_str(), &e3);
      if (*e1 || *e2 || *e3 || e1 == parts[1].c_str() || e2 == parts[2].c_str() || e3 == parts[3].c_str()) fail("invalid bedGraph line: " + l);
      rows.push_back(Row{parts[0], s, e, sc});
    }
    std::stable_sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.chrom != b.chrom ? a.chrom < b.chrom : a.start < b.start; });
    FileOut fo(out_path);
    const double eps = 1e-6;
    size_t i = 0;
    while (i < rows.size()) {
      size_t j = i + 1;
      int64_t mn = rows[i].start, mx = rows[i].end;
      while (j < rows.size() && rows[j].chrom == rows[j - 1].chrom && rows[j].start == rows[j - 1].end &&
             !(std::fabs(rows[j].score - rows[j - 1].score) > eps)) {
        mn = std::min(mn, rows[j].start);
        mx = std::max(mx, rows[j].end);
        j++;
      }
      fo.put(rows[i].chrom + "\t" + std::to_string(mn) + "\t" + std::to_string(mx) + "\t" + fmt_f64(rows[i].score) + "\n");
      i = j;
    }
  });
}

double orc_scale_factor(int32_t method, float base_scale, uint64_t bin_size, uint64_t n_reads) {
  return scale_factor(method, base_scale, bin_size, n_reads);
}

int orc_bam_stats(const char *bam_path, uint64_t bin_size, uint64_t *n_mapped, uint64_t *n_unmapped,
                  uint64_t *genome_len, uint64_t *chunk_len, uint64_t *n_chunks, int32_t *n_refs) {
  return guarded([&] {
    BamStats st = bam_stats_new(bam_path);
    *n_mapped = st.n_mapped;
    *n_unmapped = st.n_unmapped;
    *genome_len = st.genome_len;
    *n_refs = (int32_t)st.header.names.size();
    uint64_t L = chunk_length(st, bin_size);
    *chunk_len = L;
    uint64_t n = 0;
    if (L)
      for (size_t i = 0; i < st.header.names.size(); i++)
        if (st.in_stats[i]) n += (st.header.lens[i] + L - 1) / L;
    *n_chunks = n;
  });
}

int orc_format_f64(double v, char *buf, size_t cap) {
  std::string s = fmt_f64(v);
  if (s.size() + 1 > cap) return -1;
  memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}

int orc_count_records(const char *bam_path, uint64_t *n_records, uint64_t *inflated_bytes, uint64_t *n_blocks) {
  return guarded([&] {
    Bgzf bz(bam_path);
    read_header(bz);
    uint64_t n = 0;
    std::vector<uint8_t> buf(1 << 16);
    for (;;) {
      uint8_t b4[4];
      size_t g = bz.read(b4, 4);
      if (g == 0) break;
      if (g != 4) fail("truncated BAM record");
      uint32_t bs = rd32(b4);
      if (buf.size() < bs) buf.resize(bs);
      if (bz.read(buf.data(), bs) != bs) fail("truncated BAM record");
      n++;
    }
    *n_records = n;
    *inflated_bytes = bz.bytes_inflated;
    *n_blocks = bz.n_blocks_loaded;
  });
}

}  // extern "C"

#include "bam_writers.inl"
