// bam_writers.inl — CPU restatement of the BAM-writing tools (TEST INFRASTRUCTURE ONLY, part of the oracle; included at the end
// of bamnado_oracle.cpp).  Follows bamnado/src/bam_splitter.rs:54-120 (BamFilterer::split_async), bam_modifier.rs:55-230
// (BamModifier::run) and spike_in_analysis.rs:162-400 (BamSplitter), record by record, and writes BGZF with zlib.
//
// Pinned by the reference: the @PG record (ID/PN bamnado, VN = crate version; bam_splitter.rs:203-236, bam_utils.rs:803-829).
// Unpinned against the real crates (noodles is not vendored): the order of contigs in the output of split / modify (the
// reference iterates a HashMap, any order; header order here), the PP links noodles' Programs::add appends, and that a
// RecordBuf re-encoded by `modify --tn5-shift` equals the input record apart from pos, bin, mate pos and tlen.
namespace {

struct BgzfWriter {
  FILE *f;
  std::vector<uint8_t> buf;
  explicit BgzfWriter(const std::string &path) : f(fopen(path.c_str(), "wb")) {
    if (!f) fail("cannot create " + path);
  }
  ~BgzfWriter() {
    if (f) fclose(f);
  }
  void block(const uint8_t *p, size_t n) {
    std::vector<uint8_t> out(n + n / 100 + 128);
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) fail("zlib init failed");
    zs.next_in = const_cast<uint8_t *>(p);
    zs.avail_in = (uInt)n;
    zs.next_out = out.data() + 18;
    zs.avail_out = (uInt)(out.size() - 26);
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) fail("zlib deflate failed");
    const size_t c = zs.total_out;
    deflateEnd(&zs);
    const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(out.data(), hdr, 16);
    const uint32_t total = (uint32_t)(18 + c + 8), crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), p, (uInt)n);
    out[16] = (uint8_t)(total - 1), out[17] = (uint8_t)((total - 1) >> 8);
    for (int i = 0; i < 4; i++) out[18 + c + i] = (uint8_t)(crc >> (8 * i)), out[22 + c + i] = (uint8_t)((uint32_t)n >> (8 * i));
    if (fwrite(out.data(), 1, total, f) != total) fail("write error");
  }
  void write(const uint8_t *p, size_t n) {
    buf.insert(buf.end(), p, p + n);
    size_t off = 0;
    while (buf.size() - off >= 0xff00) block(buf.data() + off, 0xff00), off += 0xff00;
    buf.erase(buf.begin(), buf.begin() + (long)off);
  }
  void flush() {
    if (!buf.empty()) block(buf.data(), buf.size());
    buf.clear();
  }
  void finish() {
    flush();
    block(nullptr, 0);  // EOF marker: an empty block
    fclose(f);
    f = nullptr;
  }
};

std::string header_text_of(const std::string &bam) {
  Bgzf bz(bam);
  uint8_t b[8];
  if (bz.read(b, 8) != 8 || memcmp(b, "BAM\1", 4) != 0) fail("not a BAM file");
  std::string text(rd32(b + 4), '\0');
  if (bz.read((uint8_t *)&text[0], text.size()) != text.size()) fail("truncated header text");
  return text;
}

// add_bamnado_program_group (bam_utils.rs:757-780)
std::string text_with_program(std::string text, const std::string &cl) {
  while (!text.empty() && text.back() == '\0') text.pop_back();
  if (!text.empty() && text.back() != '\n') text += '\n';
  std::vector<std::string> ids, pps;
  std::stringstream ss(text);
  std::string line;
  while (std::getline(ss, line)) {
    if (line.rfind("@PG\t", 0) != 0) continue;
    std::string id, pp, field;
    std::stringstream ls(line.substr(4));
    while (std::getline(ls, field, '\t')) {
      if (field.rfind("ID:", 0) == 0) id = field.substr(3);
      if (field.rfind("PP:", 0) == 0) pp = field.substr(3);
    }
    ids.push_back(id);
    pps.push_back(pp);
  }
  std::string fields = "\tPN:bamnado\tVN:0.9.0";
  if (!cl.empty()) fields += "\tCL:" + cl;
  std::vector<std::string> leaves;
  for (auto &id : ids)
    if (std::find(pps.begin(), pps.end(), id) == pps.end()) leaves.push_back(id);
  if (leaves.empty()) return text + "@PG\tID:bamnado" + fields + "\n";
  for (auto &leaf : leaves) {
    std::string id = "bamnado";
    if (std::find(ids.begin(), ids.end(), id) != ids.end()) id += "-" + leaf;
    ids.push_back(id);
    text += "@PG\tID:" + id + fields + "\tPP:" + leaf + "\n";
  }
  return text;
}

void write_bam_header(BgzfWriter &w, const Header &h, const std::string &text) {
  std::vector<uint8_t> out = {'B', 'A', 'M', 1};
  auto put32 = [&](uint32_t v) {
    for (int i = 0; i < 4; i++) out.push_back((uint8_t)(v >> (8 * i)));
  };
  put32((uint32_t)text.size());
  out.insert(out.end(), text.begin(), text.end());
  put32((uint32_t)h.names.size());
  for (size_t r = 0; r < h.names.size(); r++) {
    put32((uint32_t)h.names[r].size() + 1);
    out.insert(out.end(), h.names[r].begin(), h.names[r].end());
    out.push_back(0);
    put32((uint32_t)h.lens[r]);
  }
  w.write(out.data(), out.size());
  w.flush();  // the header has its own block(s)
}

uint32_t reg2bin(int64_t beg, int64_t end) {
  --end;
  if (beg >> 14 == end >> 14) return (uint32_t)(4681 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (uint32_t)(585 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (uint32_t)(73 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (uint32_t)(9 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (uint32_t)(1 + (beg >> 26));
  return 0;
}

void put_rec(BgzfWriter &w, const uint8_t *payload, uint32_t bs) {
  uint8_t b4[4] = {(uint8_t)bs, (uint8_t)(bs >> 8), (uint8_t)(bs >> 16), (uint8_t)(bs >> 24)};
  w.write(b4, 4);
  w.write(payload, bs);
}

// split_async / run: for every contig, the records of query(contig:1-length) that pass is_valid; modify optionally shifts them
void filter_to_bam(const std::string &bam, const orc_filter_cfg *fc, bool modify, bool tn5, const std::string &out_path, const std::string &cl,
                   uint64_t *n_written) {
  ChunkReader rd(bam);
  Header h = read_header(rd.bz);
  auto F = make_filter(fc, h);
  BgzfWriter w(out_path);
  write_bam_header(w, h, text_with_program(header_text_of(bam), cl));
  std::vector<uint8_t> rec(1 << 16);
  uint64_t written = 0;
  for (size_t ref = 0; ref < h.names.size(); ref++) {
    if (ref >= rd.bai.refs.size() || h.lens[ref] == 0) continue;
    for (auto &ch : bai_query(rd.bai.refs[ref], 1, h.lens[ref])) {
      rd.bz.seek(ch.beg);
      while (rd.bz.tell() < ch.end) {
        uint8_t b4[4];
        size_t g = rd.bz.read(b4, 4);
        if (g == 0) break;
        if (g != 4) fail("truncated BAM record");
        const uint32_t bs = rd32(b4);
        if (bs < 32) fail("invalid BAM record size");
        if (rec.size() < bs) rec.resize(bs);
        if (rd.bz.read(rec.data(), bs) != bs) fail("truncated BAM record");
        Rec r{rec.data(), bs};
        if (r.ref_id() != (int32_t)ref || r.pos() < 0) continue;
        if (32 + r.l_read_name() + 4 * (uint64_t)r.n_cigar() > bs) continue;
        const uint64_t span = cigar_span(r);
        if (span == 0) continue;
        const uint64_t a_start = (uint64_t)r.pos() + 1, a_end = a_start + span - 1;
        if (a_start > h.lens[ref] || a_end < 1) continue;
        const int valid = is_valid(*F, r);
        if (valid < 0) fail("read filter error (`?` in the reference)");
        if (!valid) continue;
        if (modify && tn5) {  // bam_modifier.rs:103-219
          if (!(r.flag() & 0x2)) continue;  // "Skipping record": not written
          const bool rev = r.flag() & 0x10;
          int64_t start = std::max<int64_t>(r.pos(), 0), end = start + (int64_t)r.l_seq();
          if (rev) end -= 5;
          else start += 4;
          const int64_t chromsize = (int64_t)h.lens[ref];
          if (!(start >= 0 && end <= chromsize)) continue;
          if (!(start >= 0 && start <= chromsize)) continue;
          const int32_t tlen = r.tlen();
          const int32_t new_tlen = tlen < 0 ? (tlen > INT32_MAX - 9 ? INT32_MAX : tlen + 9) : tlen - 9;
          if (r.next_pos() < 0) fail("Missing mate alignment start");
          int64_t mate = r.next_pos();
          if (rev) {
            const int64_t adj = mate + 4;
            if (adj >= 0 && adj <= chromsize) mate = adj;
          }
          auto st32 = [&](size_t off, uint32_t v) {
            for (int i = 0; i < 4; i++) rec[off + i] = (uint8_t)(v >> (8 * i));
          };
          st32(4, (uint32_t)start);
          const uint32_t bin = reg2bin(start, start + (int64_t)span);
          rec[10] = (uint8_t)bin, rec[11] = (uint8_t)(bin >> 8);
          st32(24, (uint32_t)mate);
          st32(28, (uint32_t)new_tlen);
        }
        put_rec(w, rec.data(), bs);
        written++;
      }
    }
  }
  w.finish();
  if (n_written) *n_written = written;
}

}  // namespace

extern "C" {

int orc_bam_split(const char *bam, const orc_filter_cfg *fc, const char *out_path, const char *command_line, uint64_t *n_written) {
  return guarded([&] { filter_to_bam(bam, fc, false, false, out_path, command_line ? command_line : "", n_written); });
}

int orc_bam_modify(const char *bam, const orc_filter_cfg *fc, int32_t tn5_shift, const char *out_path, const char *command_line, uint64_t *n_written) {
  return guarded([&] { filter_to_bam(bam, fc, true, tn5_shift != 0, out_path, command_line ? command_line : "", n_written); });
}

// BamSplitter::split (spike_in_analysis.rs:233-400); stats[10] in the order of SplitStats; paths[4] = endogenous, exogenous, both, unmapped
int orc_bam_split_exogenous(const char *bam, const orc_filter_cfg *fc, const char *exogenous_prefix, const char *const *paths, const char *command_line,
                            uint64_t stats[10]) {
  return guarded([&] {
    Bgzf bz(bam);
    Header h = read_header(bz);
    auto F = make_filter(fc, h);
    const std::string text = text_with_program(header_text_of(bam), command_line ? command_line : ""), prefix = exogenous_prefix;
    std::vector<std::unique_ptr<BgzfWriter>> w;
    for (int k = 0; k < 4; k++) {
      w.emplace_back(new BgzfWriter(paths[k]));
      write_bam_header(*w[k], h, text);
    }
    std::vector<char> exo(h.names.size());
    for (size_t r = 0; r < h.names.size(); r++) exo[r] = h.names[r].compare(0, prefix.size(), prefix) == 0;
    uint64_t s[10] = {0};
    std::vector<uint8_t> rec(1 << 16);
    for (;;) {
      uint8_t b4[4];
      size_t g = bz.read(b4, 4);
      if (g == 0) break;
      if (g != 4) fail("truncated BAM record");
      const uint32_t bs = rd32(b4);
      if (bs < 32) fail("invalid BAM record size");
      if (rec.size() < bs) rec.resize(bs);
      if (bz.read(rec.data(), bs) != bs) fail("truncated BAM record");
      Rec r{rec.data(), bs};
      s[0]++;
      const uint32_t flag = r.flag();
      if (flag & 0x4) s[1]++, put_rec(*w[3], rec.data(), bs);
      else if (flag & 0x200) s[2]++, put_rec(*w[3], rec.data(), bs);
      else if (flag & 0x400) s[3]++, put_rec(*w[3], rec.data(), bs);
      else if (flag & 0x100) s[4]++, put_rec(*w[3], rec.data(), bs);
      else {
        const int valid = is_valid(*F, r);
        if (valid < 0 || (valid == 1 && (size_t)r.ref_id() >= h.names.size())) fail("read filter error (`?` in the reference)");
        if (!valid) {
          s[5]++;
          continue;
        }
        if (exo[r.ref_id()]) s[8]++, put_rec(*w[1], rec.data(), bs);
        else s[9]++, put_rec(*w[0], rec.data(), bs);
      }
    }
    for (auto &x : w) x->finish();
    for (int k = 0; k < 10; k++) stats[k] = s[k];
  });
}

}  // extern "C"
