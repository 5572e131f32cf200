// bigwig_tools.inl — host side of `bigwig-compare` / `bigwig-aggregate` (bamnado/src/bigwig_compare.rs:616-790); included at the
// end of engine.cpp (it uses the engine's device buffers, staging ring and BigWig encoder).  The host reads the file layout only:
// header, chromosome B+ tree, R-tree leaves.  Section payloads go to the device as they are, are inflated by the zlib variant of
// K1, become items, and every output bin is computed there (bigwig_tools.cuh); the output is written by the same device encoder
// as `bam-coverage -o x.bw`.

namespace {

struct BwInput {
  std::string path;
  int fd = -1;
  const uint8_t *map = nullptr;
  uint64_t size = 0;
  uint32_t ubuf = 0;                                                 // uncompressBufSize: 0 = sections are stored raw
  std::vector<std::tuple<std::string, uint32_t, uint32_t>> chroms;   // (name, id, length), sorted by id
  std::vector<std::pair<uint64_t, uint64_t>> leaves;                 // (offset, size) of every data section, in file order
  ~BwInput() {
    if (map) munmap(const_cast<uint8_t *>(map), size);
    if (fd >= 0) close(fd);
  }
  template <class T>
  T rd(uint64_t off) const {
    if (off + sizeof(T) > size) fail("truncated BigWig: " + path);
    T v;
    memcpy(&v, map + off, sizeof(T));
    return v;
  }
};

void bw_walk_chroms(BwInput &in, uint64_t off, uint32_t key_size, int depth) {
  if (depth > 16) fail("malformed chromosome tree in " + in.path);
  const uint8_t is_leaf = in.rd<uint8_t>(off);
  const uint16_t count = in.rd<uint16_t>(off + 2);
  off += 4;
  for (uint16_t k = 0; k < count; k++, off += key_size + 8) {
    if (off + key_size + 8 > in.size) fail("truncated BigWig: " + in.path);
    if (is_leaf) {
      std::string name((const char *)in.map + off, key_size);
      name.resize(strnlen(name.c_str(), key_size));
      in.chroms.emplace_back(name, in.rd<uint32_t>(off + key_size), in.rd<uint32_t>(off + key_size + 4));
    } else {
      bw_walk_chroms(in, in.rd<uint64_t>(off + key_size), key_size, depth + 1);
    }
  }
}

void bw_walk_rtree(BwInput &in, uint64_t off, int depth) {
  if (depth > 32) fail("malformed R-tree in " + in.path);
  const uint8_t is_leaf = in.rd<uint8_t>(off);
  const uint16_t count = in.rd<uint16_t>(off + 2);
  off += 4;
  for (uint16_t k = 0; k < count; k++) {
    if (is_leaf) {
      in.leaves.emplace_back(in.rd<uint64_t>(off + 16), in.rd<uint64_t>(off + 24));
      off += 32;
    } else {
      bw_walk_rtree(in, in.rd<uint64_t>(off + 16), depth + 1);
      off += 24;
    }
  }
}

void bw_open(const std::string &path, BwInput &in) {
  in.path = path;
  in.fd = ::open(path.c_str(), O_RDONLY);
  if (in.fd < 0) fail("File not found: " + path);
  struct stat st;
  if (fstat(in.fd, &st) != 0 || st.st_size < 64) fail("not a BigWig file: " + path);
  in.size = (uint64_t)st.st_size;
  void *m = mmap(nullptr, in.size, PROT_READ, MAP_SHARED, in.fd, 0);
  if (m == MAP_FAILED) fail("cannot map " + path);
  in.map = (const uint8_t *)m;
  if (in.rd<uint32_t>(0) != 0x888FFC26u) fail("not a BigWig file (or a byte-swapped one, which is not supported): " + path);
  const uint64_t chrom_off = in.rd<uint64_t>(8), index_off = in.rd<uint64_t>(24);
  in.ubuf = in.rd<uint32_t>(52);
  if (in.rd<uint32_t>(chrom_off) != 0x78CA8C91u) fail("bad chromosome tree in " + path);
  const uint32_t key_size = in.rd<uint32_t>(chrom_off + 8);
  if (in.rd<uint64_t>(chrom_off + 16)) bw_walk_chroms(in, chrom_off + 32, key_size, 0);
  std::sort(in.chroms.begin(), in.chroms.end(), [](const auto &x, const auto &y) { return std::get<1>(x) < std::get<1>(y); });
  if (in.rd<uint32_t>(index_off) != 0x2468ACE0u) fail("bad R-tree in " + path);
  if (in.rd<uint64_t>(index_off + 8)) bw_walk_rtree(in, index_off + 48, 0);
  std::sort(in.leaves.begin(), in.leaves.end());
}

// All intervals of one BigWig as one device array; range_of[name] = its items of that chromosome (load_intervals_for_chrom).
void bw_load_items(DeviceCtx &ctx, cudaStream_t st, const BwInput &in, DevBuf &d_items, std::map<std::string, std::pair<uint64_t, uint64_t>> &range_of) {
  range_of.clear();
  for (auto &c : in.chroms) range_of[std::get<0>(c)] = {0, 0};
  const uint32_t ns = (uint32_t)in.leaves.size();
  d_items.ensure(16);
  if (ns == 0) return;
  const bool compressed = in.ubuf != 0;
  if (in.ubuf > 65536u) fail("BigWig sections larger than 64 KiB are not supported: " + in.path);
  DevBuf d_comp, d_coff, d_uoff, d_out, d_len, d_status, d_err, d_work, d_sec_off, d_headers, d_item_off, d_flag, d_range;
  auto release = [&] {
    for (DevBuf *b : {&d_comp, &d_coff, &d_uoff, &d_out, &d_len, &d_status, &d_err, &d_work, &d_sec_off, &d_headers, &d_item_off, &d_flag, &d_range}) b->release();
  };
  try {
    // section payloads back to back (16-byte aligned starts), as the inflate kernel and the header reads want them
    std::vector<uint64_t> coff(ns + 1, 0), uoff(ns + 1, 0), sec_off(ns);
    for (uint32_t k = 0; k < ns; k++) {
      if (in.leaves[k].first + in.leaves[k].second > in.size) fail("truncated BigWig: " + in.path);
      coff[k + 1] = coff[k] + ((in.leaves[k].second + 15) & ~15ull);
    }
    std::vector<uint8_t> packed(coff[ns]);
    {
      const int threads = (int)std::min<uint32_t>(std::max(2u, host_cores_share()), ns / 64 + 1);
      std::vector<std::thread> th;
      for (int t = 0; t < threads; t++)
        th.emplace_back([&, t] {
          for (uint32_t k = (uint32_t)t; k < ns; k += (uint32_t)threads) memcpy(packed.data() + coff[k], in.map + in.leaves[k].first, in.leaves[k].second);
        });
      for (auto &t : th) t.join();
    }
    d_comp.ensure(packed.size() + COMP_PAD);
    upload_bytes(ctx, packed.data(), packed.size(), d_comp.as<uint8_t>(), st);
    const uint8_t *sec_data = d_comp.as<uint8_t>();
    std::vector<uint32_t> raw_len(ns);
    if (compressed) {
      const uint64_t stride = ((uint64_t)in.ubuf + 15) & ~15ull;
      for (uint32_t k = 0; k <= ns; k++) uoff[k] = (uint64_t)k * stride;
      d_out.ensure((size_t)ns * stride + 64);
      d_len.ensure((size_t)ns * 4);
      d_status.ensure((size_t)ns * 4);
      d_err.ensure(16);
      d_work.ensure(16);
      CU(cudaMemsetAsync(d_err.p, 0, 16, st));
      CU(cudaMemsetAsync(d_work.p, 0, 16, st));
      CU(cudaMemsetAsync(d_len.p, 0, (size_t)ns * 4, st));
      // a zlib stream ends where its Adler-32 ends, not where its padded slot ends: the streams are given as (begin, end) pairs
      std::vector<uint64_t> cexact(2 * (size_t)ns);
      for (uint32_t k = 0; k < ns; k++) cexact[2 * k] = coff[k], cexact[2 * k + 1] = coff[k] + in.leaves[k].second;
      upload(d_coff, cexact.data(), cexact.size());
      upload(d_uoff, uoff.data(), uoff.size());
      InflateArgs ia{};
      ia.comp = d_comp.as<uint8_t>();
      ia.blk_coff = d_coff.as<uint64_t>();
      ia.blk_uoff = d_uoff.as<uint64_t>();
      ia.out = d_out.as<uint8_t>();
      ia.status = d_status.as<int32_t>();
      ia.error_flag = d_err.as<int32_t>();
      ia.work_counter = d_work.as<uint32_t>();
      ia.crc = ctx.d_crc;
      ia.n_blocks = ns;
      ia.verify_crc = 1;
      ia.out_len = d_len.as<uint32_t>();
      ia.coff_stride = 2;
      CU(launch_inflate_zlib(ia, ctx.sm_count, st));
      std::vector<int32_t> status(ns);
      CU(cudaMemcpyAsync(status.data(), d_status.p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(raw_len.data(), d_len.p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      for (uint32_t k = 0; k < ns; k++)
        if (status[k] != 0) fail("cannot inflate section " + std::to_string(k) + " of " + in.path + " (status " + std::to_string(status[k]) + ")");
      for (uint32_t k = 0; k < ns; k++) sec_off[k] = uoff[k];
      sec_data = d_out.as<uint8_t>();
    } else {
      for (uint32_t k = 0; k < ns; k++) sec_off[k] = coff[k], raw_len[k] = (uint32_t)in.leaves[k].second;
    }
    upload(d_sec_off, sec_off.data(), sec_off.size());
    d_headers.ensure((size_t)ns * 24);
    CU(launch_bwr_headers(sec_data, d_sec_off.as<uint64_t>(), ns, d_headers.as<uint32_t>(), ctx.sm_count, st));
    std::vector<uint32_t> hdr((size_t)ns * 6);
    CU(cudaMemcpyAsync(hdr.data(), d_headers.p, (size_t)ns * 24, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // item counts and chromosome ranges (sections are stored in (chromosome id, start) order)
    std::vector<uint64_t> item_off(ns + 1, 0);
    std::map<uint32_t, std::pair<uint64_t, uint64_t>> by_id;
    uint32_t last_id = 0;
    for (uint32_t k = 0; k < ns; k++) {
      const uint32_t id = hdr[(size_t)k * 6], type = hdr[(size_t)k * 6 + 5] & 0xffu, count = hdr[(size_t)k * 6 + 5] >> 16;
      const uint32_t unit = type == 1 ? 12u : type == 2 ? 8u : type == 3 ? 4u : 0u;
      if (!unit) fail("unknown section type in " + in.path);
      if (raw_len[k] < 24 + (uint64_t)count * unit) fail("short section in " + in.path);
      if (k && id < last_id) fail("sections out of chromosome order in " + in.path);
      last_id = id;
      item_off[k + 1] = item_off[k] + count;
      auto it = by_id.find(id);
      if (it == by_id.end()) by_id[id] = {item_off[k], item_off[k + 1]};
      else it->second.second = item_off[k + 1];
    }
    const uint64_t n_items = item_off[ns];
    d_items.ensure(std::max<uint64_t>(n_items, 1) * sizeof(BwItem));
    upload(d_item_off, item_off.data(), item_off.size());
    d_flag.ensure(16);
    CU(cudaMemsetAsync(d_flag.p, 0, 16, st));
    BwrEmitArgs ea{};
    ea.data = sec_data;
    ea.sec_off = d_sec_off.as<uint64_t>();
    ea.item_off = d_item_off.as<uint64_t>();
    ea.n_secs = ns;
    ea.items = d_items.as<BwItem>();
    ea.bad = d_flag.as<unsigned int>();
    CU(launch_bwr_emit_items(ea, ctx.sm_count, st));
    std::vector<uint64_t> ranges;
    for (auto &c : in.chroms) {
      auto it = by_id.find(std::get<1>(c));
      const std::pair<uint64_t, uint64_t> r = it == by_id.end() ? std::pair<uint64_t, uint64_t>{0, 0} : it->second;
      range_of[std::get<0>(c)] = r;
      ranges.push_back(r.first);
      ranges.push_back(r.second);
    }
    if (!in.chroms.empty()) {
      upload(d_range, ranges.data(), ranges.size());
      CU(launch_bwr_check_sorted(d_items.as<BwItem>(), d_range.as<uint64_t>(), (uint32_t)in.chroms.size(), d_flag.as<unsigned int>(), ctx.sm_count, st));
    }
    unsigned int bad = 0;
    CU(cudaMemcpyAsync(&bad, d_flag.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (bad) fail("BigWig with empty, overlapping or unsorted intervals is not supported: " + in.path);
  } catch (...) {
    release();
    throw;
  }
  release();
}

// compare_bigwigs_with_threads / aggregate_bigwigs_with_threads: per-bin values of every chromosome of the FIRST input, in its
// chromosome order, merged into items and written as a BigWig.
void bigwig_tool(const std::vector<std::string> &inputs, const std::string &out_path, int mode, uint32_t bin_size, double pseudocount,
                 const std::vector<double> &scales, int device) {
  if (inputs.empty()) fail("At least one BigWig file is required", ERR_INVALID);
  if ((int)inputs.size() > BWT_MAX_FILES) fail("too many BigWig inputs", ERR_INVALID);
  if (bin_size == 0) fail("bin_size must be > 0", ERR_INVALID);
  DeviceCtx &ctx = ctx_for(device);
  std::lock_guard<std::mutex> lock(ctx.mu);
  cudaStream_t st = ctx.sets[0].stream;
  const double t0 = now_ms();
  const size_t nf = inputs.size();
  std::vector<std::unique_ptr<BwInput>> in(nf);
  std::vector<DevBuf> d_items(nf), d_range(nf);
  DevBuf d_bin_first, d_chrom_len, d_bins, d_tiles, d_total, d_pair_off, d_pair_val, d_pair_w, d_out_items, d_cif;
  auto release = [&] {
    for (auto &b : d_items) b.release();
    for (auto &b : d_range) b.release();
    for (DevBuf *b : {&d_bin_first, &d_chrom_len, &d_bins, &d_tiles, &d_total, &d_pair_off, &d_pair_val, &d_pair_w, &d_out_items, &d_cif}) b->release();
  };
  try {
    std::vector<std::map<std::string, std::pair<uint64_t, uint64_t>>> range_of(nf);
    for (size_t f = 0; f < nf; f++) {
      in[f].reset(new BwInput());
      bw_open(inputs[f], *in[f]);
      bw_load_items(ctx, st, *in[f], d_items[f], range_of[f]);
    }
    const auto &chroms_in = in[0]->chroms;  // chromosome order of the first input (its ids)
    const uint32_t n_chrom = (uint32_t)chroms_in.size();
    std::vector<std::tuple<std::string, uint32_t, uint32_t>> chroms;
    std::vector<uint64_t> bin_first(n_chrom + 1, 0);
    std::vector<uint32_t> chrom_len(n_chrom);
    for (uint32_t c = 0; c < n_chrom; c++) {
      chroms.emplace_back(std::get<0>(chroms_in[c]), c, std::get<2>(chroms_in[c]));
      chrom_len[c] = std::get<2>(chroms_in[c]);
      bin_first[c + 1] = bin_first[c] + ((uint64_t)chrom_len[c] + bin_size - 1) / bin_size;
    }
    const uint64_t n_bins = bin_first[n_chrom];
    BwtArgs a{};
    a.n_files = (int)nf;
    a.mode = mode;
    a.pseudocount = pseudocount;
    a.n_chrom = n_chrom;
    a.bin_size = bin_size;
    a.n_bins = n_bins;
    for (size_t f = 0; f < nf; f++) {
      std::vector<uint64_t> r;
      for (uint32_t c = 0; c < n_chrom; c++) {
        auto it = range_of[f].find(std::get<0>(chroms[c]));
        if (it == range_of[f].end()) fail("chromosome " + std::get<0>(chroms[c]) + " is not in " + inputs[f]);
        r.push_back(it->second.first);
        r.push_back(it->second.second);
      }
      upload(d_range[f], r.data(), r.size());
      a.items[f] = d_items[f].as<BwItem>();
      a.range[f] = d_range[f].as<uint64_t>();
      a.scale[f] = scales[f];
    }
    upload(d_bin_first, bin_first.data(), bin_first.size());
    upload(d_chrom_len, chrom_len.data(), chrom_len.size());
    a.bin_first = d_bin_first.as<uint64_t>();
    a.chrom_len = d_chrom_len.as<uint32_t>();
    d_bins.ensure(std::max<uint64_t>(n_bins, 1) * 4);
    a.bins = d_bins.as<float>();
    d_total.ensure(16);
    const double t1 = now_ms();
    if (mode == BWT_MEDIAN) {
      if (n_bins > 0xffffffffull) fail("too many bins");
      d_pair_off.ensure(std::max<uint64_t>(n_bins, 1) * 8);
      a.pair_off = d_pair_off.as<uint64_t>();
      CU(launch_bwt_median_count(a, ctx.sm_count, st));
      CU(launch_u64_scan(a.pair_off, (uint32_t)n_bins, d_total.as<unsigned long long>(), st));
      unsigned long long n_pairs = 0;
      CU(cudaMemcpyAsync(&n_pairs, d_total.p, 8, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      d_pair_val.ensure(std::max<uint64_t>(n_pairs, 1) * 4);
      d_pair_w.ensure(std::max<uint64_t>(n_pairs, 1) * 4);
      a.pair_val = d_pair_val.as<float>();
      a.pair_w = d_pair_w.as<uint32_t>();
      CU(launch_bwt_median_fill(a, ctx.sm_count, st));
    } else {
      CU(launch_bwt_bins(a, ctx.sm_count, st));
    }
    // consecutive equal bins -> one item
    const uint64_t nt64 = (n_bins + COL_TILE - 1) / COL_TILE;
    if (nt64 > 0xffffffffull) fail("too many bins");
    a.n_tiles = (uint32_t)nt64;
    d_tiles.ensure(nt64 * 8 + 16);
    a.tile_count = d_tiles.as<uint64_t>();
    CU(launch_bwt_heads(a, ctx.sm_count, st));
    CU(launch_u64_scan(a.tile_count, a.n_tiles, d_total.as<unsigned long long>(), st));
    unsigned long long n_out = 0;
    CU(cudaMemcpyAsync(&n_out, d_total.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    d_out_items.ensure(std::max<uint64_t>(n_out, 1) * sizeof(BwItem));
    d_cif.ensure((n_chrom + 1) * 8);
    a.out_items = d_out_items.as<BwItem>();
    a.chrom_item_first = d_cif.as<uint64_t>();
    CU(launch_bwt_emit(a, ctx.sm_count, st));
    std::vector<uint64_t> cif(n_chrom + 1, 0);
    if (n_chrom) CU(cudaMemcpyAsync(cif.data(), d_cif.p, (size_t)n_chrom * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cif[n_chrom] = n_out;
    for (uint32_t c = n_chrom; c-- > 0;)
      if (bin_first[c + 1] == bin_first[c]) cif[c] = cif[c + 1];  // a zero-length chromosome owns no bins and no items
    const double t2 = now_ms();
    g_unmapper.reap();
    encode_bigwig_file(ctx, st, out_path, chroms, cif, d_out_items, t2);
    if (env_u64("BND_DEBUG_TIMING", 0))
      fprintf(stderr, "[bnd] bigwig tool (mode %d): %zu inputs, %llu bins, %llu items out; read + inflate + items %.2f ms, bins + merge %.2f ms, encode + write %.2f ms\n", mode,
              nf, (unsigned long long)n_bins, (unsigned long long)n_out, t1 - t0, t2 - t1, now_ms() - t2);
  } catch (...) {
    release();
    throw;
  }
  release();
}

}  // namespace

void bigwig_compare(const char *bw1, const char *bw2, const char *out, int comparison, uint32_t bin_size, const double *pseudocount, const double *scale1,
                    const double *scale2, int device) {
  if (!bw1 || !bw2 || !out) fail("missing BigWig path", ERR_INVALID);
  if (comparison != BWT_SUBTRACTION && comparison != BWT_RATIO && comparison != BWT_LOGRATIO) fail("unknown comparison", ERR_INVALID);
  // a small default pseudocount avoids division by zero in the ratio modes (bigwig_compare.rs:646)
  bigwig_tool({bw1, bw2}, out, comparison, bin_size, pseudocount ? *pseudocount : 1e-12, {scale1 ? *scale1 : 1.0, scale2 ? *scale2 : 1.0}, device);
}

void bigwig_aggregate(const char *const *paths, int n, const char *out, int mode, uint32_t bin_size, const double *pseudocount, const double *scale_factors,
                      int device) {
  if (n < 1 || !paths) fail("At least one BigWig file is required", ERR_INVALID);
  if (mode < BWT_SUM || mode > BWT_MEDIAN) fail("unknown aggregation mode", ERR_INVALID);
  std::vector<std::string> in;
  std::vector<double> sf;
  for (int i = 0; i < n; i++) {
    in.push_back(paths[i]);
    sf.push_back(scale_factors ? scale_factors[i] : 1.0);
  }
  bigwig_tool(in, out, mode, bin_size, pseudocount ? *pseudocount : 0.0, sf, device);
}

// A BigWig from explicit intervals sorted by (chromosome index, start): the device encoder on caller-provided rows (what a
// bedGraph -> BigWig conversion needs; the tests build their inputs with it).
void bigwig_from_intervals(const char *const *chrom_names, const uint32_t *chrom_lens, uint32_t n_chrom, const uint32_t *chrom_idx, const uint32_t *start,
                           const uint32_t *end, const float *value, uint64_t n, const char *out_path, int device) {
  if (!out_path || (n_chrom && (!chrom_names || !chrom_lens))) fail("missing argument", ERR_INVALID);
  std::vector<std::tuple<std::string, uint32_t, uint32_t>> chroms;
  for (uint32_t c = 0; c < n_chrom; c++) chroms.emplace_back(chrom_names[c], c, chrom_lens[c]);
  std::vector<uint64_t> first(n_chrom + 1, 0);
  std::vector<BwItem> items(n);
  for (uint64_t i = 0; i < n; i++) {
    if (chrom_idx[i] >= n_chrom) fail("chromosome index out of range", ERR_INVALID);
    if (end[i] <= start[i] || end[i] > chrom_lens[chrom_idx[i]]) fail("interval out of range for BigWig", ERR_INVALID);
    if (i && (chrom_idx[i] < chrom_idx[i - 1] || (chrom_idx[i] == chrom_idx[i - 1] && start[i] < end[i - 1])))
      fail("intervals must be sorted by (chromosome, start) and must not overlap", ERR_INVALID);
    first[chrom_idx[i] + 1]++;
    items[i] = BwItem{start[i], end[i], value[i]};
  }
  for (uint32_t c = 0; c < n_chrom; c++) first[c + 1] += first[c];
  DeviceCtx &ctx = ctx_for(device);
  std::lock_guard<std::mutex> lock(ctx.mu);
  cudaStream_t st = ctx.sets[0].stream;
  DevBuf d_items;
  try {
    upload(d_items, items.data(), items.size());
    g_unmapper.reap();
    encode_bigwig_file(ctx, st, out_path, chroms, first, d_items, now_ms());
  } catch (...) {
    d_items.release();
    throw;
  }
  d_items.release();
}
