// bam_write.inl — host side of `split`, `split-exogenous` and `modify` (bamnado/src/bam_splitter.rs:54-120,
// spike_in_analysis.rs:162-400, bam_modifier.rs:55-230); included by engine.cpp in front of Pileup::run_pass, which hands every
// indexed segment to BamWriteJob::segment instead of the pile-up kernel while a job is attached.
//
// The records a job keeps are appended to one device stream per output file (class), in file order.  When a stream could
// overflow with the next segment, the pipeline is drained and the stream's complete 65 280-byte blocks are compressed on the
// device (bam_write.cuh), packed back to back, copied out through the pinned bounce buffers and written; the tail (less than a
// block) moves to the front.  The host writes nothing but what the device produced, plus the constant 28-byte EOF block.

namespace {

// add_bamnado_program_group (bam_utils.rs:757-780): @PG ID:bamnado PN:bamnado VN:<version> CL:<command line>.  noodles'
// Programs::add links the new program to every leaf of the existing @PG chains (PP), suffixing the id when it is taken.
std::string sam_text_with_program(const std::string &text_in, const std::string &command_line) {
  std::string text = text_in;
  while (!text.empty() && text.back() == '\0') text.pop_back();
  if (!text.empty() && text.back() != '\n') text.push_back('\n');
  std::vector<std::string> ids, pps;
  size_t pos = 0;
  while (pos < text.size()) {
    size_t eol = text.find('\n', pos);
    if (eol == std::string::npos) eol = text.size();
    if (text.compare(pos, 4, "@PG\t") == 0) {
      std::string id, pp;
      size_t f = pos + 4;
      while (f < eol) {
        size_t t = text.find('\t', f);
        if (t == std::string::npos || t > eol) t = eol;
        if (text.compare(f, 3, "ID:") == 0) id = text.substr(f + 3, t - f - 3);
        if (text.compare(f, 3, "PP:") == 0) pp = text.substr(f + 3, t - f - 3);
        f = t + 1;
      }
      ids.push_back(id);
      pps.push_back(pp);
    }
    pos = eol + 1;
  }
  std::string fields = "\tPN:bamnado\tVN:" BND_BAMNADO_VERSION;
  if (!command_line.empty()) fields += "\tCL:" + command_line;
  std::vector<std::string> leaves;
  for (size_t i = 0; i < ids.size(); i++)
    if (std::find(pps.begin(), pps.end(), ids[i]) == pps.end()) leaves.push_back(ids[i]);
  if (leaves.empty()) {
    text += "@PG\tID:bamnado" + fields + "\n";
  } else {
    for (const std::string &leaf : leaves) {
      std::string id = "bamnado";
      if (std::find(ids.begin(), ids.end(), id) != ids.end()) id += "-" + leaf;
      ids.push_back(id);
      text += "@PG\tID:" + id + fields + "\tPP:" + leaf + "\n";
    }
  }
  return text;
}

std::vector<uint8_t> bam_header_bytes(const BamHeader &h, const std::string &text) {
  std::vector<uint8_t> out;
  auto put32 = [&](uint32_t v) {
    for (int i = 0; i < 4; i++) out.push_back((uint8_t)(v >> (8 * i)));
  };
  out.insert(out.end(), {'B', 'A', 'M', 1});
  put32((uint32_t)text.size());
  out.insert(out.end(), text.begin(), text.end());
  put32((uint32_t)h.names.size());
  for (size_t r = 0; r < h.names.size(); r++) {
    put32((uint32_t)h.names[r].size() + 1);
    out.insert(out.end(), h.names[r].begin(), h.names[r].end());
    out.push_back(0);
    put32((uint32_t)h.lens[r]);
  }
  return out;
}

const uint8_t BGZF_EOF[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};

struct ClassStream {
  std::string path;
  int fd = -1;
  uint64_t file_off = 0;
  DevBuf buf, first_rec;
  uint64_t cap = 0;      // bytes of `buf` the stream may use
  uint64_t ub = 0;       // upper bound of the bytes in the stream (the exact length lives on the device)
  uint64_t raw_bytes = 0, blocks = 0;
  std::vector<uint8_t> *mem = nullptr;  // bnd_bgzf_compress: the blocks go to memory instead of a file
};

}  // namespace

struct BamWriteJob {
  DeviceCtx *ctx = nullptr;
  int mode = BAMW_SPLIT, tn5 = 0, n_classes = 1;
  bool store_only = false;
  ClassStream cls[BAMW_MAX_CLASSES];
  DevBuf d_class_len, d_counts, d_ref_exo, d_ref_len;
  DevBuf rec_cls[MAX_SETS], tile_sum[MAX_SETS];
  cudaEvent_t ev_scan[MAX_SETS] = {};
  int prev_set = -1;
  DevBuf d_slots, d_sizes, d_offsets, d_tokens, d_blob[2];
  cudaStream_t wstream = nullptr;  // D2H of a packed batch, issued by the writer thread while the next batch is compressed
  FilterDev fdev{};
  int32_t n_ref = 0;
  double compress_ms = 0, write_ms = 0;
  uint64_t flushes = 0;
  static constexpr uint32_t BATCH = 4096;  // BGZF blocks compressed per launch (256 MiB of slots)

  ~BamWriteJob() {
    for (auto &c : cls) {
      if (c.fd >= 0) close(c.fd);
      c.buf.release();
      c.first_rec.release();
    }
    for (DevBuf *b : {&d_class_len, &d_counts, &d_ref_exo, &d_ref_len, &d_slots, &d_sizes, &d_offsets, &d_tokens, &d_blob[0], &d_blob[1]}) b->release();
    if (wstream) cudaStreamDestroy(wstream);
    for (int i = 0; i < MAX_SETS; i++) {
      rec_cls[i].release();
      tile_sum[i].release();
      if (ev_scan[i]) cudaEventDestroy(ev_scan[i]);
    }
  }

  void write_all(ClassStream &c, const uint8_t *p, uint64_t n) {
    if (c.mem) {
      c.mem->insert(c.mem->end(), p, p + n);
      c.file_off += n;
      return;
    }
    // One pwrite stream fills page-cache pages at ~4 GB/s on the B200 hosts, and parallel pwrites to one file serialise on the
    // inode lock (profiles/r1_sweeps.md).  Large pieces therefore go through a shared mapping of the file's new tail, filled by
    // several threads with non-temporal copies; the pages are backed first unless the file system has ample room (a store into
    // a hole of a full file system would raise SIGBUS instead of an error).
    const unsigned nt = n >= (8u << 20) && env_u64("BND_BAM_WRITE_MMAP", 1) ? std::min<unsigned>(24u, std::max(2u, host_cores_share() * 3 / 2)) : 1u;
    if (nt > 1 && ftruncate(c.fd, (off_t)(c.file_off + n)) == 0) {
      bool backed = true;
      struct statvfs vfs;
      const bool roomy = fstatvfs(c.fd, &vfs) == 0 && (double)vfs.f_bavail * (double)vfs.f_frsize >= 2.0 * (double)n + 1e9;
      if (!roomy && posix_fallocate(c.fd, (off_t)c.file_off, (off_t)n) != 0) backed = false;
      const uint64_t a0 = c.file_off & ~4095ull, delta = c.file_off - a0;
      void *m = backed ? mmap(nullptr, n + delta, PROT_READ | PROT_WRITE, MAP_SHARED, c.fd, (off_t)a0) : MAP_FAILED;
      if (m != MAP_FAILED) {
        uint8_t *dst = (uint8_t *)m + delta;
        std::vector<std::thread> th;
        auto piece = [&](unsigned k) {
          const uint64_t a = (n * k / nt) & ~63ull, b = k + 1 == nt ? n : (n * (k + 1) / nt) & ~63ull;
          stream_copy(dst + a, p + a, b - a);
        };
        for (unsigned k = 1; k < nt; k++) th.emplace_back(piece, k);
        piece(0);
        for (auto &t : th) t.join();
        g_unmapper.submit((char *)m, n + delta);
        c.file_off += n;
        return;
      }
    }
    while (n) {
      ssize_t w = pwrite(c.fd, p, n, (off_t)c.file_off);
      if (w <= 0) fail("write error on " + c.path);
      p += w, n -= (uint64_t)w, c.file_off += (uint64_t)w;
    }
  }

  // a packed batch on the device -> pinned bounce buffers -> file; runs on the writer thread with its own stream
  void drain_blob(ClassStream &c, const uint8_t *d_src, uint64_t total, uint64_t CH) {
    CU(cudaSetDevice(ctx->device));
    PinnedBounce &pb = ctx->bounce;
    uint64_t pending_len[PinnedBounce::N] = {0, 0};
    uint64_t k = 0;
    for (uint64_t off = 0; off < total; off += CH, k++) {  // the copy of chunk k overlaps the write of chunk k - 1
      const int slot = (int)(k % PinnedBounce::N);
      const uint64_t len = std::min<uint64_t>(CH, total - off);
      CU(cudaMemcpyAsync(pb.buf[slot], d_src + off, len, cudaMemcpyDeviceToHost, wstream));
      CU(cudaEventRecord(pb.ev[slot], wstream));
      pending_len[slot] = len;
      if (k > 0) {
        const int prev = (int)((k - 1) % PinnedBounce::N);
        CU(cudaEventSynchronize(pb.ev[prev]));
        write_all(c, pb.buf[prev], pending_len[prev]);
      }
    }
    if (k > 0) {
      const int prev = (int)((k - 1) % PinnedBounce::N);
      CU(cudaEventSynchronize(pb.ev[prev]));
      write_all(c, pb.buf[prev], pending_len[prev]);
    }
  }

  // blocks [0, n_blocks) of a device stream -> compressed, packed, written behind the file's current end.  Batches of BATCH blocks:
  // while batch k is copied out and written by the writer thread, batch k + 1 is compressed and packed into the other blob.
  void compress_and_write(ClassStream &c, const uint8_t *d_raw, uint64_t raw_len, const uint32_t *d_first_rec, cudaStream_t st) {
    const uint64_t n_blocks = (raw_len + BGZF_RAW - 1) / BGZF_RAW;
    const uint32_t batch = (uint32_t)std::min<uint64_t>(BATCH, std::max<uint64_t>(n_blocks, 1));
    d_slots.ensure((size_t)batch * BGZF_SLOT);
    d_sizes.ensure((size_t)batch * 4);
    d_offsets.ensure((size_t)batch * 8);
    std::vector<uint32_t> sizes(batch);
    std::vector<uint64_t> offs(batch);
    const uint64_t CH = std::max<uint64_t>(env_u64("BND_TEXT_CHUNK_BYTES", 64ull << 20), 1u << 20);
    ctx->bounce.ensure(CH);
    if (!wstream) CU(cudaStreamCreateWithFlags(&wstream, cudaStreamNonBlocking));
    std::thread writer;
    std::exception_ptr werr;
    double w_busy = 0;
    auto join_writer = [&] {
      if (writer.joinable()) writer.join();
      if (werr) {
        std::exception_ptr e = werr;
        werr = nullptr;
        std::rethrow_exception(e);
      }
    };
    try {
      uint32_t which = 0;
      for (uint64_t b0 = 0; b0 < n_blocks; b0 += batch, which ^= 1u) {
        const uint32_t nb = (uint32_t)std::min<uint64_t>(batch, n_blocks - b0);
        const double t0 = now_ms();
        BamwDeflateArgs da{};
        da.raw = d_raw + b0 * BGZF_RAW;
        da.raw_len = std::min<uint64_t>(raw_len - b0 * BGZF_RAW, (uint64_t)nb * BGZF_RAW);
        da.n_blocks = nb;
        da.first_rec = d_first_rec ? d_first_rec + b0 : nullptr;
        da.out = d_slots.as<uint8_t>();
        da.out_size = d_sizes.as<uint32_t>();
        d_tokens.ensure((size_t)bamw_deflate_grid(nb, ctx->sm_count) * BGZF_RAW * 4);
        da.tokens = d_tokens.as<uint32_t>();
        da.crc = ctx->d_crc;
        da.store_only = store_only ? 1 : 0;
        CU(launch_bamw_deflate(da, ctx->sm_count, st));
        CU(cudaMemcpyAsync(sizes.data(), d_sizes.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        uint64_t total = 0;
        for (uint32_t k = 0; k < nb; k++) {
          if (sizes[k] < 28 || sizes[k] > BGZF_SLOT) fail("BGZF compression failed on the device");
          offs[k] = total;
          total += sizes[k];
        }
        // the blob about to be overwritten was handed to the writer two batches ago: with one writer at a time it is free once
        // the previous write has been joined, which also keeps the file offsets in order
        join_writer();
        DevBuf &blob = d_blob[which];
        blob.ensure(total + 64);
        CU(cudaMemcpyAsync(d_offsets.p, offs.data(), (size_t)nb * 8, cudaMemcpyHostToDevice, st));
        BwCompactArgs ca{};
        ca.staging = d_slots.as<uint8_t>();
        ca.comp_size = d_sizes.as<uint32_t>();
        ca.offset = d_offsets.as<uint64_t>();
        ca.n_sections = nb;
        ca.stride = BGZF_SLOT;
        ca.blob = blob.as<uint8_t>();
        CU(launch_bw_compact(ca, ctx->sm_count, st));
        CU(cudaStreamSynchronize(st));
        compress_ms += now_ms() - t0;
        const uint8_t *src = blob.as<uint8_t>();
        writer = std::thread([this, &c, src, total, CH, &werr, &w_busy] {
          const double t1 = now_ms();
          try {
            drain_blob(c, src, total, CH);
          } catch (...) {
            werr = std::current_exception();
          }
          w_busy += now_ms() - t1;
        });
        c.blocks += nb;
      }
      join_writer();
    } catch (...) {
      if (writer.joinable()) writer.join();
      throw;
    }
    write_ms += w_busy;
    c.raw_bytes += raw_len;
  }

  void begin(DeviceCtx &c, const BamHeader &hdr, const FilterDev &f, const uint64_t *d_ref_len_ptr, int mode_, int tn5_, const std::string &exo_prefix,
             const std::vector<std::string> &paths, const std::string &command_line) {
    ctx = &c;
    mode = mode_, tn5 = tn5_;
    n_classes = mode == BAMW_SPLIT_EXOGENOUS ? 4 : 1;
    store_only = env_u64("BND_BAM_STORE_ONLY", 0) != 0;
    if ((int)paths.size() != n_classes) fail("wrong number of output paths", ERR_INVALID);
    fdev = f;
    n_ref = (int32_t)hdr.names.size();
    (void)d_ref_len_ptr;
    cudaStream_t st = c.sets[0].stream;
    upload(d_ref_len, hdr.lens.data(), hdr.lens.size());
    std::vector<uint8_t> exo(std::max<size_t>(hdr.names.size(), 1), 0);
    if (mode == BAMW_SPLIT_EXOGENOUS)
      for (size_t r = 0; r < hdr.names.size(); r++) exo[r] = hdr.names[r].compare(0, exo_prefix.size(), exo_prefix) == 0 ? 1 : 0;
    upload(d_ref_exo, exo.data(), exo.size());
    d_class_len.ensure(BAMW_MAX_CLASSES * 8);
    d_counts.ensure(BW_N_COUNTERS * 8);
    CU(cudaMemsetAsync(d_class_len.p, 0, BAMW_MAX_CLASSES * 8, st));
    CU(cudaMemsetAsync(d_counts.p, 0, BW_N_COUNTERS * 8, st));
    for (int i = 0; i < MAX_SETS; i++) CU(cudaEventCreateWithFlags(&ev_scan[i], cudaEventDisableTiming));
    const uint64_t cap0 = std::max<uint64_t>(env_u64("BND_BAM_STREAM_BYTES", 2ull << 30), 4 * (uint64_t)BGZF_RAW);
    const std::vector<uint8_t> header = bam_header_bytes(hdr, sam_text_with_program(hdr.text, command_line));
    DevBuf d_hdr;
    try {
      upload(d_hdr, header.data(), header.size());
      for (int k = 0; k < n_classes; k++) {
        ClassStream &s = cls[k];
        s.path = paths[k];
        s.fd = ::open(s.path.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
        if (s.fd < 0) fail("cannot create " + s.path);
        s.cap = mode == BAMW_SPLIT_EXOGENOUS && k == 2 ? 4 * (uint64_t)BGZF_RAW : cap0;  // "both genomes" never receives a record
        s.buf.ensure(s.cap + 64);
        s.first_rec.ensure((s.cap / BGZF_RAW + 4) * 4);
        CU(cudaMemsetAsync(s.first_rec.p, 0xff, (s.cap / BGZF_RAW + 4) * 4, st));
        CU(cudaMemsetAsync(s.first_rec.p, 0, 4, st));
        compress_and_write(s, d_hdr.as<uint8_t>(), header.size(), nullptr, st);  // the header in its own block(s), like htslib and noodles
        s.raw_bytes = 0, s.blocks = 0;
      }
    } catch (...) {
      d_hdr.release();
      throw;
    }
    d_hdr.release();
  }

  // Everything queued has to be complete (the caller synchronised the segment streams).
  void flush(bool final) {
    cudaStream_t st = ctx->sets[0].stream;
    uint64_t len[BAMW_MAX_CLASSES];
    CU(cudaMemcpy(len, d_class_len.p, sizeof len, cudaMemcpyDeviceToHost));
    for (int k = 0; k < n_classes; k++) {
      ClassStream &s = cls[k];
      if (len[k] > s.cap) fail("internal error: BAM output stream overflow");
      const uint64_t nb = final ? (len[k] + BGZF_RAW - 1) / BGZF_RAW : len[k] / BGZF_RAW;
      if (nb) compress_and_write(s, s.buf.as<uint8_t>(), std::min<uint64_t>(len[k], nb * BGZF_RAW), s.first_rec.as<uint32_t>(), st);
      const uint64_t rem = len[k] - std::min<uint64_t>(len[k], nb * BGZF_RAW);
      if (nb) {
        if (rem) CU(cudaMemcpyAsync(s.buf.p, s.buf.as<uint8_t>() + nb * BGZF_RAW, rem, cudaMemcpyDeviceToDevice, st));  // rem < one block <= nb blocks: no overlap
        CU(cudaMemcpyAsync(s.first_rec.p, s.first_rec.as<uint32_t>() + nb, 4, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemsetAsync(s.first_rec.as<uint32_t>() + 1, 0xff, (s.cap / BGZF_RAW + 3) * 4, st));
        if (rem == 0) CU(cudaMemsetAsync(s.first_rec.p, 0, 4, st));  // the next record starts the next block
      }
      len[k] = rem;
      s.ub = rem;
      if (final) {
        write_all(s, BGZF_EOF, sizeof BGZF_EOF);
        if (ftruncate(s.fd, (off_t)s.file_off) != 0 || close(s.fd) != 0) {
          s.fd = -1;
          fail("write error on " + s.path);
        }
        s.fd = -1;
      }
    }
    CU(cudaMemcpyAsync(d_class_len.p, len, sizeof len, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    flushes++;
  }

  // room for `bytes` more in every stream?  (called before a segment is queued)
  void before_segment(uint64_t bytes) {
    bool need = false;
    for (int k = 0; k < n_classes; k++)
      if (cls[k].ub + bytes > cls[k].cap && !(mode == BAMW_SPLIT_EXOGENOUS && k == 2)) need = true;
    if (!need) return;
    for (auto &s : ctx->sets) CU(cudaStreamSynchronize(s.stream));
    // the bound counts every inflated byte of every segment; what the filter kept is known on the device: with the exact
    // lengths most alarms are false (a filter that keeps a third of the reads flushed 23 times for 14 GB at the BASELINE
    // size), and a flush only happens once a stream is at least half full
    uint64_t len[BAMW_MAX_CLASSES];
    CU(cudaMemcpy(len, d_class_len.p, sizeof len, cudaMemcpyDeviceToHost));
    need = false;
    for (int k = 0; k < n_classes; k++) {
      cls[k].ub = len[k];
      if (mode == BAMW_SPLIT_EXOGENOUS && k == 2) continue;
      if (len[k] + bytes > cls[k].cap || len[k] > cls[k].cap / 2) need = true;
    }
    if (!need) return;
    flush(false);
    for (int k = 0; k < n_classes; k++) {
      ClassStream &s = cls[k];
      if (mode == BAMW_SPLIT_EXOGENOUS && k == 2) continue;
      if (s.ub + bytes <= s.cap) continue;
      // one segment alone exceeds the stream: a larger buffer, the tail carried over
      const uint64_t cap = (s.ub + bytes) * 2;
      DevBuf nb, nf;
      nb.ensure(cap + 64);
      nf.ensure((cap / BGZF_RAW + 4) * 4);
      CU(cudaMemset(nf.p, 0xff, (cap / BGZF_RAW + 4) * 4));
      CU(cudaMemcpy(nf.p, s.first_rec.p, 4, cudaMemcpyDeviceToDevice));
      if (s.ub) CU(cudaMemcpy(nb.p, s.buf.p, s.ub, cudaMemcpyDeviceToDevice));
      s.buf.release();
      s.first_rec.release();
      s.buf = nb, s.first_rec = nf, s.cap = cap;
    }
  }

  // segment `S` has been indexed on its stream: classify, continue the class scans behind the previous segment, gather
  void segment(SegSet &S, int set_index, uint64_t bytes) {
    cudaStream_t st = S.stream;
    const size_t max_rec = S.rec_off.cap / 8;
    rec_cls[set_index].ensure(max_rec + 16);
    tile_sum[set_index].ensure((max_rec / BAMW_TILE + 2) * BAMW_MAX_CLASSES * 8);
    BamwArgs a{};
    a.buf = S.infl.as<uint8_t>();
    a.rec_off = S.rec_off.as<uint64_t>();
    a.st = S.st;
    a.f = fdev;
    a.mode = mode, a.tn5_shift = tn5, a.n_classes = n_classes;
    a.n_ref = n_ref;
    a.ref_len = d_ref_len.as<uint64_t>();
    a.ref_is_exo = d_ref_exo.as<uint8_t>();
    a.rec_cls = rec_cls[set_index].as<uint8_t>();
    a.tile_sum = tile_sum[set_index].as<uint64_t>();
    a.class_len = d_class_len.as<uint64_t>();
    a.counts = d_counts.as<unsigned long long>();
    for (int k = 0; k < n_classes; k++) a.stream[k] = cls[k].buf.as<uint8_t>(), a.first_rec[k] = cls[k].first_rec.as<uint32_t>();
    if (prev_set >= 0) CU(cudaStreamWaitEvent(st, ev_scan[prev_set], 0));  // class_len is chained from segment to segment
    CU(launch_bamw_segment(a, ctx->sm_count, st));
    CU(cudaEventRecord(ev_scan[set_index], st));
    prev_set = set_index;
    CU(launch_bamw_gather(a, ctx->sm_count, st));
    for (int k = 0; k < n_classes; k++) cls[k].ub += bytes;
  }
};
