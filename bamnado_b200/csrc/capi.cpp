// capi.cpp — the extern "C" boundary declared in include/bamnado_b200.h, a thin shell over bnd::Pileup.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bamnado_b200.h"
#include "engine.h"
#include "kernels.h"
#include "outputs.h"

namespace {
thread_local std::string g_err;
thread_local int g_err_class = BND_OK;

template <class Fn>
int guarded(Fn fn) {
  try {
    fn();
    g_err_class = BND_OK;
    return 0;
  } catch (const bnd::Error &e) {
    g_err = e.what();
    g_err_class = e.cls;
    return e.cls;
  } catch (const std::bad_alloc &) {
    g_err = "out of host memory";
    g_err_class = BND_ERR_RUNTIME;
    return BND_ERR_RUNTIME;
  } catch (const std::exception &e) {
    g_err = e.what();
    g_err_class = BND_ERR_RUNTIME;
    return BND_ERR_RUNTIME;
  }
}

bnd::Pileup *P(bnd_pileup *p) {
  if (!p) bnd::fail("null pileup handle", bnd::ERR_INVALID);
  return reinterpret_cast<bnd::Pileup *>(p);
}

static_assert(sizeof(bnd::RunRow) == sizeof(bnd_run), "row layout must match the C ABI");
static_assert(bnd::ST_N_COUNTERS == BND_N_COUNTERS, "counter layout must match the C ABI");

void rows_out(std::vector<bnd::RunRow> &rows, bnd_run **runs, uint64_t *n_runs) {
  *n_runs = rows.size();
  *runs = (bnd_run *)malloc(std::max<size_t>(rows.size(), 1) * sizeof(bnd_run));
  if (!*runs) throw std::bad_alloc();
  if (!rows.empty()) memcpy(*runs, rows.data(), rows.size() * sizeof(bnd_run));
}
}  // namespace

extern "C" {

const char *bnd_last_error(void) { return g_err.c_str(); }
int bnd_last_error_class(void) { return g_err_class; }
void bnd_free(void *p) { free(p); }
int bnd_abi_version(void) { return BND_ABI_VERSION; }
int bnd_device_count(void) { return bnd::device_count(); }
uint64_t bnd_launch_count(void) { return bnd::launch_count(); }

bnd_pileup *bnd_pileup_create(const bnd_pileup_cfg *cfg, const bnd_filter_cfg *filter) {
  bnd::Pileup *p = nullptr;
  int rc = guarded([&] { p = new bnd::Pileup(cfg, filter); });
  return rc == 0 ? reinterpret_cast<bnd_pileup *>(p) : nullptr;
}
void bnd_pileup_destroy(bnd_pileup *p) { delete reinterpret_cast<bnd::Pileup *>(p); }

int bnd_pileup_to_bedgraph(bnd_pileup *p, const char *out_path) {
  return guarded([&] { P(p)->to_bedgraph(out_path); });
}
int bnd_pileup_to_bigwig(bnd_pileup *p, const char *out_path) {
  return guarded([&] { bnd::pileup_to_bigwig(*P(p), out_path); });
}
int bnd_pileup_chromosome(bnd_pileup *p, const char *chrom, bnd_run **runs, uint64_t *n_runs) {
  return guarded([&] {
    std::vector<bnd::RunRow> rows;
    P(p)->chromosome(chrom, rows);
    rows_out(rows, runs, n_runs);
  });
}
int bnd_pileup_runs(bnd_pileup *p, bnd_run **runs, uint64_t *n_runs) {
  return guarded([&] {
    std::vector<bnd::RunRow> rows;
    P(p)->runs(rows);
    rows_out(rows, runs, n_runs);
  });
}
int bnd_pileup_norm_factor(bnd_pileup *p, double *factor) {
  return guarded([&] { *factor = P(p)->norm_factor(); });
}
int bnd_pileup_stats(bnd_pileup *p, uint64_t stats[BND_N_COUNTERS]) {
  return guarded([&] { P(p)->stats(stats); });
}
int bnd_pileup_accumulate(bnd_pileup *p) {
  return guarded([&] { P(p)->accumulate(); });
}
int bnd_pileup_stage(bnd_pileup *p) {
  return guarded([&] { P(p)->stage(); });
}
int bnd_pileup_accumulate_resident(bnd_pileup *p) {
  return guarded([&] { P(p)->accumulate_resident(); });
}
int bnd_pileup_set_total_stats(bnd_pileup *p, const uint64_t stats[BND_N_COUNTERS]) {
  return guarded([&] { P(p)->set_total_stats(stats); });
}
int bnd_pileup_finalize(bnd_pileup *p, uint64_t *n_runs) {
  return guarded([&] {
    uint64_t n = P(p)->finalize();
    if (n_runs) *n_runs = n;
  });
}
int bnd_pileup_bedgraph_text(bnd_pileup *p, char **text, uint64_t *n_bytes) {
  return guarded([&] {
    std::string s;
    P(p)->bedgraph_text(s);
    *n_bytes = s.size();
    *text = (char *)malloc(s.size() + 1);
    if (!*text) throw std::bad_alloc();
    memcpy(*text, s.data(), s.size());
    (*text)[s.size()] = 0;
  });
}
int bnd_pileup_timings(bnd_pileup *p, double ms[8]) {
  return guarded([&] {
    const bnd::Timings &t = P(p)->timings();
    ms[0] = t.inflate_ms, ms[1] = t.index_ms, ms[2] = t.pileup_ms, ms[3] = t.finalize_ms;
    ms[4] = t.h2d_ms, ms[5] = t.text_ms, ms[6] = t.wall_ms, ms[7] = t.read_ms;
  });
}
int bnd_pileup_shape(bnd_pileup *p, uint64_t shape[16]) {
  return guarded([&] {
    const bnd::Shape &s = P(p)->shape();
    memset(shape, 0, 16 * sizeof(uint64_t));
    shape[0] = s.records, shape[1] = s.comp_bytes, shape[2] = s.inflated_bytes, shape[3] = s.blocks;
    shape[4] = s.bins, shape[5] = s.launches, shape[6] = s.chunk_len, shape[7] = s.n_chunks, shape[8] = s.segments;
  });
}
int bnd_pileup_preallocate_output(bnd_pileup *p, const char *out_path) {
  return guarded([&] { P(p)->preallocate_output(out_path); });
}
int bnd_pileup_format_bedgraph(bnd_pileup *p, uint64_t *bytes_per_ref, uint64_t n_refs) {
  return guarded([&] {
    std::vector<uint64_t> v;
    P(p)->format_bedgraph(v);
    if (v.size() != n_refs) bnd::fail("bytes_per_ref must have one entry per reference of the BAM header", bnd::ERR_INVALID);
    for (size_t i = 0; i < v.size(); i++) bytes_per_ref[i] = v[i];
  });
}
int bnd_pileup_preallocated_bytes(bnd_pileup *p, uint64_t *n_bytes) {
  return guarded([&] { *n_bytes = P(p)->preallocated_bytes(); });
}
int bnd_pileup_write_bedgraph_pieces(bnd_pileup *p, const char *out_path, const uint64_t *bytes_all, int32_t world, int32_t rank, uint64_t preallocated) {
  return guarded([&] { P(p)->write_bedgraph_pieces(out_path, bytes_all, world, rank, preallocated); });
}

int bnd_signal_for_chromosome(const bnd_pileup_cfg *cfg, const bnd_filter_cfg *filter, const char *chrom, float **signal,
                              uint64_t *n) {
  return guarded([&] {
    // get_signal_for_chromosome builds BamPileup::new(.., Raw, scale, use_fragment, filter, collapse = true, ..)
    bnd_pileup_cfg c = *cfg;
    c.norm_method = 0;
    c.collapse = 1;
    c.has_shift = 0;
    c.has_truncate = 0;
    c.shard_world = 0;
    bnd::Pileup pile(&c, filter);
    std::vector<float> v;
    pile.signal(chrom, v);
    *n = v.size();
    *signal = (float *)malloc(std::max<size_t>(v.size(), 1) * sizeof(float));
    if (!*signal) throw std::bad_alloc();
    if (!v.empty()) memcpy(*signal, v.data(), v.size() * sizeof(float));
  });
}

int bnd_multi_to_tsv(const bnd_pileup_cfg *cfgs, const bnd_filter_cfg *filters, int32_t n_bams, const char *out_path) {
  return guarded([&] { bnd::multi_to_tsv(cfgs, filters, n_bams, out_path); });
}

int bnd_bin_counts(const bnd_pileup_cfg *cfgs, const bnd_filter_cfg *filters, int32_t n_samples, uint64_t **counts, uint64_t *n_bins,
                   char **chrom_names, uint64_t **chrom_len, uint64_t **chrom_first_row, uint64_t *n_chroms, uint64_t *library_sizes) {
  return guarded([&] {
    bnd::Pileup::BinCounts bc;
    bnd::Pileup::count_bins(cfgs, filters, n_samples, bc);
    std::string names;
    for (size_t i = 0; i < bc.chroms.size(); i++) {
      if (i) names.push_back('\n');
      names += bc.chroms[i];
    }
    auto dup = [](const void *src, size_t bytes) {
      void *m = malloc(bytes ? bytes : 1);
      if (!m) bnd::fail("out of host memory");
      if (bytes) memcpy(m, src, bytes);
      return m;
    };
    *counts = (uint64_t *)dup(bc.counts.data(), bc.counts.size() * 8);
    *n_bins = bc.n_bins;
    *chrom_names = (char *)dup(names.c_str(), names.size() + 1);
    *chrom_len = (uint64_t *)dup(bc.chrom_len.data(), bc.chrom_len.size() * 8);
    *chrom_first_row = (uint64_t *)dup(bc.chrom_first_row.data(), bc.chrom_first_row.size() * 8);
    *n_chroms = bc.chroms.size();
    for (size_t i = 0; i < bc.library_sizes.size(); i++) library_sizes[i] = bc.library_sizes[i];
  });
}

int bnd_multi_prepare(bnd_pileup *p, uint64_t *n_bins) {
  return guarded([&] { *n_bins = P(p)->multi_prepare(false); });
}
int bnd_multi_prepare_resident(bnd_pileup *p, uint64_t *n_bins) {
  return guarded([&] { *n_bins = P(p)->multi_prepare(true); });
}
int bnd_multi_change_bits(bnd_pileup *const *ps, int32_t n, void *dev_bits) {
  return guarded([&] {
    std::vector<bnd::Pileup *> v;
    for (int i = 0; ps && i < n; i++) v.push_back(P(ps[i]));
    bnd::Pileup::multi_change_bits(v.data(), (int)v.size(), (uint32_t *)dev_bits);
  });
}
int bnd_multi_or_bits(int32_t device, void *dev_dst, const void *dev_src, uint32_t n_src, uint64_t n_words) {
  return guarded([&] { bnd::Pileup::multi_or_bits(device, (uint32_t *)dev_dst, (const uint32_t *)dev_src, n_src, n_words); });
}
int bnd_multi_compact_columns(bnd_pileup *const *ps, int32_t n, const void *dev_bits, uint64_t *n_rows, const void **dev_vals) {
  return guarded([&] {
    std::vector<bnd::Pileup *> v;
    for (int i = 0; ps && i < n; i++) v.push_back(P(ps[i]));
    const uint32_t *vals = nullptr;
    *n_rows = bnd::Pileup::multi_compact_columns(v.data(), (int)v.size(), (const uint32_t *)dev_bits, &vals);
    if (dev_vals) *dev_vals = vals;
  });
}
int bnd_multi_format_tsv(bnd_pileup *p, const void *dev_cols, uint64_t col_stride, int32_t n_cols, const int32_t *col_order, uint64_t row_lo,
                         uint64_t row_hi, int32_t sizes_only, uint64_t *bytes_per_ref, uint64_t n_refs) {
  return guarded([&] {
    std::vector<uint64_t> v;
    P(p)->multi_format_tsv((const uint32_t *)dev_cols, col_stride, n_cols, col_order, row_lo, row_hi, v, sizes_only != 0);
    if (v.size() != n_refs) bnd::fail("n_refs does not match the BAM header", bnd::ERR_INVALID);
    for (size_t i = 0; i < v.size(); i++) bytes_per_ref[i] = v[i];
  });
}
int bnd_multi_write_tsv_pieces(bnd_pileup *p, const char *out_path, const char *header, const uint64_t *bytes_all, int32_t world, int32_t rank) {
  return guarded([&] { P(p)->multi_write_tsv_pieces(out_path, header, bytes_all, world, rank); });
}

int bnd_bigwig_compare(const char *bw1, const char *bw2, const char *out_path, int32_t comparison, uint32_t bin_size, const double *pseudocount,
                       const double *scale_factor_bw1, const double *scale_factor_bw2) {
  return guarded([&] { bnd::bigwig_compare(bw1, bw2, out_path, comparison, bin_size, pseudocount, scale_factor_bw1, scale_factor_bw2, -1); });
}
int bnd_bigwig_aggregate(const char *const *bw_paths, int32_t n, const char *out_path, int32_t method, uint32_t bin_size, const double *pseudocount,
                         const double *scale_factors) {
  return guarded([&] { bnd::bigwig_aggregate(bw_paths, n, out_path, 10 + method, bin_size, pseudocount, scale_factors, -1); });
}
int bnd_bigwig_from_intervals(const char *const *chrom_names, const uint32_t *chrom_lens, uint32_t n_chrom, const uint32_t *chrom_idx, const uint32_t *start,
                              const uint32_t *end, const float *value, uint64_t n, const char *out_path) {
  return guarded([&] { bnd::bigwig_from_intervals(chrom_names, chrom_lens, n_chrom, chrom_idx, start, end, value, n, out_path, -1); });
}
int bnd_bam_split(const char *bam_path, const bnd_filter_cfg *filter, const char *out_path, const char *command_line, int32_t device, uint64_t *counters) {
  return guarded([&] {
    if (!out_path) bnd::fail("missing output path", bnd::ERR_INVALID);
    bnd::bam_write_tool(bam_path, filter, bnd::BAMW_SPLIT, 0, nullptr, {out_path}, command_line, device, counters, nullptr);
  });
}
int bnd_bam_modify(const char *bam_path, const bnd_filter_cfg *filter, int32_t tn5_shift, const char *out_path, const char *command_line, int32_t device,
                   uint64_t *counters) {
  return guarded([&] {
    if (!out_path) bnd::fail("missing output path", bnd::ERR_INVALID);
    bnd::bam_write_tool(bam_path, filter, bnd::BAMW_MODIFY, tn5_shift ? 1 : 0, nullptr, {out_path}, command_line, device, counters, nullptr);
  });
}
int bnd_bam_split_exogenous(const char *bam_path, const bnd_filter_cfg *filter, const char *exogenous_prefix, const char *out_prefix, const char *command_line,
                            int32_t device, uint64_t *counters) {
  return guarded([&] {
    if (!out_prefix || !exogenous_prefix) bnd::fail("missing output prefix or exogenous prefix", bnd::ERR_INVALID);
    std::vector<std::string> paths;
    for (const char *ext : {"endogenous.bam", "exogenous.bam", "both.bam", "unmapped.bam"}) paths.push_back(bnd::path_with_extension(out_prefix, ext));
    bnd::bam_write_tool(bam_path, filter, bnd::BAMW_SPLIT_EXOGENOUS, 0, exogenous_prefix, paths, command_line, device, counters, nullptr);
  });
}
int bnd_collapse_bedgraph(const char *in_path, const char *out_path) {
  return guarded([&] { bnd::collapse_bedgraph_file(in_path, out_path); });
}

double bnd_scale_factor(int32_t method, float base_scale, uint64_t bin_size, uint64_t n_reads) {
  // NormalizationMethod::scale_factor (signal_normalization.rs:50-74); method: 0 raw, 1 rpkm, 2 cpm
  const double base = (double)base_scale;
  if (method == 0) return base;
  if (n_reads == 0) return 0.0;
  const double rpm = 1000000.0 / (double)n_reads;
  if (method == 2) return base * rpm;
  const double per_kb = 1000.0 / (double)bin_size;
  return base * rpm * per_kb;
}

int bnd_pileup_refs(bnd_pileup *p, char **names, uint64_t **lens, uint64_t *n_refs) {
  return guarded([&] {
    const bnd::BamHeader &h = P(p)->header();
    std::string joined;
    for (size_t i = 0; i < h.names.size(); i++) {
      if (i) joined.push_back('\n');
      joined += h.names[i];
    }
    *n_refs = h.names.size();
    *names = (char *)malloc(joined.size() + 1);
    *lens = (uint64_t *)malloc(std::max<size_t>(h.lens.size(), 1) * sizeof(uint64_t));
    if (!*names || !*lens) throw std::bad_alloc();
    memcpy(*names, joined.c_str(), joined.size() + 1);
    for (size_t i = 0; i < h.lens.size(); i++) (*lens)[i] = h.lens[i];
  });
}

int bnd_bam_stats(const char *bam_path, uint64_t bin_size, uint64_t out[6]) {
  return guarded([&] {
    bnd_pileup_cfg c;
    memset(&c, 0, sizeof c);
    c.bam_path = bam_path;
    c.bin_size = bin_size;
    c.scale_factor = 1.f;
    c.device = -1;
    bnd::Pileup pile(&c, nullptr);
    const bnd::Geometry &g = pile.geometry();
    out[0] = g.n_mapped, out[1] = g.n_unmapped, out[2] = g.genome_len, out[3] = g.L, out[4] = g.total_chunks();
    out[5] = pile.header().names.size();
  });
}

int bnd_cli_main(int argc, const char *const *argv) { return bnd::cli_main(argc, argv); }

int bnd_bgzf_inflate(const uint8_t *bgzf, uint64_t n_bytes, uint8_t *out, uint64_t out_cap, uint64_t *out_len, int32_t *status,
                     uint64_t status_cap, uint64_t *n_blocks) {
  return guarded([&] {
    std::vector<uint8_t> o;
    std::vector<int32_t> st;
    bnd::device_inflate(bgzf, n_bytes, o, st, -1);
    if (o.size() > out_cap || st.size() > status_cap) bnd::fail("output buffer too small", bnd::ERR_INVALID);
    if (!o.empty()) memcpy(out, o.data(), o.size());
    if (!st.empty()) memcpy(status, st.data(), st.size() * sizeof(int32_t));
    *out_len = o.size();
    *n_blocks = st.size();
  });
}

int bnd_bgzf_compress(const uint8_t *data, uint64_t n_bytes, uint8_t **out, uint64_t *out_len) {
  return guarded([&] {
    if (!out || !out_len || (n_bytes && !data)) bnd::fail("missing argument", bnd::ERR_INVALID);
    std::vector<uint8_t> o;
    bnd::device_bgzf_compress(data, n_bytes, o, -1);
    uint8_t *p = (uint8_t *)malloc(o.size() ? o.size() : 1);
    if (!p) bnd::fail("out of memory");
    memcpy(p, o.data(), o.size());
    *out = p;
    *out_len = o.size();
  });
}

}  // extern "C"
