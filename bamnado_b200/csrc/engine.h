// engine.h — host engine of the B200 coverage path: streams compressed BAM bytes through pinned buffers into the
// kernels of kernels.cu and keeps the per-bin difference array, counters and compacted runs resident in HBM.
//
// It stands where BamPileup stands in the reference (bamnado/src/coverage_analysis.rs:136-760); the C ABI in
// capi.cpp is a thin shell over this class.
#pragma once
#include <stdint.h>

#include <memory>
#include <string>
#include <vector>

#include "../../include/bamnado_b200.h"
#include "bamio.h"
#include "filter_host.h"
#include "kernel_args.h"

namespace bnd {

struct DeviceCtx;  // per-process, per-device pools (streams, pinned staging ring, segment buffers)

struct Timings {
  double inflate_ms = 0, index_ms = 0, pileup_ms = 0, finalize_ms = 0, h2d_ms = 0, text_ms = 0, wall_ms = 0, read_ms = 0;
};
struct Shape {
  uint64_t records = 0, comp_bytes = 0, inflated_bytes = 0, blocks = 0, bins = 0, launches = 0, chunk_len = 0, n_chunks = 0, segments = 0;
};

class Pileup {
 public:
  // drop_scaffold_chunks: scaffold contigs are left out of the chunk list itself (bin_counts.rs:190-200)
  Pileup(const bnd_pileup_cfg *cfg, const bnd_filter_cfg *filter, bool drop_scaffold_chunks = false);
  ~Pileup();
  Pileup(const Pileup &) = delete;

  // --- whole-object operations (mirror BamPileup) ---
  void to_bedgraph(const std::string &path);
  void to_bigwig(const std::string &path);
  void runs(std::vector<RunRow> &out);                      // all rows of this shard, (ref_id, start) order
  void chromosome(const std::string &chrom, std::vector<RunRow> &out);
  void signal(const std::string &chrom, std::vector<float> &out);
  double norm_factor() const;
  void stats(uint64_t out[ST_N_COUNTERS]) const;

  // --- staged operations ---
  void accumulate();            // host file -> pinned -> device, hot path
  void stage();                 // copy the shard's compressed bytes into HBM
  void accumulate_resident();   // hot path over the staged bytes
  void set_total_stats(const uint64_t s[ST_N_COUNTERS]);
  uint64_t finalize();          // scan + collapse; returns the number of rows
  void bedgraph_text(std::string &out);
  // sharded jobs, one output file: text on the device + bytes per reference, then this shard's pieces at the offsets that
  // follow from every rank's table (bytes_all[rank][ref])
  void preallocate_output(const std::string &path);  // one rank, before its pile-up: pages of the merged file allocated meanwhile
  void format_bedgraph(std::vector<uint64_t> &bytes_per_ref);
  uint64_t preallocated_bytes();                      // stops that allocation; bytes of the file that have pages
  void write_bedgraph_pieces(const std::string &path, const uint64_t *bytes_all, int world, int rank, uint64_t preallocated = 0);
  const Timings &timings() const { return tm_; }
  const Shape &shape() const { return shape_; }

  const BamHeader &header() const { return hdr_; }
  const Geometry &geometry() const { return geo_; }
  const bnd_pileup_cfg &config() const { return cfg_; }
  const std::string &bam_path() const { return bam_path_; }
  // BamStats::is_paired_end (bam_utils.rs:564-582): any of the first ~1000 records carries the 0x1 flag
  bool is_paired_end();
  // MultiBamPileup (coverage_analysis.rs:783-848): raw per-bin counts of every BAM, collapsed where all samples agree
  // and written as TSV text formatted on the device
  static void multi_to_tsv(std::vector<Pileup *> &bams, const std::string &path);
  // one BAM per rank: raw dense counts stay on the device, every rank marks "my count changes here" with one bit per
  // bin, the masks of all ranks are OR-ed (gathered over NVLink by the caller), every rank compacts its own column at the shared
  // heads, receives its slice of rows of every column and formats + writes that slice of the one TSV
  // (a rank may hold several BAMs: the group functions take all of them, the shared table lives in the first handle)
  uint64_t multi_prepare(bool resident = false);
  static void multi_change_bits(Pileup *const *ps, int n, uint32_t *dev_bits);
  static void multi_or_bits(int device, uint32_t *dst, const uint32_t *src, uint32_t n_src, uint64_t n_words);
  static uint64_t multi_compact_columns(Pileup *const *ps, int n, const uint32_t *dev_bits, const uint32_t **dev_vals);
  void multi_format_tsv(const uint32_t *dev_cols, uint64_t stride, int n_cols, const int32_t *col_order, uint64_t row_lo, uint64_t row_hi,
                        std::vector<uint64_t> &bytes_per_ref, bool sizes_only = false);
  void multi_write_tsv_pieces(const std::string &path, const char *header, const uint64_t *bytes_all, int world, int rank);
  // bin_counts::count_bins (bin_counts.rs:100-289): samples x bins matrix of raw counts on the grid shared by all BAMs
  struct BinCounts {
    uint64_t bin_size = 0, n_bins = 0;
    std::vector<std::string> chroms;        // grid chromosomes, sorted by name
    std::vector<uint64_t> chrom_len;        // their lengths
    std::vector<uint64_t> chrom_first_row;  // [chroms + 1] first matrix row of each chromosome
    std::vector<uint64_t> counts;           // [n_bins][n_samples]
    std::vector<uint64_t> library_sizes;    // reads kept by each sample's filter (per read and chunk, like the pile-up)
  };
  static void count_bins(const bnd_pileup_cfg *cfgs, const bnd_filter_cfg *filters, int n_samples, BinCounts &out);
  // split / modify / split-exogenous (bam_splitter.rs:54-120, bam_modifier.rs:55-230, spike_in_analysis.rs:233-400): one pass over
  // the file, kept records gathered per output on the device, BGZF-compressed there and written.  mode = BAMW_*; paths: one
  // output (split, modify) or four (endogenous, exogenous, both, unmapped).  counters[BW_N_COUNTERS], file_bytes[4].
  void write_bam(int mode, int tn5_shift, const std::string &exogenous_prefix, const std::vector<std::string> &paths, const std::string &command_line,
                 uint64_t *counters, uint64_t *file_bytes);

 private:
  struct Range {  // what part of the file / genome a pass covers
    uint64_t chunk_lo = 0, chunk_hi = 0;
    uint64_t byte_begin = 0, byte_end = 0;  // file offsets of BGZF block starts
    uint64_t entry_uoff = 0;                // offset of the first record inside the first block
    bool bounded = false;                   // stop once records pass (end_ref, end_pos)
    int end_ref = 0;
    uint64_t end_pos = 0;
  };
  struct Impl;
  struct TextJob;
  std::unique_ptr<Impl> im_;
  void format_text(TextJob &job);
  uint64_t place_pieces(TextJob &job, const uint64_t *bytes_all, int world, int rank, uint64_t base = 0) const;
  void write_pieces(TextJob &job, int fd, const std::string &path, uint64_t total, uint64_t prealloc, bool truncate_at_end, char *premap = nullptr,
                    uint64_t premap_len = 0, bool sparse_ok = false);

  void open();
  Range range_for_chunks(uint64_t chunk_lo, uint64_t chunk_hi) const;
  void run_pass(const Range &r, bool resident);
  void reset_accumulators(const Range &r);
  uint64_t finalize_range(bool dense_counts, bool emit_runs = true);
  void fetch_rows(std::vector<RunRow> &out);
  void ensure_finalized();
  uint64_t estimate_text_bytes(bool whole_job) const;
  uint64_t shard_bins_upper() const { return geo_.first_bin_of_chunk(shard_.chunk_hi) - geo_.first_bin_of_chunk(shard_.chunk_lo); }

  bnd_pileup_cfg cfg_;
  std::string bam_path_;
  FilterHost filt_;
  BamHeader hdr_;
  Bai bai_;
  Geometry geo_;
  uint64_t file_size_ = 0;
  Range shard_;
  bool accumulated_ = false, finalized_ = false, total_stats_set_ = false;
  uint64_t stats_[ST_N_COUNTERS] = {0};
  Timings tm_;
  Shape shape_;
};

// K1 exposed on its own (bnd_bgzf_inflate): inflate BGZF blocks held in host memory on the device.
void device_inflate(const uint8_t *bgzf, uint64_t n_bytes, std::vector<uint8_t> &out, std::vector<int32_t> &status, int device);

// the BGZF compressor of the BAM writers on its own (bnd_bgzf_compress)
void device_bgzf_compress(const uint8_t *data, uint64_t n, std::vector<uint8_t> &out, int device);

int device_count();

// collapse-bedgraph (src/cli.rs:1149-1205 + bedgraph_utils.rs:5-45): the raw bytes of the bedGraph are parsed and collapsed on
// the device.  Outputs are the collapsed groups in (chrom, start) order; out.chrom indexes `names` (byte-string order).
struct BdgRows {
  std::vector<uint32_t> chrom;
  std::vector<long long> start, end;
  std::vector<double> score;
};
void collapse_bedgraph_text(const uint8_t *text, uint64_t n, std::vector<std::string> &names, BdgRows &out, int device);

// bigwig-compare / bigwig-aggregate (bamnado/src/bigwig_compare.rs:616-790) on the device.  comparison: 0 subtraction, 1 ratio,
// 2 log-ratio; mode: 10 sum, 11 mean, 12 max, 13 min, 14 median (kernel_args.h BWT_*).  Null pointers = the reference's defaults.
void bigwig_compare(const char *bw1, const char *bw2, const char *out, int comparison, uint32_t bin_size, const double *pseudocount, const double *scale1,
                    const double *scale2, int device);
void bigwig_aggregate(const char *const *paths, int n, const char *out, int mode, uint32_t bin_size, const double *pseudocount, const double *scale_factors,
                      int device);

// the three BAM-writing tools behind one call (a Pileup is opened on `bam_path` for the pass)
void bam_write_tool(const char *bam_path, const bnd_filter_cfg *filter, int mode, int tn5_shift, const char *exogenous_prefix,
                    const std::vector<std::string> &paths, const char *command_line, int device, uint64_t *counters, uint64_t *file_bytes);
// Path::with_extension as BamSplitter::new uses it (spike_in_analysis.rs:171-174): "<prefix minus its extension>.<ext>"
std::string path_with_extension(const std::string &prefix, const std::string &ext);

void bigwig_from_intervals(const char *const *chrom_names, const uint32_t *chrom_lens, uint32_t n_chrom, const uint32_t *chrom_idx, const uint32_t *start,
                           const uint32_t *end, const float *value, uint64_t n, const char *out_path, int device);

}  // namespace bnd
