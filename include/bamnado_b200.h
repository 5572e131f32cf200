/*
 * bamnado_b200.h — C ABI of the B200-native BamNado coverage path.
 *
 * This is the drop-in boundary: the thin `extern "C"` surface a Rust `-sys` crate / pyo3
 * module / CLI binds instead of calling the reference's Rust engine directly.  Plain pointers and sizes only;
 * no torch, no C++ types.  Every entry point cites the reference interface it replaces (paths are relative to
 * the reference repository, alsmith151/BamNado v0.9.0).
 *
 * All functions returning `int` return 0 on success, non-zero on failure; `bnd_last_error()` then holds a
 * thread-local message and `bnd_last_error_class()` the class, so a binding can raise the reference's
 * exception types (RuntimeError / KeyError / ValueError; bamnado-python/src/lib.rs:22-24,109-116,265-270).
 *
 * There is no CPU fallback: the library needs a CUDA device (sm_100a build) for every compute call.
 */
#ifndef BAMNADO_B200_H
#define BAMNADO_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BND_ABI_VERSION 4

/* error classes */
enum { BND_OK = 0, BND_ERR_RUNTIME = 1, BND_ERR_NOT_FOUND = 2, BND_ERR_INVALID = 3 };

/* Filter counters, in the order of BamReadFilterStats (bamnado/src/read_filter.rs:31-58). */
enum {
  BND_N_TOTAL = 0, BND_FAIL_PROPER_PAIR, BND_FAIL_MAPQ, BND_FAIL_LENGTH, BND_FAIL_BLACKLIST, BND_FAIL_BARCODE,
  BND_FAIL_READ_GROUP, BND_FAIL_STRAND, BND_FAIL_TAG, BND_FAIL_FRAGMENT_LENGTH, BND_FAIL_DUPLICATE,
  BND_FAIL_SECONDARY, BND_FAIL_SUPPLEMENTARY, BND_N_COUNTERS
};

/* BamReadFilter::new + with_exclude_* (bamnado/src/read_filter.rs:438-492) as plain data.  File-based fields
 * mirror what the CLI / Python layer loads before constructing the filter (src/cli.rs:648-720,
 * bamnado-python/src/lib.rs:167-219). */
typedef struct bnd_filter_cfg {
  int32_t strand;            /* 0 both, 1 forward, 2 reverse (bam_utils.rs:95-133) */
  int32_t proper_pair;
  uint32_t min_mapq;
  uint32_t min_length;
  uint32_t max_length;
  int32_t has_min_fragment_length;
  uint32_t min_fragment_length;
  int32_t has_max_fragment_length;
  uint32_t max_fragment_length;
  int32_t exclude_duplicates, exclude_secondary, exclude_supplementary;
  const char *read_group;        /* NULL = none */
  const char *filter_tag;        /* NULL = none */
  const char *filter_tag_value;  /* NULL = none */
  const char *blacklist_bed;     /* NULL = none; BED path, >= 4 columns (bam_utils.rs:650-684) */
  int32_t blacklist_merge_overlaps; /* 1 on the single-BAM CLI path (src/cli.rs:672-678) */
  const char *barcode_csv;       /* NULL = none; CSV with a header (bam_utils.rs:186-269) */
  int32_t barcode_csv_column;    /* -1: column named `barcode`; >= 0: that column index (CellBarcodesMulti) */
  int32_t has_barcode_list;      /* explicit list (Python ReadFilter.whitelisted_barcodes) */
  uint64_t n_barcodes;
  const char *const *barcodes;
} bnd_filter_cfg;

/* BamPileup::new (bamnado/src/coverage_analysis.rs:298-324) as plain data, plus device/shard placement. */
typedef struct bnd_pileup_cfg {
  const char *bam_path;
  uint64_t bin_size;
  int32_t norm_method;       /* 0 raw, 1 rpkm, 2 cpm (signal_normalization.rs:13-21) */
  float scale_factor;
  int32_t use_fragment;
  int32_t collapse;
  int32_t ignore_scaffold;
  int32_t has_shift;  int32_t shift[4];   /* genomic_intervals.rs:50-59 */
  int32_t has_truncate; int32_t trunc_has5; uint64_t trunc5; int32_t trunc_has3; uint64_t trunc3;
  int32_t n_threads;         /* BigWig writer threads (coverage_analysis.rs:752-754); <= 0: 6 */
  int32_t device;            /* CUDA device ordinal; < 0: current device */
  int32_t shard_rank;        /* this process's shard of the genomic chunks (0 .. shard_world-1) */
  int32_t shard_world;       /* <= 1: unsharded */
} bnd_pileup_cfg;

typedef struct bnd_run {     /* one output row: rust_lapper::Interval<usize,u32> + contig (bam_utils.rs:32) */
  int32_t ref_id;
  uint32_t val;
  uint64_t start, stop;      /* 1-based half-open, as the reference keeps them */
} bnd_run;

typedef struct bnd_pileup bnd_pileup;  /* opaque; single owner, not thread-safe */

const char *bnd_last_error(void);
int bnd_last_error_class(void);
void bnd_free(void *p);                /* frees buffers returned through out-pointers */
int bnd_abi_version(void);
int bnd_device_count(void);
/* CUDA kernel launches issued by this library since the process started (bench.py's gpu_launches) */
uint64_t bnd_launch_count(void);

/* ---- BamPileup (coverage_analysis.rs:136-760) ------------------------------------------------------------ */
/* BamPileup::new + BamReadFilter::new; opens BAM header + BAI (BamStats::new, bam_utils.rs:311-378). */
bnd_pileup *bnd_pileup_create(const bnd_pileup_cfg *cfg, const bnd_filter_cfg *filter);
void bnd_pileup_destroy(bnd_pileup *p);
/* to_bedgraph (coverage_analysis.rs:618-632) */
int bnd_pileup_to_bedgraph(bnd_pileup *p, const char *out_path);
/* to_bigwig (coverage_analysis.rs:642-759, replaces the bigtools writer at :748-755): items, 1 024-item sections, one zlib
 * stream per section and the zoom summaries are built on the device; the host lays out header, chromosome tree and R-trees */
int bnd_pileup_to_bigwig(bnd_pileup *p, const char *out_path);
/* pileup_chromosome (coverage_analysis.rs:337-430): runs of one contig; unknown contig -> BND_ERR_NOT_FOUND */
int bnd_pileup_chromosome(bnd_pileup *p, const char *chrom, bnd_run **runs, uint64_t *n_runs);
/* pileup_genome / pileup_to_polars rows (coverage_analysis.rs:436-598), sorted by (ref_id, start) */
int bnd_pileup_runs(bnd_pileup *p, bnd_run **runs, uint64_t *n_runs);
/* normalization_factor (coverage_analysis.rs:326-333); valid after a pileup ran on this handle */
int bnd_pileup_norm_factor(bnd_pileup *p, double *factor);
/* filter.stats() snapshot (read_filter.rs:805-808) */
int bnd_pileup_stats(bnd_pileup *p, uint64_t stats[BND_N_COUNTERS]);

/* ---- staged interface used for sharded (multi-GPU) runs and device-resident timing ------------------------ */
/* Stream this shard's BAM bytes host->device and accumulate the per-bin difference array + filter counters. */
int bnd_pileup_accumulate(bnd_pileup *p);
/* Copy this shard's compressed BAM bytes into HBM once (not part of any timed hot path). */
int bnd_pileup_stage(bnd_pileup *p);
/* Run the hot path (inflate, record index, filter, scatter, scan, collapse) over the staged bytes. */
int bnd_pileup_accumulate_resident(bnd_pileup *p);
/* Replace the counters by job-wide totals (after an all-reduce over ranks) before normalising. */
int bnd_pileup_set_total_stats(bnd_pileup *p, const uint64_t stats[BND_N_COUNTERS]);
/* Scan + normalise + collapse on device; returns the number of output rows this shard owns. */
int bnd_pileup_finalize(bnd_pileup *p, uint64_t *n_runs);
/* bedGraph text of this shard's rows, in final (chrom byte order, start) order (device formats the text). */
int bnd_pileup_bedgraph_text(bnd_pileup *p, char **text, uint64_t *n_bytes);
/* per-stage device time of the last accumulate call, ms: [inflate, record index, filter+scatter, finalize, h2d] */
int bnd_pileup_timings(bnd_pileup *p, double ms[8]);
/* shape of the last run: [records visited, compressed bytes, inflated bytes, BGZF blocks, bins, kernel launches,
   chunk length, n chunks, segments (= K1 launches), 7 reserved] */
int bnd_pileup_shape(bnd_pileup *p, uint64_t shape[16]);
/* Sharded jobs write ONE bedGraph (write_bedgraph, coverage_analysis.rs:43-75, sorts all rows by (chrom, start)):
 * 1. every rank formats its rows on the device and gets the text bytes it holds per reference (header order, n_refs entries);
 * 2. the caller gathers those tables from all ranks (all-gather of n_refs x u64 - the only exchange besides the counters);
 * 3. every rank fills its own pieces of the same file: references in byte-string order, inside a reference the shards in rank
 *    order.  bytes_all is [world][n_refs], row `rank` must be what step 1 returned.  No rank waits for another one; the file is
 *    complete once every rank has returned (the caller's barrier).
 * 0. optional, ONE rank, before its pile-up: bnd_pileup_preallocate_output creates the file and allocates its estimated pages in
 *    the background while the pile-ups run (otherwise every rank allocates its own pieces when it writes them, and those
 *    allocations queue up on the file's lock: ~1 us per page on tmpfs, 2.1 GB = 0.5 s).  After its pile-up that rank calls
 *    bnd_pileup_preallocated_bytes (stops the allocation, returns how many leading bytes of the file have pages); the value travels
 *    with the tables of step 2 and every rank passes it to step 3 as `preallocated` (0 = nothing was allocated in advance). */
int bnd_pileup_preallocate_output(bnd_pileup *p, const char *out_path);
int bnd_pileup_preallocated_bytes(bnd_pileup *p, uint64_t *n_bytes);
int bnd_pileup_format_bedgraph(bnd_pileup *p, uint64_t *bytes_per_ref, uint64_t n_refs);
int bnd_pileup_write_bedgraph_pieces(bnd_pileup *p, const char *out_path, const uint64_t *bytes_all, int32_t world, int32_t rank, uint64_t preallocated);

/* ---- Python front door (bamnado-python/src/lib.rs:248-304) ------------------------------------------------ */
/* get_signal_for_chromosome: dense float32[chrom_size]; unknown contig -> BND_ERR_NOT_FOUND */
int bnd_signal_for_chromosome(const bnd_pileup_cfg *cfg, const bnd_filter_cfg *filter, const char *chrom,
                              float **signal, uint64_t *n);

/* ---- MultiBamPileup::to_tsv (coverage_analysis.rs:766-865) ------------------------------------------------ */
int bnd_multi_to_tsv(const bnd_pileup_cfg *cfgs, const bnd_filter_cfg *filters, int32_t n_bams, const char *out_path);

/* ---- bin_counts::count_bins (bamnado/src/bin_counts.rs:100-289; front half of `bam-normalize`, src/cli.rs:1455, and of
 * compute_scale_factors, bamnado-python/src/lib.rs:505) -------------------------------------------------------------------
 * Raw read (or fragment) counts of n_samples BAMs on one shared grid of bins.  bin_size, use_fragment and ignore_scaffold
 * are taken from cfgs[0] (the reference passes one value for all samples); normalisation, collapse, shift and truncate
 * fields are ignored (the reference counts raw, uncollapsed, unshifted).  Every BAM must have the same contig names and
 * lengths (else BND_ERR_RUNTIME with the reference's message).  Outputs (library-allocated, release with bnd_free):
 *   counts            [n_bins][n_samples] row-major, BinCounts::counts
 *   chrom_names       grid chromosomes in row order (sorted by name), '\n'-separated, NUL-terminated
 *   chrom_len         [n_chroms] contig lengths
 *   chrom_first_row   [n_chroms + 1] first matrix row of each chromosome; bin i of a chromosome covers the 1-based inclusive
 *                     interval [1 + i * bin_size, min((i + 1) * bin_size, len)] (BinRegion, bin_counts.rs:27-34)
 *   library_sizes     caller-provided array of n_samples: BinCounts::library_sizes (reads kept by each sample's filter) */
int bnd_bin_counts(const bnd_pileup_cfg *cfgs, const bnd_filter_cfg *filters, int32_t n_samples, uint64_t **counts,
                   uint64_t *n_bins, char **chrom_names, uint64_t **chrom_len, uint64_t **chrom_first_row, uint64_t *n_chroms,
                   uint64_t *library_sizes);

/* ---- MultiBamPileup with one BAM per GPU (coverage_analysis.rs:783-864 run as one rank per BAM) -----------------
 * The pile-ups run independently; what the ranks exchange is (a) one BIT per bin "my count changes here" (hg38 at bin 1: 387 MB
 * per rank instead of the 3.1 GB of a byte mask), (b) slices of the compacted count columns, (c) n_refs x u64 byte counts.  The
 * collectives run on memory the caller owns (e.g. torch CUDA tensors over NCCL/NVLink); every pointer named dev_* is a DEVICE
 * pointer.
 * A rank may hold several BAMs (8 BAMs on 1, 2 or 4 GPUs): the group calls take all handles of the rank; the shared table
 * (row heads + this rank's columns) is kept in ps[0], which is also the handle the text calls are made on.
 *   1. bnd_multi_prepare          per BAM: pile-up + dense raw counts; returns the number of bins (must agree on all ranks);
 *                                 _resident runs over bytes staged by bnd_pileup_stage (device-resident timing)
 *   2. bnd_multi_change_bits      dev_bits[ceil(n_bins / 32)] (u32 words, bit i = bin i opens a new row for one of the rank's BAMs)
 *   3. (caller) all-gather the masks; bnd_multi_or_bits folds n_src masks laid out back to back into dev_dst (may alias mask 0)
 *   4. bnd_multi_compact_columns  the rank's counts at the shared row heads stay on the device: *dev_vals -> u32[n][n_rows]
 *   5. (caller) all-to-all: rank k receives rows [row_lo, row_hi) = [k n / world, (k + 1) n / world) of every column
 *   6. bnd_multi_format_tsv       TSV text of that slice on the device; BAM c's count of row r is
 *                                 dev_cols[col_order[c] * col_stride + (r - row_lo)] (col_order NULL = identity);
 *                                 returns the text bytes this rank holds per reference (header order).  A rank may cut its slice
 *                                 into sub-slices to bound the text held in HBM: sizes_only = 1 computes the byte table without
 *                                 keeping text; the sub-slices then take part in step 7 as consecutive virtual ranks
 *   7. (caller) all-gather those tables; bnd_multi_write_tsv_pieces fills this rank's pieces of the ONE file (references in
 *      byte-string order like collapse_equal_bins sorts them, slices in rank order); rank 0 also writes `header`
 *      ("chrom\tstart\tend\t<bam>...\n", the same string on every rank).  The file is complete when all ranks have returned. */
int bnd_multi_prepare(bnd_pileup *p, uint64_t *n_bins);
int bnd_multi_prepare_resident(bnd_pileup *p, uint64_t *n_bins);
int bnd_multi_change_bits(bnd_pileup *const *ps, int32_t n, void *dev_bits);
int bnd_multi_or_bits(int32_t device, void *dev_dst, const void *dev_src, uint32_t n_src, uint64_t n_words);
int bnd_multi_compact_columns(bnd_pileup *const *ps, int32_t n, const void *dev_bits, uint64_t *n_rows, const void **dev_vals);
int bnd_multi_format_tsv(bnd_pileup *p, const void *dev_cols, uint64_t col_stride, int32_t n_cols, const int32_t *col_order, uint64_t row_lo,
                         uint64_t row_hi, int32_t sizes_only, uint64_t *bytes_per_ref, uint64_t n_refs);
int bnd_multi_write_tsv_pieces(bnd_pileup *p, const char *out_path, const char *header, const uint64_t *bytes_all, int32_t world, int32_t rank);

/* ---- collapse-bedgraph (src/cli.rs:1149-1205 + bedgraph_utils.rs:5-45) ------------------------------------ */
int bnd_collapse_bedgraph(const char *in_path, const char *out_path);

/* ---- bigwig-compare / bigwig-aggregate (bamnado/src/bigwig_compare.rs:616-790; src/cli.rs:259-297, 299-330, 1081-1147) -------
 * The sections of every input are inflated on the device, every output bin is computed by its own thread replaying the
 * reference's sweep for that bin (f64, same order of additions: subtraction / ratio / sum / mean / max / min / median are
 * bit-identical, log-ratio within 1e-6), equal neighbouring bins merge, and the result is written by the device BigWig encoder.
 * Chromosomes and their order are those of the first input.  NULL pseudocount / scale factors = the reference's defaults
 * (compare: 1e-12 and 1.0; aggregate: 0.0 and 1.0).
 *   comparison: 0 subtraction (bw1 - bw2), 1 ratio (bw1 / (bw2 + pc)), 2 log-ratio (ln((bw1 + pc) / (bw2 + pc)))
 *   method:     0 sum, 1 mean, 2 max, 3 min, 4 median (weighted lower median over the bases of a bin) */
int bnd_bigwig_compare(const char *bw1, const char *bw2, const char *out_path, int32_t comparison, uint32_t bin_size, const double *pseudocount,
                       const double *scale_factor_bw1, const double *scale_factor_bw2);
int bnd_bigwig_aggregate(const char *const *bw_paths, int32_t n, const char *out_path, int32_t method, uint32_t bin_size, const double *pseudocount,
                         const double *scale_factors);
/* A BigWig from explicit intervals (sorted by (chromosome index, start), disjoint): the device encoder on caller-provided rows. */
int bnd_bigwig_from_intervals(const char *const *chrom_names, const uint32_t *chrom_lens, uint32_t n_chrom, const uint32_t *chrom_idx, const uint32_t *start,
                              const uint32_t *end, const float *value, uint64_t n, const char *out_path);

/* ---- split / modify / split-exogenous (the BAM-writing consumers of the read filter) -----------------------
 * One pass over the BAM: records inflated and indexed by the coverage kernels, BamReadFilter::is_valid on the device, the kept
 * records gathered per output file, BGZF-compressed on the device (dynamic-Huffman DEFLATE, record-parallel LZ77) and written.
 * The output header is the input's plus the reference's `@PG ID:bamnado PN:bamnado VN:<version> CL:<command_line>` record
 * (bam_utils.rs:757-780).  The input needs its .bai.  `counters` (may be NULL) receives BND_BW_N_COUNTERS values.
 *   bnd_bam_split            BamFilterer::split_async (bam_splitter.rs:54-120): every placed record the per-contig queries return
 *                            and the filter accepts, contigs in header order (the reference walks a HashMap: any order)
 *   bnd_bam_modify           BamModifier::run (bam_modifier.rs:55-230): the same, and with tn5_shift every properly paired record is
 *                            rewritten (+4 / -5 shift, tlen, mate start, bin), the others are dropped like the reference drops them
 *   bnd_bam_split_exogenous  BamSplitter::split (spike_in_analysis.rs:233-400): <prefix>.endogenous.bam, .exogenous.bam (reference
 *                            name starts with exogenous_prefix), .both.bam (always empty), .unmapped.bam (unmapped, QC-fail,
 *                            duplicate, secondary); counters[0..10) are SplitStats in its field order */
enum {
  BND_BW_N_TOTAL = 0, BND_BW_N_UNMAPPED, BND_BW_N_QCFAIL, BND_BW_N_DUPLICATE, BND_BW_N_SECONDARY, BND_BW_N_FILTERED, BND_BW_N_LOW_MAPQ,
  BND_BW_N_BOTH, BND_BW_N_EXOGENOUS, BND_BW_N_ENDOGENOUS, BND_BW_N_WRITTEN, BND_BW_N_DROPPED_BY_SHIFT, BND_BW_N_ERRORS, BND_BW_N_COUNTERS = 16
};
int bnd_bam_split(const char *bam_path, const bnd_filter_cfg *filter, const char *out_path, const char *command_line, int32_t device, uint64_t *counters);
int bnd_bam_modify(const char *bam_path, const bnd_filter_cfg *filter, int32_t tn5_shift, const char *out_path, const char *command_line, int32_t device,
                   uint64_t *counters);
int bnd_bam_split_exogenous(const char *bam_path, const bnd_filter_cfg *filter, const char *exogenous_prefix, const char *out_prefix, const char *command_line,
                            int32_t device, uint64_t *counters);

/* ---- NormalizationMethod::scale_factor (signal_normalization.rs:50-74) ------------------------------------ */
double bnd_scale_factor(int32_t method, float base_scale, uint64_t bin_size, uint64_t n_reads);

/* reference dictionary of the BAM behind a handle (BamStats::chromsizes_ref_id, bam_utils.rs:495-502): names joined by '\n'
 * and lengths in header order; free both with bnd_free */
int bnd_pileup_refs(bnd_pileup *p, char **names, uint64_t **lens, uint64_t *n_refs);

/* ---- BamStats (bam_utils.rs:311-436): [n_mapped, n_unmapped, genome_len, chunk_len, n_chunks, n_refs] ----- */
int bnd_bam_stats(const char *bam_path, uint64_t bin_size, uint64_t out[6]);

/* ---- run_cli (src/cli.rs:737-757): same subcommands/flags/exit codes for the coverage path ---------------- */
int bnd_cli_main(int argc, const char *const *argv);

/* ---- kernel-level entry points (parity tests and micro-benchmarks) ---------------------------------------- */
/* Inflate `n_blocks` BGZF blocks held in host memory on the device; writes the concatenated output into `out`
 * (capacity out_cap) and per-block status (0 ok) into `status`.  CRC32 and ISIZE are verified on the device. */
int bnd_bgzf_inflate(const uint8_t *bgzf, uint64_t n_bytes, uint8_t *out, uint64_t out_cap, uint64_t *out_len,
                     int32_t *status, uint64_t status_cap, uint64_t *n_blocks);

/* The BGZF compressor of the BAM writers on its own: `data` cut into 65 280-byte blocks, each a gzip member with a dynamic-Huffman
 * DEFLATE stream built on the device, followed by the EOF marker.  *out is released with bnd_free. */
int bnd_bgzf_compress(const uint8_t *data, uint64_t n_bytes, uint8_t **out, uint64_t *out_len);

#ifdef __cplusplus
}
#endif
#endif
