/*
 * bamnado_oracle.h — C interface of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference
 * (alsmith151/BamNado v0.9.0) coverage hot path, used as the checker for the
 * CUDA product in `tests/`, `__graft_entry__.smoke()` and as the
 * `cpu_baseline` / `--impl reference` leg of `bench.py`.  Nothing under
 * `bamnado_b200/` may include, link or call it.
 *
 * Parity pinning: the oracle reproduces the reference's own known-answer
 * tests (bamnado/src/coverage_analysis.rs:989-1094: max 117.0 raw,
 * 10134.2568359375 CPM/RPKM on test/data/test.bam at bin 1000;
 * bamnado/src/bedgraph_utils.rs:82-129; bamnado/src/signal_normalization.rs:81-103)
 * — see tests/test_oracle_golden.py.  The reference itself (Rust) cannot be
 * built in this image (no cargo/rustc), so there is no oracle/_ref.
 */
#ifndef BAMNADO_ORACLE_H
#define BAMNADO_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* counters in the order of BamReadFilterStats (bamnado/src/read_filter.rs:31-58) */
enum {
  ORC_N_TOTAL = 0, ORC_FAIL_PROPER_PAIR, ORC_FAIL_MAPQ, ORC_FAIL_LENGTH, ORC_FAIL_BLACKLIST,
  ORC_FAIL_BARCODE, ORC_FAIL_READ_GROUP, ORC_FAIL_STRAND, ORC_FAIL_TAG, ORC_FAIL_FRAGMENT_LENGTH,
  ORC_FAIL_DUPLICATE, ORC_FAIL_SECONDARY, ORC_FAIL_SUPPLEMENTARY, ORC_N_COUNTERS
};

/* BamReadFilter::new argument list (bamnado/src/read_filter.rs:438-492) as plain data */
typedef struct {
  int32_t strand;            /* 0 both, 1 forward, 2 reverse */
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
  const char *blacklist_bed;     /* NULL = none; BED path (>=4 columns) */
  int32_t blacklist_merge_overlaps; /* 1 on the single-BAM CLI path (src/cli.rs:672-678) */
  const char *barcode_csv;       /* NULL = none; CSV with header column `barcode` */
  int32_t barcode_csv_column;    /* -1: column named `barcode`; >=0: that column (CellBarcodesMulti) */
  int32_t has_barcode_list;      /* explicit list (Python ReadFilter.whitelisted_barcodes) */
  uint64_t n_barcodes;
  const char *const *barcodes;
} orc_filter_cfg;

/* BamPileup::new argument list (bamnado/src/coverage_analysis.rs:298-324) */
typedef struct {
  const char *bam_path;
  uint64_t bin_size;
  int32_t norm_method;       /* 0 raw, 1 rpkm, 2 cpm */
  float scale_factor;
  int32_t use_fragment;
  int32_t collapse;
  int32_t ignore_scaffold;
  int32_t has_shift;  int32_t shift[4];
  int32_t has_truncate; int32_t trunc_has5; uint64_t trunc5; int32_t trunc_has3; uint64_t trunc3;
  int32_t n_threads;         /* worker threads over genomic chunks; <=0: all cores */
  int32_t faithful_io;       /* 1: re-open BAM + re-read .bai per chunk as the reference does */
} orc_pileup_cfg;

typedef struct {             /* one output row (Iv + contig) */
  int32_t ref_id;
  uint32_t val;
  uint64_t start, stop;
} orc_run;

const char *orc_last_error(void);
void orc_free(void *p);

/* pileup_genome: all runs over all contigs, sorted by (ref_id, start). */
int orc_pileup_runs(const orc_pileup_cfg *cfg, const orc_filter_cfg *f, orc_run **runs, uint64_t *n_runs,
                    uint64_t stats[ORC_N_COUNTERS], double *norm_factor);
/* pileup_chromosome (coverage_analysis.rs:337-430) */
int orc_pileup_chromosome(const orc_pileup_cfg *cfg, const orc_filter_cfg *f, const char *chrom, orc_run **runs,
                          uint64_t *n_runs, uint64_t stats[ORC_N_COUNTERS]);
/* get_signal_for_chromosome (bamnado-python/src/lib.rs:248-304): dense f32[chrom_size]; returns 2 for unknown chrom */
int orc_signal_for_chromosome(const orc_pileup_cfg *cfg, const orc_filter_cfg *f, const char *chrom, float **out,
                              uint64_t *n);
/* to_bedgraph (coverage_analysis.rs:618-632) */
int orc_pileup_to_bedgraph(const orc_pileup_cfg *cfg, const orc_filter_cfg *f, const char *out_path,
                           uint64_t stats[ORC_N_COUNTERS], double *norm_factor);
/* bench helper: to_bedgraph restricted to a few chromosomes (bounded sample of a large BAM) */
int orc_pileup_chroms_to_bedgraph(const orc_pileup_cfg *cfg, const orc_filter_cfg *f, const char *const *chroms, int32_t n_chroms,
                                  const char *out_path, uint64_t stats[ORC_N_COUNTERS], double *norm_factor);
/* MultiBamPileup::to_tsv (coverage_analysis.rs:783-864), intended semantics */
int orc_multi_to_tsv(const orc_pileup_cfg *cfgs, const orc_filter_cfg *filters, int32_t n_bams, const char *out_path);
/* collapse-bedgraph subcommand (src/cli.rs:1149-1205 + bedgraph_utils.rs:5-45) */
int orc_collapse_bedgraph(const char *in_path, const char *out_path);
/* NormalizationMethod::scale_factor (signal_normalization.rs:50-74) */
double orc_scale_factor(int32_t method, float base_scale, uint64_t bin_size, uint64_t n_reads);
/* BamStats (bam_utils.rs:311-436) */
int orc_bam_stats(const char *bam_path, uint64_t bin_size, uint64_t *n_mapped, uint64_t *n_unmapped,
                  uint64_t *genome_len, uint64_t *chunk_len, uint64_t *n_chunks, int32_t *n_refs);
/* ryu-style shortest f64 text used for the score column */
int orc_format_f64(double v, char *buf, size_t cap);
/* collapse_equal_bins on explicit rows (coverage_analysis.rs:81-134): rows sorted by (chrom,start) then grouped;
   scores is row-major n_rows x n_cols u32; output written in place, returns new row count */
uint64_t orc_collapse_equal_bins(int32_t *chrom, uint64_t *start, uint64_t *end, uint32_t *scores, uint64_t n_rows,
                                 uint32_t n_cols);
/* count records by a full sequential scan (bench bookkeeping) */
int orc_count_records(const char *bam_path, uint64_t *n_records, uint64_t *inflated_bytes, uint64_t *n_blocks);

/* BAM-writing tools (bam_writers.inl): split (bam_splitter.rs:54-120), modify (bam_modifier.rs:55-230), split-exogenous
   (spike_in_analysis.rs:233-400; paths[4] = endogenous, exogenous, both, unmapped; stats[10] in the order of SplitStats) */
int orc_bam_split(const char *bam, const orc_filter_cfg *fc, const char *out_path, const char *command_line, uint64_t *n_written);
int orc_bam_modify(const char *bam, const orc_filter_cfg *fc, int32_t tn5_shift, const char *out_path, const char *command_line, uint64_t *n_written);
int orc_bam_split_exogenous(const char *bam, const orc_filter_cfg *fc, const char *exogenous_prefix, const char *const *paths, const char *command_line,
                            uint64_t stats[10]);

#ifdef __cplusplus
}
#endif
#endif
