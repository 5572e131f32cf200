// kernels.h — host-callable launchers of the sm_100a kernels (defined in kernels.cu).
#pragma once
#include "cuda_rt.h"
#include "kernel_args.h"

namespace bnd {

// K1 — per-BGZF-block DEFLATE inflate + CRC32 (inflate.cuh)
void make_crc_tables(CrcTables *t);  // host: tables matching the lane group width the kernel is built with
int inflate_ctas_per_sm();
cudaError_t launch_inflate(const InflateArgs &a, int sm_count, cudaStream_t s);

// K2 — record index (records.cuh): guess+walk, fix+scan, compact
cudaError_t launch_inflate_zlib(const InflateArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_record_index(const RecArgs &a, cudaStream_t s);
cudaError_t launch_carry_copy(const uint8_t *src_buf, uint8_t *dst_buf, const SegState *prev, SegState *cur,
                              uint64_t carry_cap, cudaStream_t s);

// K3+K4 — filter + interval + scatter (pileup.cuh); persistent grid-stride over the segment's records
cudaError_t launch_pileup(const PileArgs &a, int sm_count, cudaStream_t s);

// K5 — scan / collapse / compaction and its consumers (finalize.cuh)
cudaError_t launch_bins_reduce(const FinArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_tile_scan(const FinArgs &a, cudaStream_t s);
cudaError_t launch_bins_emit(const FinArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_expand_runs(const RunArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_dense_signal(const SignalArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_text_len(const TextArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_text_scan(const TextArgs &a, cudaStream_t s);
cudaError_t launch_text_write(const TextArgs &a, int sm_count, cudaStream_t s);

// K6 — collapse-bedgraph and multi-BAM collapse (collapse.cuh)
cudaError_t launch_bdg_sorted_check(const BdgArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bdg_count(const BdgArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bdg_emit(const BdgArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_u64_scan(uint64_t *v, uint32_t n, unsigned long long *total_out, cudaStream_t s);
// collapse-bedgraph reader (bedgraph_parse.cuh)
cudaError_t launch_bdg_newline_count(const BdgParseArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bdg_line_starts(const BdgParseArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bdg_parse_lines(const BdgParseArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bdg_scatter_slow(const BdgSlow *slow, const double *vals, uint64_t n, double *score, int sm_count, cudaStream_t s);
cudaError_t launch_bdg_row_count(const BdgRowArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bdg_compact_rows(const BdgRowArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bdg_name_heads(const BdgNameArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bdg_name_segments(const BdgNameArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bdg_assign_chrom(const BdgNameArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bdg_gather_rows(const uint32_t *perm, uint64_t n, const uint32_t *chrom, const long long *start, const long long *end, const double *score,
                                   uint32_t *o_chrom, long long *o_start, long long *o_end, double *o_score, int sm_count, cudaStream_t s);
cudaError_t launch_multi_bits(const MultiArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_multi_or(uint32_t *dst, const uint32_t *src, uint32_t n_src, uint64_t n_words, int sm_count, cudaStream_t s);
cudaError_t launch_multi_text_len(const MultiTextArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_multi_text_write(const MultiTextArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bin_matrix(const BinMatrixArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_multi_count(const MultiArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_multi_emit(const MultiArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_multi_rows(const RunArgs &a, int sm_count, cudaStream_t s);
// BigWig on the device (bigwig.cuh): items, sections + zlib streams, zoom summaries
cudaError_t launch_bw_row_first(const uint64_t *run_bin, uint64_t n_runs, const GeomDev &g, uint64_t *row_first, cudaStream_t s);
cudaError_t launch_bw_items(const BwItemArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bw_deflate(const BwDeflateArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bw_compact(const BwCompactArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bw_zoom_windows(const BwZoomArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bw_zoom_compact(const BwZoomArgs &a, unsigned long long total, int sm_count, cudaStream_t s);
// stable sort permutation of rows by (chrom rank, start): CUB radix sort (library) — only used when the input is unsorted
cudaError_t sort_rows_by_chrom_start(const uint32_t *d_chrom, const long long *d_start, uint64_t n, uint32_t *d_perm, cudaStream_t s);

// number of kernel launches issued through this file since process start (bench's gpu_launches)
unsigned long long launch_count();

// bigwig-compare / bigwig-aggregate (bigwig_tools.cuh)
cudaError_t launch_bwr_headers(const uint8_t *data, const uint64_t *sec_off, uint32_t n_secs, uint32_t *headers, int sm_count, cudaStream_t s);
cudaError_t launch_bwr_emit_items(const BwrEmitArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bwr_check_sorted(const BwItem *items, const uint64_t *range, uint32_t n_chrom, unsigned int *bad, int sm_count, cudaStream_t s);
cudaError_t launch_bwt_bins(const BwtArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bwt_median_count(const BwtArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bwt_median_fill(const BwtArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bwt_heads(const BwtArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bwt_emit(const BwtArgs &a, int sm_count, cudaStream_t s);

// BAM writers (bam_write.cuh): classify + scan of one segment; gather into the class streams; BGZF compression of a stream
cudaError_t launch_bamw_segment(const BamwArgs &a, int sm_count, cudaStream_t s);
cudaError_t launch_bamw_gather(const BamwArgs &a, int sm_count, cudaStream_t s);
unsigned bamw_deflate_grid(uint32_t n_blocks, int sm_count);  // CTAs the next call launches (token scratch is per CTA)
cudaError_t launch_bamw_deflate(const BamwDeflateArgs &a, int sm_count, cudaStream_t s);

}  // namespace bnd
