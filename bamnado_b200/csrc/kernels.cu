// kernels.cu — the single device translation unit: all sm_100a kernels of the coverage path and their launchers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (see bamnado_b200/build.py).
#include "kernels.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <numeric>
#include <vector>
#if !defined(BND_EMUL)
#include <cub/device/device_radix_sort.cuh>
#endif

#include "bam_write.cuh"
#include "bedgraph_parse.cuh"
#include "bigwig.cuh"
#include "bigwig_tools.cuh"
#include "collapse.cuh"
#include "finalize.cuh"
#include "inflate_spec.cuh"
#include "pileup.cuh"
#include "records.cuh"

namespace bnd {

namespace {
std::atomic<unsigned long long> g_launches{0};
inline cudaError_t launched() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}
inline unsigned grid_for(uint64_t items, uint64_t per_cta, uint64_t cap) {
  uint64_t n = (items + per_cta - 1) / per_cta;
  if (n < 1) n = 1;
  if (n > cap) n = cap;
  return (unsigned)n;
}
}  // namespace

unsigned long long launch_count() { return g_launches.load(); }

namespace {
// K1 shapes: CTA size / resident CTAs per SM / root table bits / span bits / match-list length.  One shape ships; the
// others exist for the tests (a short match list forces rounds that cannot take all 32 lanes) and for A/B timing.
struct InflateShape {
  int threads, ctas_per_sm;
};
const InflateShape kShapes[] = {{320, 3}, {320, 3}, {256, 4}, {32, 30}, {320, 3}};
int shape_index() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("BND_INFLATE_VARIANT");
    v = e ? atoi(e) : 0;
    if (v < 0 || v >= (int)(sizeof(kShapes) / sizeof(kShapes[0]))) v = 0;
  }
  return v;
}
template <int T, int MINB, int LB, int DB, int SPAN, int NT, bool CRC_SMEM, bool ZLIB = false>
cudaError_t launch_inflate_spec_t(const InflateArgs &a, int sm_count, int ctas_per_sm, cudaStream_t s) {
  auto kern = bgzf_inflate_spec_kernel<T, MINB, LB, DB, SPAN, NT, CRC_SMEM, ZLIB>;
  constexpr int warps = T / 32;
  const size_t smem = inflate_spec_smem_bytes<LB, DB, NT, SPAN>(warps, CRC_SMEM);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  unsigned grid = grid_for(a.n_blocks, warps, (uint64_t)sm_count * ctas_per_sm);
  BND_LAUNCH(kern, dim3(grid), dim3(T), smem, s, a);
  return launched();
}
}  // namespace

int inflate_ctas_per_sm() { return kShapes[shape_index()].ctas_per_sm; }

void make_crc_tables(CrcTables *T) {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    T->t[i] = c;
  }
  // z[b][x]: the CRC register value (x << 8b) advanced over 32 words of zero bits (one word per lane of the warp)
  for (int b = 0; b < 4; b++)
    for (uint32_t x = 0; x < 256; x++) {
      uint32_t c = x << (8 * b);
      for (int k = 0; k < 4 * 32; k++) c = T->t[c & 0xff] ^ (c >> 8);
      T->z[b][x] = c;
    }
}

cudaError_t launch_inflate(const InflateArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_blocks == 0) return cudaSuccess;
  const InflateShape &v = kShapes[shape_index()];
  switch (shape_index()) {
    case 1: return launch_inflate_spec_t<320, 3, 9, 8, 256, 384, true>(a, sm_count, v.ctas_per_sm, s);  // round-1 shape: 256-bit spans, ten warps per CTA
    case 2: return launch_inflate_spec_t<256, 4, 9, 8, 256, 256, true>(a, sm_count, v.ctas_per_sm, s);  // short match list (tests)
    case 3: return launch_inflate_spec_t<32, 30, 9, 8, 320, 384, false>(a, sm_count, v.ctas_per_sm, s);  // 30 one-warp CTAs / SM, CRC tables through L1 (A/B: 8 % slower)
    case 4: return launch_inflate_spec_t<320, 3, 9, 8, 320, 384, true>(a, sm_count, v.ctas_per_sm, s);  // 320-bit spans, list of 384 (shipped until the 6-byte list made room: 2.9 % slower)
    default: return launch_inflate_spec_t<320, 3, 9, 8, 384, 448, true>(a, sm_count, v.ctas_per_sm, s);  // 30 warps / SM in three CTAs, 384-bit spans, 448 matches per round
  }
}

// zlib streams (BigWig sections): the same walk, Adler-32 instead of CRC-32, sizes reported per stream
cudaError_t launch_inflate_zlib(const InflateArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_blocks == 0) return cudaSuccess;
  return launch_inflate_spec_t<320, 3, 9, 8, 384, 448, true, true>(a, sm_count, 3, s);
}

cudaError_t launch_record_index(const RecArgs &a, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  constexpr int T = 256;
  auto k1 = rec_guess_walk_kernel<T>;
  auto k2 = rec_fix_scan_kernel<256>;
  auto k3 = rec_compact_kernel<T>;
  unsigned grid = (unsigned)(((uint64_t)a.n_tiles * 32 + T - 1) / T);
  BND_LAUNCH(k1, dim3(grid), dim3(T), 0, s, a);
  cudaError_t e = launched();
  if (e != cudaSuccess) return e;
  BND_LAUNCH(k2, dim3(1), dim3(256), 0, s, a);
  e = launched();
  if (e != cudaSuccess) return e;
  BND_LAUNCH(k3, dim3(grid), dim3(T), 0, s, a);
  return launched();
}

cudaError_t launch_carry_copy(const uint8_t *src_buf, uint8_t *dst_buf, const SegState *prev, SegState *cur,
                              uint64_t carry_cap, cudaStream_t s) {
  BND_LAUNCH(carry_copy_kernel, dim3(8), dim3(256), 0, s, src_buf, dst_buf, prev, cur, carry_cap);
  return launched();
}

cudaError_t launch_pileup(const PileArgs &a, int sm_count, cudaStream_t s) {
  BND_LAUNCH(pileup_kernel, dim3((unsigned)(sm_count * 16)), dim3(PILE_THREADS), 0, s, a);
  return launched();
}

cudaError_t launch_bins_reduce(const FinArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bins_reduce_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(FIN_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_tile_scan(const FinArgs &a, cudaStream_t s) {
  auto k = tile_scan_kernel<1024>;
  BND_LAUNCH(k, dim3(1), dim3(1024), 0, s, a);
  return launched();
}
cudaError_t launch_bins_emit(const FinArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bins_emit_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(FIN_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_expand_runs(const RunArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_runs == 0) return cudaSuccess;
  BND_LAUNCH(expand_runs_kernel, dim3(grid_for(a.n_runs, 256, (uint64_t)sm_count * 16)), dim3(256), 0, s, a);
  return launched();
}
cudaError_t launch_dense_signal(const SignalArgs &a, int sm_count, cudaStream_t s) {
  if (a.chrom_size == 0) return cudaSuccess;
  BND_LAUNCH(dense_signal_kernel, dim3(grid_for(a.chrom_size, 256, (uint64_t)sm_count * 16)), dim3(256), 0, s, a);
  return launched();
}
cudaError_t launch_text_len(const TextArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(text_len_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(TEXT_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_text_scan(const TextArgs &a, cudaStream_t s) {
  auto k = text_scan_kernel<1024>;
  BND_LAUNCH(k, dim3(1), dim3(1024), 0, s, a);
  return launched();
}
cudaError_t launch_text_write(const TextArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(text_write_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(TEXT_THREADS), 0, s, a);
  return launched();
}

cudaError_t launch_bdg_sorted_check(const BdgArgs &a, int sm_count, cudaStream_t s) {
  if (a.n < 2) return cudaSuccess;
  BND_LAUNCH(bdg_sorted_check_kernel, dim3(grid_for(a.n, COL_THREADS, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_bdg_count(const BdgArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bdg_count_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_bdg_emit(const BdgArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bdg_emit_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_u64_scan(uint64_t *v, uint32_t n, unsigned long long *total_out, cudaStream_t s) {
  auto k = u64_scan_kernel<1024>;
  BND_LAUNCH(k, dim3(1), dim3(1024), 0, s, v, n, total_out);
  return launched();
}
cudaError_t launch_bin_matrix(const BinMatrixArgs &a, int sm_count, cudaStream_t s) {
  const uint64_t n = a.g.bin_hi - a.g.bin_lo;
  if (n == 0) return cudaSuccess;
  BND_LAUNCH(bin_matrix_kernel, dim3(grid_for(n, COL_THREADS, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_multi_bits(const MultiArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_bins == 0) return cudaSuccess;
  BND_LAUNCH(multi_bits_kernel, dim3(grid_for(a.n_bins, COL_THREADS, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_multi_or(uint32_t *dst, const uint32_t *src, uint32_t n_src, uint64_t n_words, int sm_count, cudaStream_t s) {
  if (n_words == 0) return cudaSuccess;
  BND_LAUNCH(multi_or_kernel, dim3(grid_for(n_words, COL_THREADS, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, dst, src, n_src, n_words);
  return launched();
}
cudaError_t launch_multi_text_len(const MultiTextArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(multi_text_len_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(TEXT_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_multi_text_write(const MultiTextArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(multi_text_write_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(TEXT_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_multi_count(const MultiArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(multi_count_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_multi_emit(const MultiArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(multi_emit_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_multi_rows(const RunArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_runs == 0) return cudaSuccess;
  BND_LAUNCH(multi_rows_kernel, dim3(grid_for(a.n_runs, 256, (uint64_t)sm_count * 16)), dim3(256), 0, s, a);
  return launched();
}

// ---- collapse-bedgraph reader (bedgraph_parse.cuh) ---------------------------------------------------------------------------
cudaError_t launch_bdg_newline_count(const BdgParseArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bdg_newline_count_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(BDG_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_bdg_line_starts(const BdgParseArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bdg_line_starts_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(BDG_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_bdg_parse_lines(const BdgParseArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_lines == 0) return cudaSuccess;
  BND_LAUNCH(bdg_parse_lines_kernel, dim3(grid_for(a.n_lines, 256, (uint64_t)sm_count * 16)), dim3(256), 0, s, a);
  return launched();
}
cudaError_t launch_bdg_scatter_slow(const BdgSlow *slow, const double *vals, uint64_t n, double *score, int sm_count, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  BND_LAUNCH(bdg_scatter_slow_kernel, dim3(grid_for(n, 256, (uint64_t)sm_count * 8)), dim3(256), 0, s, slow, vals, n, score);
  return launched();
}
cudaError_t launch_bdg_row_count(const BdgRowArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bdg_row_count_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_bdg_compact_rows(const BdgRowArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bdg_compact_rows_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_bdg_name_heads(const BdgNameArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bdg_name_heads_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_bdg_name_segments(const BdgNameArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bdg_name_segments_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_bdg_assign_chrom(const BdgNameArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bdg_assign_chrom_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_bdg_gather_rows(const uint32_t *perm, uint64_t n, const uint32_t *chrom, const long long *start, const long long *end, const double *score,
                                   uint32_t *o_chrom, long long *o_start, long long *o_end, double *o_score, int sm_count, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  BND_LAUNCH(bdg_gather_rows_kernel, dim3(grid_for(n, 256, (uint64_t)sm_count * 8)), dim3(256), 0, s, perm, n, chrom, start, end, score, o_chrom, o_start, o_end, o_score);
  return launched();
}

// ---- BigWig (bigwig.cuh) ------------------------------------------------------------------------------------------------
cudaError_t launch_bw_row_first(const uint64_t *run_bin, uint64_t n_runs, const GeomDev &g, uint64_t *row_first, cudaStream_t s) {
  BND_LAUNCH(bw_row_first_kernel, dim3((unsigned)((g.n_ref + 1 + 127) / 128)), dim3(128), 0, s, run_bin, n_runs, g, row_first);
  return launched();
}
cudaError_t launch_bw_items(const BwItemArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_runs == 0) return cudaSuccess;
  BND_LAUNCH(bw_items_kernel, dim3(grid_for(a.n_runs, 256, (uint64_t)sm_count * 16)), dim3(256), 0, s, a);
  return launched();
}
size_t bw_deflate_smem_bytes(uint32_t header_bytes, uint32_t unit_bytes) {
  const uint32_t raw = header_bytes + unit_bytes * BW_ITEMS;
  return ((raw + 3u) & ~3u) + ((bw_comp_bound(raw) + 3) / 4 + 1) * 4;
}
cudaError_t launch_bw_deflate(const BwDeflateArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_sections == 0) return cudaSuccess;
  const size_t smem = bw_deflate_smem_bytes(a.header_bytes, a.unit_bytes);
  static size_t attr_done = 0;
  if (smem > attr_done) {
    cudaError_t e = cudaFuncSetAttribute(bw_deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_done = smem;
  }
  BND_LAUNCH(bw_deflate_kernel, dim3(grid_for(a.n_sections, 1, (uint64_t)sm_count * 8)), dim3(BW_THREADS), smem, s, a);
  return launched();
}
cudaError_t launch_bw_compact(const BwCompactArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_sections == 0) return cudaSuccess;
  BND_LAUNCH(bw_compact_kernel, dim3(grid_for(a.n_sections, 8, (uint64_t)sm_count * 8)), dim3(256), 0, s, a);
  return launched();
}
cudaError_t launch_bw_zoom_windows(const BwZoomArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bw_zoom_windows_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(BW_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_bw_zoom_compact(const BwZoomArgs &a, unsigned long long total, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bw_zoom_compact_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(BW_THREADS), 0, s, a);
  cudaError_t e = launched();
  if (e != cudaSuccess) return e;
  BND_LAUNCH(bw_zoom_chrom_first_kernel, dim3((a.n_chrom + 1 + 127) / 128), dim3(128), 0, s, a, total);
  return launched();
}

// ---- bigwig-compare / bigwig-aggregate (bigwig_tools.cuh) --------------------------------------------------------------------
cudaError_t launch_bwr_headers(const uint8_t *data, const uint64_t *sec_off, uint32_t n_secs, uint32_t *headers, int sm_count, cudaStream_t s) {
  if (n_secs == 0) return cudaSuccess;
  BND_LAUNCH(bwr_headers_kernel, dim3(grid_for(n_secs, 256, (uint64_t)sm_count * 8)), dim3(256), 0, s, data, sec_off, n_secs, headers);
  return launched();
}
cudaError_t launch_bwr_emit_items(const BwrEmitArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_secs == 0) return cudaSuccess;
  BND_LAUNCH(bwr_emit_items_kernel, dim3(grid_for(a.n_secs, 8, (uint64_t)sm_count * 8)), dim3(256), 0, s, a);
  return launched();
}
cudaError_t launch_bwr_check_sorted(const BwItem *items, const uint64_t *range, uint32_t n_chrom, unsigned int *bad, int sm_count, cudaStream_t s) {
  if (n_chrom == 0) return cudaSuccess;
  BND_LAUNCH(bwr_check_sorted_kernel, dim3((unsigned)sm_count, std::min<uint32_t>(n_chrom, 1024u)), dim3(256), 0, s, items, range, n_chrom, bad);
  return launched();
}
cudaError_t launch_bwt_bins(const BwtArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_bins == 0) return cudaSuccess;
  BND_LAUNCH(bwt_bins_kernel, dim3(grid_for(a.n_bins, 256, (uint64_t)sm_count * 16)), dim3(256), 0, s, a);
  return launched();
}
cudaError_t launch_bwt_median_count(const BwtArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_bins == 0) return cudaSuccess;
  BND_LAUNCH(bwt_median_count_kernel, dim3(grid_for(a.n_bins, 256, (uint64_t)sm_count * 16)), dim3(256), 0, s, a);
  return launched();
}
cudaError_t launch_bwt_median_fill(const BwtArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_bins == 0) return cudaSuccess;
  BND_LAUNCH(bwt_median_fill_kernel, dim3(grid_for(a.n_bins, 256, (uint64_t)sm_count * 16)), dim3(256), 0, s, a);
  return launched();
}
cudaError_t launch_bwt_heads(const BwtArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bwt_heads_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_bwt_emit(const BwtArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_tiles == 0) return cudaSuccess;
  BND_LAUNCH(bwt_emit_kernel, dim3(grid_for(a.n_tiles, 1, (uint64_t)sm_count * 8)), dim3(COL_THREADS), 0, s, a);
  return launched();
}

// ---- BAM writers (bam_write.cuh) ---------------------------------------------------------------------------------------------
cudaError_t launch_bamw_segment(const BamwArgs &a, int sm_count, cudaStream_t s) {
  BND_LAUNCH(bamw_classify_kernel, dim3((unsigned)sm_count * 8), dim3(BAMW_TILE), 0, s, a);
  if (cudaError_t e = launched()) return e;
  BND_LAUNCH(bamw_scan_kernel, dim3(1), dim3(BAMW_THREADS), 0, s, a);
  return launched();
}
cudaError_t launch_bamw_gather(const BamwArgs &a, int sm_count, cudaStream_t s) {
  BND_LAUNCH(bamw_gather_kernel, dim3((unsigned)sm_count * 8), dim3(BAMW_TILE), 0, s, a);
  return launched();
}
unsigned bamw_deflate_grid(uint32_t n_blocks, int sm_count) { return grid_for(n_blocks, 1, (uint64_t)sm_count * 2); }
cudaError_t launch_bamw_deflate(const BamwDeflateArgs &a, int sm_count, cudaStream_t s) {
  if (a.n_blocks == 0) return cudaSuccess;
  const size_t smem = BGZF_RAW + 16;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(bamw_deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  BND_LAUNCH(bamw_deflate_kernel, dim3(bamw_deflate_grid(a.n_blocks, sm_count)), dim3(BAMW_THREADS), smem, s, a);
  return launched();
}

cudaError_t sort_rows_by_chrom_start(const uint32_t *d_chrom, const long long *d_start, uint64_t n, uint32_t *d_perm, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
#if defined(BND_EMUL)
  (void)s;  // tests/emul: device memory is host memory
  std::iota(d_perm, d_perm + n, 0u);
  std::stable_sort(d_perm, d_perm + n, [&](uint32_t x, uint32_t y) {
    return d_chrom[x] != d_chrom[y] ? d_chrom[x] < d_chrom[y] : d_start[x] < d_start[y];
  });
  return cudaSuccess;
#else
  // LSD: stable sort by start (sign bit flipped so that i64 order == u64 order), then stable sort by chrom rank
  if (n > 0x7fffffffull) return cudaErrorInvalidValue;
  const int N = (int)n;
  uint64_t *k0 = nullptr, *k1 = nullptr;
  uint32_t *v0 = nullptr, *c0 = nullptr, *c1 = nullptr;
  void *tmp = nullptr;
  cudaError_t e = cudaSuccess;
  auto done = [&](cudaError_t err) {
    cudaFree(k0), cudaFree(k1), cudaFree(v0), cudaFree(c0), cudaFree(c1), cudaFree(tmp);
    return err;
  };
  if ((e = cudaMalloc((void **)&k0, n * 8)) || (e = cudaMalloc((void **)&k1, n * 8)) || (e = cudaMalloc((void **)&v0, n * 4)) ||
      (e = cudaMalloc((void **)&c0, n * 4)) || (e = cudaMalloc((void **)&c1, n * 4)))
    return done(e);
  std::vector<long long> h_start(n);
  std::vector<uint32_t> iota(n), h_chrom(n);
  std::iota(iota.begin(), iota.end(), 0u);
  if ((e = cudaMemcpyAsync(h_start.data(), d_start, n * 8, cudaMemcpyDeviceToHost, s)) ||
      (e = cudaMemcpyAsync(h_chrom.data(), d_chrom, n * 4, cudaMemcpyDeviceToHost, s)) || (e = cudaStreamSynchronize(s)))
    return done(e);
  std::vector<uint64_t> keys(n);
  for (uint64_t i = 0; i < n; i++) keys[i] = (uint64_t)h_start[i] ^ 0x8000000000000000ull;
  if ((e = cudaMemcpyAsync(k0, keys.data(), n * 8, cudaMemcpyHostToDevice, s)) || (e = cudaMemcpyAsync(v0, iota.data(), n * 4, cudaMemcpyHostToDevice, s)))
    return done(e);
  size_t tb = 0, tb2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, d_perm, N, 0, 64, s);
  cub::DeviceRadixSort::SortPairs(nullptr, tb2, c0, c1, d_perm, v0, N, 0, 32, s);
  tb = std::max(tb, tb2);
  if ((e = cudaMalloc(&tmp, tb))) return done(e);
  if ((e = cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, v0, d_perm, N, 0, 64, s))) return done(e);
  std::vector<uint32_t> perm(n), gathered(n);
  if ((e = cudaMemcpyAsync(perm.data(), d_perm, n * 4, cudaMemcpyDeviceToHost, s)) || (e = cudaStreamSynchronize(s))) return done(e);
  for (uint64_t i = 0; i < n; i++) gathered[i] = h_chrom[perm[i]];
  if ((e = cudaMemcpyAsync(c0, gathered.data(), n * 4, cudaMemcpyHostToDevice, s))) return done(e);
  if ((e = cub::DeviceRadixSort::SortPairs(tmp, tb, c0, c1, d_perm, v0, N, 0, 32, s))) return done(e);
  if ((e = cudaMemcpyAsync(d_perm, v0, n * 4, cudaMemcpyDeviceToDevice, s)) || (e = cudaStreamSynchronize(s))) return done(e);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return done(cudaSuccess);
#endif
}

}  // namespace bnd
