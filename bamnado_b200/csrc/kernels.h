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

// number of kernel launches issued through this file since process start (bench's gpu_launches)
unsigned long long launch_count();

}  // namespace bnd
