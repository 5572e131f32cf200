// common.cuh — shared device helpers for the sm_100a kernels of the BamNado coverage path.
//
// Everything here is byte/integer plumbing (unaligned little-endian loads, sub-warp tiles, launch macro).  The
// path has no GEMM-shaped work, so nothing goes through tensor cores; the rules that matter are coalescing,
// shared-memory staging and keeping enough independent warps in flight.
#pragma once
#include <stdint.h>
#if defined(BND_EMUL)
#include "cuda_emul.h"  // tests/emul: CPU SIMT emulator, test infrastructure only
#define BND_LAUNCH(kern, grid, block, smem, stream, ...) \
  emul::launch((grid), (block), (smem), [=]() { kern(__VA_ARGS__); })
#define BND_DYN_SMEM(name) unsigned char *name = emul::dyn_smem()
#else
#include <cuda_runtime.h>
#define BND_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define BND_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace bnd {

template <class T>
__host__ __device__ __forceinline__ T tmin(T a, T b) {
  return a < b ? a : b;
}
template <class T>
__host__ __device__ __forceinline__ T tmax(T a, T b) {
  return a > b ? a : b;
}

// A power-of-two group of G consecutive lanes inside one warp, driven with raw warp intrinsics.
template <int G>
struct Tile {
  static_assert(G >= 1 && G <= 32 && (G & (G - 1)) == 0, "tile width must be a power of two <= 32");
  unsigned mask;
  int lane;   // rank inside the tile
  int shift;  // first warp lane of the tile
  __device__ __forceinline__ Tile() {
    int wl = (int)(threadIdx.x & 31u);
    lane = wl & (G - 1);
    shift = wl - lane;
    mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << shift);
  }
  template <class T>
  __device__ __forceinline__ T shfl(T v, int src) const {
    return __shfl_sync(mask, v, src, G);
  }
  template <class T>
  __device__ __forceinline__ T shfl_up(T v, unsigned d) const {
    return __shfl_up_sync(mask, v, d, G);
  }
  template <class T>
  __device__ __forceinline__ T shfl_down(T v, unsigned d) const {
    return __shfl_down_sync(mask, v, d, G);
  }
  template <class T>
  __device__ __forceinline__ T shfl_xor(T v, int d) const {
    return __shfl_xor_sync(mask, v, d, G);
  }
  __device__ __forceinline__ unsigned ballot(int pred) const { return (__ballot_sync(mask, pred) & mask) >> shift; }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};

// ---- shared-memory window addresses --------------------------------------------------------------------------------
// Hot loops address shared memory through 32-bit window addresses and ld/st.shared PTX: the compiler otherwise keeps
// re-deriving 64-bit generic pointers (S2R + IMAD chains in the SASS of the inflate walk, profiles/).
#if defined(BND_EMUL)
typedef uintptr_t saddr_t;
__device__ __forceinline__ saddr_t smem_addr(const void *p) { return (saddr_t)p; }
__device__ __forceinline__ uint32_t lds_u16(saddr_t a) { return *(const uint16_t *)a; }
__device__ __forceinline__ uint32_t lds_u32(saddr_t a) { return *(const uint32_t *)a; }
__device__ __forceinline__ void sts_u32(saddr_t a, uint32_t v) { *(uint32_t *)a = v; }
__device__ __forceinline__ void sts_u16(saddr_t a, uint32_t v) { *(uint16_t *)a = (uint16_t)v; }
__device__ __forceinline__ void cp_async_u32(saddr_t dst, const void *src) { *(uint32_t *)dst = *(const uint32_t *)src; }
__device__ __forceinline__ void cp_async_wait_all() {}
__device__ __forceinline__ void mbar_init(saddr_t, uint32_t) {}
__device__ __forceinline__ void bulk_copy_g2s(saddr_t dst, const void *src, uint32_t bytes, saddr_t) { memcpy((void *)dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(saddr_t, uint32_t) {}
__device__ __forceinline__ void bulk_store_fence() {}
__device__ __forceinline__ void bulk_store_s2g(void *dst, saddr_t src, uint32_t bytes) { memcpy(dst, (const void *)src, bytes); }
#else
typedef uint32_t saddr_t;
__device__ __forceinline__ saddr_t smem_addr(const void *p) { return (saddr_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u16(saddr_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(saddr_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u32(saddr_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(saddr_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }
// asynchronous 4-byte global -> shared copy (LDGSTS): no register staging, completion awaited with cp_async_wait_all
__device__ __forceinline__ void cp_async_u32(saddr_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// ---- bulk asynchronous copy (TMA engine: UBLKCP in the SASS) completed through an mbarrier ----------------------------------
// One elected thread hands a whole tile (16-byte aligned source and destination, size a multiple of 16) to the copy engine; the
// consumers wait on the barrier's phase.  Used where a CTA stages tens of KB at once (the BGZF compressor's 64 KB block); K1's
// 1.3 KB rounds are better off with per-lane cp.async (profiles/r2_sweeps.md).
__device__ __forceinline__ void mbar_init(saddr_t mbar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Caller: every thread that read or wrote the destination through ordinary loads / stores has passed a barrier before this.
__device__ __forceinline__ void bulk_copy_g2s(saddr_t dst, const void *src, uint32_t bytes, saddr_t mbar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses to dst are ordered before the engine's writes
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(saddr_t mbar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  } while (!done);
}
// shared -> global by the copy engine (UBLKCP.G.S): every thread that wrote the tile calls bulk_store_fence() before the barrier
// that precedes the store; the issuing thread returns once the engine has READ the tile (shared memory reusable), the global
// writes complete on their own
__device__ __forceinline__ void bulk_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_s2g(void *dst, saddr_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
#endif

// ---- unaligned little-endian loads (BAM records sit at arbitrary byte offsets) ---------------------------------
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t *p) {
  uintptr_t a = (uintptr_t)p;
  const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
  unsigned sh = (unsigned)(a & 3) * 8;
  uint32_t lo = w[0];
  if (sh == 0) return lo;
  uint32_t hi = w[1];
  return __funnelshift_r(lo, hi, sh);
}
__device__ __forceinline__ uint32_t ld_u16_unaligned(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// warp-wide inclusive scan (all 32 lanes participate)
template <class T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
  const int lane = (int)(threadIdx.x & 31u);
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += o;
  }
  return v;
}
template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// CTA-wide exclusive scan for THREADS threads; `total` receives the CTA sum.  scratch: THREADS/32 + 1 entries.
template <int THREADS, class T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *scratch, T &total) {
  constexpr int NW = THREADS / 32;
  const int lane = (int)(threadIdx.x & 31u), warp = (int)(threadIdx.x >> 5);
  T inc = warp_inclusive_scan(v);
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    T w = lane < NW ? scratch[lane] : T(0);
    T wi = warp_inclusive_scan(w);
    if (lane < NW) scratch[lane] = wi - w;
    if (lane == NW - 1) scratch[NW] = wi;
  }
  __syncthreads();
  T out = scratch[warp] + inc - v;
  total = scratch[NW];
  __syncthreads();
  return out;
}

}  // namespace bnd
