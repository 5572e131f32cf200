// cuda_emul.h — a small SIMT emulator used ONLY by the test-suite (tests/emul).
//
// DEVELOPMENT AID, NOT A PRODUCT PATH.  The build box has no GPU, so kernel logic bugs would otherwise cost a
// GPU-box round trip each.  This header lets the *same* kernel sources (bamnado_b200/csrc/*.cuh) be compiled by g++
// with -DBND_EMUL and executed on the CPU: every CUDA thread of a CTA is a coroutine, warp collectives and
// __syncthreads are coroutine barriers, CTAs are spread over a few OS threads.  It is never linked into
// libbamnado_b200.so and the bamnado_b200 package cannot load it; the product fails loudly without a CUDA device.
// It does not model memory ordering, alignment faults or timing: the GPU tests (-m gpu) remain the parity gate.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __constant__ static const
#define __shared__ static thread_local
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))

struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 {
  unsigned x, y;
};
struct alignas(16) uint4 {
  unsigned x, y, z, w;
};
struct alignas(16) int4 {
  int x, y, z, w;
};
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

namespace emul {

struct Bar {
  int count = 0;
  unsigned gen = 0;
};
struct Warp {
  uint64_t slot[6][2][32];
  Bar bar[32][6];  // keyed by (lowest lane of the mask, log2 of its population): masks of different sizes may share a leader
};
inline int mask_class(unsigned mask) {
  int n = __builtin_popcount(mask), c = 0;
  while ((1 << c) < n) c++;
  return c > 5 ? 5 : c;
}
struct Cta {
  uint3 bid;
  dim3 bdim, gdim;
  Bar bar;
  int alive = 0;
  int red[3][3];  // or / and / count accumulators, slot = barrier generation % 3
  unsigned char *dyn_smem = nullptr;
  Warp *warps = nullptr;
};
struct Lane {
  uint3 tid;
  int lane;  // lane within warp
  int warp;  // warp within CTA
  void *sp = nullptr;
  bool done = false;
};
extern thread_local Lane *cur;
extern thread_local Cta *cta;

void yield();
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body);
void set_workers(int n);

inline void bar_wait(Bar &b, int expected) {
  unsigned g = b.gen;
  if (++b.count >= expected) {
    b.count = 0;
    b.gen++;
  } else {
    while (b.gen == g) yield();
  }
}
inline Warp &my_warp() { return cta->warps[cur->warp]; }
inline unsigned char *dyn_smem() { return cta->dyn_smem; }

template <class T>
inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle operand too wide");
  uint64_t b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(uint64_t b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}
// exchange: every participating lane publishes `v`, waits for the others, returns the parity buffer to read from
template <class T>
inline const uint64_t *exchange(unsigned mask, T v) {
  Warp &w = my_warp();
  int leader = __builtin_ffs((int)mask) - 1, cls = mask_class(mask);
  Bar &b = w.bar[leader][cls];
  int par = (int)(b.gen & 1);
  w.slot[cls][par][cur->lane] = to_bits(v);
  bar_wait(b, __builtin_popcount(mask));
  return w.slot[cls][par];
}
}  // namespace emul

#define threadIdx (emul::cur->tid)
#define blockIdx (emul::cta->bid)
#define blockDim (emul::cta->bdim)
#define gridDim (emul::cta->gdim)
#define warpSize 32

// ---- warp collectives ------------------------------------------------------------------------------------------
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const uint64_t *s = emul::exchange(mask, v);
  int lane = emul::cur->lane;
  int from = (lane & ~(width - 1)) | (src & (width - 1));
  return emul::from_bits<T>(s[from]);
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const uint64_t *s = emul::exchange(mask, v);
  int lane = emul::cur->lane;
  int rel = lane & (width - 1);
  return rel < (int)delta ? v : emul::from_bits<T>(s[lane - (int)delta]);
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const uint64_t *s = emul::exchange(mask, v);
  int lane = emul::cur->lane;
  int rel = lane & (width - 1);
  return rel + (int)delta >= width ? v : emul::from_bits<T>(s[lane + (int)delta]);
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  const uint64_t *s = emul::exchange(mask, v);
  int lane = emul::cur->lane;
  int from = lane ^ lanemask;
  if ((from & ~(width - 1)) != (lane & ~(width - 1))) return v;
  return emul::from_bits<T>(s[from]);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  const uint64_t *s = emul::exchange(mask, (unsigned)(pred != 0));
  unsigned out = 0;
  for (int i = 0; i < 32; i++)
    if ((mask >> i) & 1u)
      if (s[i]) out |= 1u << i;
  return out;
}
template <class T>
inline unsigned __match_any_sync(unsigned mask, T v) {
  const uint64_t *s = emul::exchange(mask, v);
  const uint64_t mine = emul::to_bits(v);
  unsigned out = 0;
  for (int i = 0; i < 32; i++)
    if (((mask >> i) & 1u) && s[i] == mine) out |= 1u << i;
  return out;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
  emul::Warp &w = emul::my_warp();
  emul::bar_wait(w.bar[__builtin_ffs((int)mask) - 1][emul::mask_class(mask)], __builtin_popcount(mask));
}

// ---- CTA barriers ----------------------------------------------------------------------------------------------
inline int emul_sync_reduce(int pred, int which) {
  // slot g%3 serves the barrier of generation g; after passing it every lane re-arms slot (g+2)%3, which no lane
  // can be using yet (to reach generation g+2 every lane must first have passed g+1, hence finished this call)
  emul::Cta *c = emul::cta;
  int slot = (int)(c->bar.gen % 3);
  if (which == 0) c->red[slot][0] |= (pred != 0);
  if (which == 1) c->red[slot][1] &= (pred != 0);
  if (which == 2) c->red[slot][2] += (pred != 0);
  emul::bar_wait(c->bar, c->alive);
  int out = which >= 0 ? c->red[slot][which] : 0;
  int re = (slot + 2) % 3;
  c->red[re][0] = 0;
  c->red[re][1] = 1;
  c->red[re][2] = 0;
  return out;
}
inline void __syncthreads() { emul_sync_reduce(0, -1); }
inline int __syncthreads_or(int pred) { return emul_sync_reduce(pred, 0); }
inline int __syncthreads_and(int pred) { return emul_sync_reduce(pred, 1); }
inline int __syncthreads_count(int pred) { return emul_sync_reduce(pred, 2); }
inline void __threadfence() {}
inline void __threadfence_block() {}

// ---- integer intrinsics ----------------------------------------------------------------------------------------
inline float __uint_as_float(unsigned u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline unsigned __float_as_uint(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
inline double __ddiv_rn(double a, double b) {
  volatile double r = a / b;
  return r;
}
inline double __dmul_rn(double a, double b) {
  volatile double r = a * b;
  return r;
}
inline double __dadd_rn(double a, double b) {
  volatile double r = a + b;
  return r;
}
inline double __dsub_rn(double a, double b) {
  volatile double r = a - b;
  return r;
}
// round-to-nearest single operations (the emulator build uses -ffp-contract=off semantics: separate statements)
inline float __fmul_rn(float a, float b) {
  volatile float r = a * b;
  return r;
}
inline float __fadd_rn(float a, float b) {
  volatile float r = a + b;
  return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline unsigned __brev(unsigned v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  return __builtin_bswap32(v);
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (unsigned)(v >> (shift & 31));
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
  return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
template <class T>
inline T __ldg(const T *p) {
  return *p;
}

// ---- atomics (global and shared alike) ---------------------------------------------------------------------------
template <class T>
inline T atomicAdd(T *p, T v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <class T>
inline T atomicSub(T *p, T v) {
  return __atomic_fetch_sub(p, v, __ATOMIC_RELAXED);
}
template <class T>
inline T atomicOr(T *p, T v) {
  return __atomic_fetch_or(p, v, __ATOMIC_RELAXED);
}
template <class T>
inline T atomicAnd(T *p, T v) {
  return __atomic_fetch_and(p, v, __ATOMIC_RELAXED);
}
template <class T>
inline T atomicExch(T *p, T v) {
  return __atomic_exchange_n(p, v, __ATOMIC_RELAXED);
}
template <class T>
inline T atomicCAS(T *p, T cmp, T v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
  return cmp;
}
template <class T>
inline T atomicMax(T *p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
  return old;
}
template <class T>
inline T atomicMin(T *p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
  return old;
}

// ---- the sliver of the CUDA runtime API the engine uses (all synchronous) ----------------------------------------
typedef int cudaError_t;
typedef struct emul_stream *cudaStream_t;
typedef struct emul_event *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNoDevice = 100 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0, cudaHostAllocDefault = 0, cudaHostAllocPortable = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
  char name[256];
  int multiProcessorCount;
  int major, minor;
  size_t totalGlobalMem;
  size_t sharedMemPerBlockOptin;
};
cudaError_t cudaMalloc(void **p, size_t n);
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t n);
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned flags);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st = 0);
cudaError_t cudaMemset(void *d, int v, size_t n);
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st = 0);
cudaError_t cudaStreamCreate(cudaStream_t *s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventQuery(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b);
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) {
  return cudaSuccess;
}
