// cuda_emul.cpp — scheduler and runtime shim of the test-only SIMT emulator (see cuda_emul.h).
#include "cuda_emul.h"

#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

namespace emul {

thread_local Lane *cur = nullptr;
thread_local Cta *cta = nullptr;

namespace {
int g_workers = 0;
constexpr size_t STACK_BYTES = 256 * 1024;

extern "C" void emul_ctx_switch(void **from_sp, void *to_sp);
asm(R"(
.text
.globl emul_ctx_switch
.type emul_ctx_switch,@function
emul_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emul_ctx_switch,.-emul_ctx_switch
)");

struct Worker {
  void *sched_sp = nullptr;
  const std::function<void()> *body = nullptr;
  std::vector<Lane> lanes;
  unsigned char *stacks = nullptr;
  size_t n_stacks = 0;
  ~Worker() {
    if (stacks) munmap(stacks, n_stacks * STACK_BYTES);
  }
};
thread_local Worker *tl_worker = nullptr;

void lane_exit();

void trampoline() {
  Worker *w = tl_worker;
  (*w->body)();
  lane_exit();
}

void lane_exit() {
  Lane *l = cur;
  l->done = true;
  Cta *c = cta;
  c->alive--;
  // a CTA barrier the remaining lanes are waiting on may now be complete
  if (c->alive > 0 && c->bar.count >= c->alive) {
    c->bar.count = 0;
    c->bar.gen++;
  }
  void *dummy;
  emul_ctx_switch(&dummy, tl_worker->sched_sp);
  abort();  // never resumed
}

void prepare_lane(Worker &w, int i) {
  unsigned char *top = w.stacks + (size_t)(i + 1) * STACK_BYTES;
  uintptr_t t = (uintptr_t)top & ~(uintptr_t)15;
  void **sp = (void **)t;
  *--sp = nullptr;                    // fake return address of the trampoline (keeps rsp % 16 == 8 at its entry)
  *--sp = (void *)&trampoline;        // `ret` target
  for (int k = 0; k < 6; k++) *--sp = nullptr;  // rbp rbx r12 r13 r14 r15
  w.lanes[i].sp = (void *)sp;
  w.lanes[i].done = false;
}

void run_cta(Worker &w, Cta &c, int nthreads) {
  cta = &c;
  c.alive = nthreads;
  c.bar = Bar();
  for (int s = 0; s < 3; s++) {
    c.red[s][0] = 0;
    c.red[s][1] = 1;
    c.red[s][2] = 0;
  }
  int nwarps = (nthreads + 31) / 32;
  for (int i = 0; i < nwarps; i++) c.warps[i] = Warp();
  for (int i = 0; i < nthreads; i++) prepare_lane(w, i);
  int remaining = nthreads;
  while (remaining > 0) {
    for (int i = 0; i < nthreads; i++) {
      Lane &l = w.lanes[i];
      if (l.done) continue;
      cur = &l;
      emul_ctx_switch(&w.sched_sp, l.sp);
      if (l.done) remaining--;
    }
  }
  cur = nullptr;
  cta = nullptr;
}
}  // namespace

void yield() {
  Lane *l = cur;
  emul_ctx_switch(&l->sp, tl_worker->sched_sp);
}

void set_workers(int n) { g_workers = n; }

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
  const int nthreads = (int)(block.x * block.y * block.z);
  const uint64_t nctas = (uint64_t)grid.x * grid.y * grid.z;
  if (nthreads <= 0 || nctas == 0) return;
  int nw = g_workers > 0 ? g_workers : (int)std::thread::hardware_concurrency();
  if (const char *e = getenv("BND_EMUL_WORKERS")) nw = atoi(e);
  if (nw < 1) nw = 1;
  if ((uint64_t)nw > nctas) nw = (int)nctas;
  std::atomic<uint64_t> next{0};
  auto work = [&]() {
    Worker w;
    tl_worker = &w;
    w.body = &body;
    w.lanes.resize(nthreads);
    w.n_stacks = nthreads;
    w.stacks = (unsigned char *)mmap(nullptr, w.n_stacks * STACK_BYTES, PROT_READ | PROT_WRITE,
                                     MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (w.stacks == MAP_FAILED) {
      perror("emul: mmap stacks");
      abort();
    }
    for (int i = 0; i < nthreads; i++) {
      unsigned x = i % block.x, y = (i / block.x) % block.y, z = i / (block.x * block.y);
      w.lanes[i].tid = uint3{x, y, z};
      w.lanes[i].lane = i & 31;
      w.lanes[i].warp = i >> 5;
    }
    std::vector<Warp> warps((nthreads + 31) / 32);
    std::vector<unsigned char> smem(smem_bytes + 64);
    Cta c;
    c.warps = warps.data();
    c.dyn_smem = (unsigned char *)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
    c.bdim = block;
    c.gdim = grid;
    for (;;) {
      uint64_t b = next.fetch_add(1);
      if (b >= nctas) break;
      c.bid = uint3{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((uint64_t)grid.x * grid.y))};
      run_cta(w, c, nthreads);
    }
    tl_worker = nullptr;
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nw; i++) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
}

}  // namespace emul

// ---- runtime shim: host memory stands in for device memory, everything is synchronous ------------------------------
struct emul_stream {
  int dummy;
};
struct emul_event {
  std::chrono::steady_clock::time_point t;
};
cudaError_t cudaMalloc(void **p, size_t n) {
  *p = nullptr;
  if (posix_memalign(p, 256, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
  memset(*p, 0xA5, n);  // device memory is not zeroed; poison it so missing memsets show up
  return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
  free(p);
  return cudaSuccess;
}
cudaError_t cudaMallocHost(void **p, size_t n) { return posix_memalign(p, 4096, n ? n : 1) == 0 ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMallocHost(p, n); }
cudaError_t cudaFreeHost(void *p) {
  free(p);
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
  memmove(d, s, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) {
  memmove(d, s, n);
  return cudaSuccess;
}
cudaError_t cudaMemset(void *d, int v, size_t n) {
  memset(d, v, n);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) {
  memset(d, v, n);
  return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t *s) {
  *s = new emul_stream();
  return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  delete s;
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) {
  *e = new emul_event();
  return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) {
  delete e;
  return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
  e->t = std::chrono::steady_clock::now();
  return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) {
  *d = 0;
  return cudaSuccess;
}
cudaError_t cudaGetDeviceCount(int *n) {
  *n = 1;
  return cudaSuccess;
}
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
  memset(p, 0, sizeof *p);
  strcpy(p->name, "SIMT emulator (tests only)");
  p->multiProcessorCount = 4;
  p->major = 10;
  p->minor = 0;
  p->totalGlobalMem = (size_t)8 << 30;
  p->sharedMemPerBlockOptin = 227 * 1024;
  return cudaSuccess;
}
cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b) {
  *free_b = *total_b = (size_t)8 << 30;
  return cudaSuccess;
}
