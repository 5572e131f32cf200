// engine.cpp — see engine.h.
#include "engine.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <thread>

#include "bigwig_layout.h"
#include "cuda_rt.h"
#include "kernels.h"

namespace bnd {

namespace {

#define CU(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) fail(std::string("CUDA error: ") + cudaGetErrorString(e_) + " in " #call);     \
  } while (0)

// Host cores this process should plan for: the machine's cores divided by the processes sharing the node (one per GPU
// under torchrun, which exports LOCAL_WORLD_SIZE); BND_HOST_SHARE overrides the divisor.
static unsigned host_cores_share() {
  const unsigned hw = std::max(4u, std::thread::hardware_concurrency());
  uint64_t share = 1;
  if (const char *e = getenv("BND_HOST_SHARE")) share = strtoull(e, nullptr, 10);
  else if (const char *e = getenv("LOCAL_WORLD_SIZE")) share = strtoull(e, nullptr, 10);
  if (share < 1) share = 1;
  return std::max(2u, (unsigned)(hw / share));
}

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Page cache -> pinned staging.  A copy whose destination is only read by the DMA engine afterwards should not pull the
// destination lines into the cache first: with non-temporal stores 16 threads move 85 GB/s out of a mapping of the file on the
// B200 hosts, against 61 GB/s for memcpy and 54 GB/s for pread (scripts/host_copy_probe.cu, profiles/r2_sweeps.md).
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2"))) static void stream_copy_avx2(uint8_t *dst, const uint8_t *src, size_t n) {
  while (n && ((uintptr_t)dst & 31)) *dst++ = *src++, n--;
  const size_t v = n / 128;
  for (size_t i = 0; i < v; i++) {
    const __m256i a = _mm256_loadu_si256((const __m256i *)(src)), b = _mm256_loadu_si256((const __m256i *)(src + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i *)(src + 64)), d = _mm256_loadu_si256((const __m256i *)(src + 96));
    _mm256_stream_si256((__m256i *)(dst), a);
    _mm256_stream_si256((__m256i *)(dst + 32), b);
    _mm256_stream_si256((__m256i *)(dst + 64), c);
    _mm256_stream_si256((__m256i *)(dst + 96), d);
    src += 128, dst += 128;
  }
  _mm_sfence();
  memcpy(dst, src, n - v * 128);
}
static void stream_copy(uint8_t *dst, const uint8_t *src, size_t n) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) stream_copy_avx2(dst, src, n);
  else memcpy(dst, src, n);
}
#else
static void stream_copy(uint8_t *dst, const uint8_t *src, size_t n) { memcpy(dst, src, n); }
#endif

uint64_t env_u64(const char *name, uint64_t dflt) {
  const char *e = getenv(name);
  if (!e || !*e) return dflt;
  return strtoull(e, nullptr, 10);
}

// Device allocations are recycled through a process-wide free list: a coverage job allocates the same few large
// buffers every time (bin array, run arrays, text), and cudaMalloc / cudaFree of hundreds of MB cost milliseconds and
// device-wide synchronisation.  Buffers are only released after the work using them has been synchronised.
struct DevPool {
  struct Entry {
    void *p;
    size_t cap;
    int device;
  };
  std::mutex mu;
  std::vector<Entry> free_list;
  size_t held = 0;
  size_t limit() {
    static size_t v = env_u64("BND_POOL_BYTES", 32ull << 30);
    return v;
  }
  void *take(size_t n, size_t &cap_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    int best = -1;
    for (size_t i = 0; i < free_list.size(); i++) {
      const Entry &e = free_list[i];
      if (e.device == dev && e.cap >= n && e.cap <= 2 * n + (4u << 20) && (best < 0 || e.cap < free_list[best].cap)) best = (int)i;
    }
    if (best < 0) return nullptr;
    Entry e = free_list[best];
    free_list.erase(free_list.begin() + best);
    held -= e.cap;
    cap_out = e.cap;
    return e.p;
  }
  void give(void *p, size_t cap) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> g(mu);
      if (held + cap <= limit()) {
        free_list.push_back(Entry{p, cap, dev});
        held += cap;
        return;
      }
    }
    cudaFree(p);
  }
  void trim() {
    std::lock_guard<std::mutex> g(mu);
    for (auto &e : free_list) cudaFree(e.p);
    free_list.clear();
    held = 0;
  }
};
DevPool g_pool;

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  // grows (never shrinks); the caller guarantees no kernel still uses the old allocation
  void ensure(size_t n) {
    if (n <= cap) return;
    release();
    size_t got = 0;
    if (void *q = g_pool.take(n, got)) {
      p = q;
      cap = got;
      return;
    }
    size_t want = n + n / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {  // give cached buffers back to the driver and retry once
      p = nullptr;
      cudaGetLastError();
      g_pool.trim();
      CU(cudaMalloc(&p, want));
    }
    cap = want;
  }
  void release() {
    if (p) g_pool.give(p, cap);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

constexpr int MAX_SETS = 8;  // segments in flight (BND_SETS, default 6): enough queued K1 work to cover the K2/K3 tails

struct SegSet {
  cudaStream_t stream = nullptr;
  DevBuf comp, coff, uoff, infl, t_entry, t_exit, t_count, t_base, t_slots, t_flag, rec_off;
  SegState *st = nullptr;
  uint32_t *work = nullptr;
  cudaEvent_t ev_idx = nullptr, ev_carry = nullptr;
  bool carry_pending = false;  // ev_carry was recorded for the segment that followed this set's last use
};

struct StageSlot {
  uint8_t *pin = nullptr;
  size_t cap = 0;
  BlockTable tab;
  cudaEvent_t ev_h2d = nullptr;
  bool h2d_pending = false;
  void ensure(size_t n) {
    if (n <= cap) return;
    if (pin) CU(cudaFreeHost(pin));
    pin = nullptr;
    cap = 0;
    CU(cudaHostAlloc((void **)&pin, n, cudaHostAllocDefault));
    cap = n;
  }
};

struct PinnedBounce {  // pinned bounce buffers for device -> host text, one pair per device context
  static constexpr int N = 2;
  uint8_t *buf[N] = {nullptr, nullptr};
  size_t cap = 0;
  cudaEvent_t ev[N] = {nullptr, nullptr};
  void ensure(size_t n) {
    if (n <= cap) return;
    for (int i = 0; i < N; i++) {
      if (buf[i]) CU(cudaFreeHost(buf[i]));
      buf[i] = nullptr;
      CU(cudaHostAlloc((void **)&buf[i], n, cudaHostAllocDefault));
      if (!ev[i]) CU(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
    cap = n;
  }
};

}  // namespace

struct DeviceCtx {
  int device = 0;
  int sm_count = 1;
  CrcTables *d_crc = nullptr;
  SegSet sets[MAX_SETS];
  int n_sets = 6;
  std::vector<StageSlot> slots;
  uint64_t seg_bytes = 0, carry_cap = 0;
  int n_readers = 4;
  PinnedBounce bounce;  // created on this context's device, used under `mu`
  std::mutex mu;  // one pass at a time per device context
};

namespace {

std::mutex g_ctx_mu;
std::map<int, std::unique_ptr<DeviceCtx>> g_ctx;

DeviceCtx &ctx_for(int device) {
  std::lock_guard<std::mutex> g(g_ctx_mu);
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    fail("bamnado_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback");
  if (device < 0) CU(cudaGetDevice(&device));
  if (device >= n) fail("CUDA device ordinal out of range", ERR_INVALID);
  auto it = g_ctx.find(device);
  if (it != g_ctx.end()) {
    CU(cudaSetDevice(device));
    return *it->second;
  }
  CU(cudaSetDevice(device));
  auto c = std::make_unique<DeviceCtx>();
  c->device = device;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 1;
  c->seg_bytes = env_u64("BND_SEGMENT_BYTES", 96ull << 20);
  c->carry_cap = env_u64("BND_CARRY_BYTES", 8ull << 20);
  c->carry_cap = (c->carry_cap + REC_TILE - 1) / REC_TILE * REC_TILE;
  // 12 readers keep up with the B200 at 35 GB/s of BAM on the 16-core hosts (8: the pile-up waits 110-130 ms per 11.5 GB)
  c->n_readers = (int)env_u64("BND_READERS", std::min(12u, std::max(2u, host_cores_share() * 3 / 4)));
  c->n_sets = (int)std::min<uint64_t>(MAX_SETS, std::max<uint64_t>(2, env_u64("BND_SETS", 6)));
  int n_slots = (int)env_u64("BND_STAGES", 4);
  if (n_slots < 2) n_slots = 2;
  c->slots.resize(n_slots);
  for (auto &s : c->slots) CU(cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
  CrcTables host_crc;
  make_crc_tables(&host_crc);
  CU(cudaMalloc((void **)&c->d_crc, sizeof(CrcTables)));
  CU(cudaMemcpy(c->d_crc, &host_crc, sizeof(CrcTables), cudaMemcpyHostToDevice));
  for (auto &s : c->sets) {
    CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CU(cudaMalloc((void **)&s.st, sizeof(SegState)));
    CU(cudaMalloc((void **)&s.work, 256));
    CU(cudaEventCreateWithFlags(&s.ev_idx, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s.ev_carry, cudaEventDisableTiming));
  }
  DeviceCtx &ref = *c;
  g_ctx[device] = std::move(c);
  return ref;
}

struct Segment {
  uint64_t begin, end;  // file offsets (BGZF block starts)
};

// ---- reader pool: file bytes -> pinned staging slots, plus the BGZF block table of each segment --------------------
// A segment is read in pieces by all reader threads at once (a single pread stream out of the page cache runs at only
// ~3 GB/s on the B200 hosts, far below the 55 GB/s of the H2D copy that follows); the thread that finishes a segment's
// last piece walks its BGZF block chain.
struct ReaderPool {
  static constexpr uint64_t PIECE = 8ull << 20;
  struct Job {
    size_t seg;
    int slot;
    uint64_t off, len;  // piece inside the segment
  };
  int fd;
  const uint8_t *map;  // read-only mapping of the whole file, or null (then the pieces are pread)
  DeviceCtx &ctx;
  const std::vector<Segment> &segs;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::deque<Job> jobs;
  std::vector<char> done;          // per segment
  std::vector<uint32_t> remaining; // pieces still being read, per segment
  std::string error;
  bool stop = false;
  std::vector<std::thread> threads;

  const bool drop_ptes = env_u64("BND_DROP_PTES", 0) != 0;
  ReaderPool(int fd_, const uint8_t *map_, DeviceCtx &c, const std::vector<Segment> &s, int n_threads) : fd(fd_), map(map_), ctx(c), segs(s) {
    for (int i = 0; i < n_threads; i++) threads.emplace_back([this] { work(); });
  }
  ~ReaderPool() {
    {
      std::lock_guard<std::mutex> g(mu);
      stop = true;
    }
    cv_job.notify_all();
    for (auto &t : threads) t.join();
  }
  void submit(size_t seg, int slot) {
    const uint64_t len = segs[seg].end - segs[seg].begin;
    {
      std::lock_guard<std::mutex> g(mu);
      if (done.size() <= seg) done.resize(seg + 1, 0), remaining.resize(seg + 1, 0);
      uint32_t n = 0;
      for (uint64_t off = 0; off < len || n == 0; off += PIECE) {
        jobs.push_back(Job{seg, slot, off, std::min<uint64_t>(PIECE, len - off)});
        n++;
        if (len == 0) break;
      }
      remaining[seg] = n;
    }
    cv_job.notify_all();
  }
  void wait(size_t seg) {
    std::unique_lock<std::mutex> g(mu);
    cv_done.wait(g, [&] { return !error.empty() || (done.size() > seg && done[seg]); });
    if (!error.empty()) fail(error);
  }
  void work() {
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> g(mu);
        cv_job.wait(g, [&] { return stop || !jobs.empty(); });
        if (stop && jobs.empty()) return;
        j = jobs.front();
        jobs.pop_front();
      }
      std::string err;
      bool last = false;
      const Segment &s = segs[j.seg];
      StageSlot &slot = ctx.slots[j.slot];
      try {
        uint64_t got = 0;
        if (map) {
          stream_copy(slot.pin + j.off, map + s.begin + j.off, j.len);
          got = j.len;
          // BND_DROP_PTES=1: drop the page-table entries of the pages this piece covers completely (the page cache keeps the
          // data), which made unmapping the file at the end of a job cheap.  Off since the mapping outlives the handle
          // (MapCache) and is torn down in the background: the entries are then worth keeping, the next pass over the same
          // file takes no page faults at all.
          if (drop_ptes) {
            const uintptr_t a0 = ((uintptr_t)(map + s.begin + j.off) + 4095) & ~(uintptr_t)4095, a1 = (uintptr_t)(map + s.begin + j.off + j.len) & ~(uintptr_t)4095;
            if (a1 > a0) madvise((void *)a0, a1 - a0, MADV_DONTNEED);
          }
        }
        while (got < j.len) {
          ssize_t r = pread(fd, slot.pin + j.off + got, j.len - got, (off_t)(s.begin + j.off + got));
          if (r < 0) fail("read error on BAM file");
          if (r == 0) fail("unexpected end of BAM file");
          got += (uint64_t)r;
        }
      } catch (const std::exception &e) {
        err = e.what();
      }
      {
        std::lock_guard<std::mutex> g(mu);
        if (!err.empty() && error.empty()) error = err;
        last = --remaining[j.seg] == 0;
      }
      if (last) {
        try {
          const uint64_t len = s.end - s.begin;
          scan_bgzf_blocks(slot.pin, len, slot.tab);
          if (slot.tab.coff.back() != len) fail("BAM segment does not end on a BGZF block boundary (corrupt file or index)");
        } catch (const std::exception &e) {
          err = e.what();
        }
        {
          std::lock_guard<std::mutex> g(mu);
          if (!err.empty() && error.empty()) error = err;
          done[j.seg] = 1;
        }
        cv_done.notify_all();
      } else if (!err.empty()) {
        cv_done.notify_all();
      }
    }
  }
};

}  // namespace

// bedGraph text of one shard: formatted on the device in header order; `pieces` say where each reference's rows go
namespace {
struct TextPiece {
  uint64_t dev_begin, dev_end;  // byte range inside the device text
  uint64_t out_off;             // where it goes in the output
};
}  // namespace

struct Pileup::TextJob {
  DevBuf d_names, d_name_off, d_skip, d_scores, d_row_len, d_tile_off, d_text, d_ref_off;
  uint64_t nbytes = 0;
  std::vector<std::pair<uint64_t, uint64_t>> span;  // per reference: [begin, end) of its rows inside the device text
  std::vector<TextPiece> pieces;                    // where those spans go in the output, sorted by dev_begin
  ~TextJob() {
    for (DevBuf *b : {&d_names, &d_name_off, &d_skip, &d_scores, &d_row_len, &d_tile_off, &d_text, &d_ref_off}) b->release();
  }
};

// Unmaps finished output mappings off the caller's critical path.  The threads are joined when the next job starts and
// when the library is unloaded, so no thread of ours outlives the code it runs.
namespace {
struct Unmapper {
  struct Job {
    std::thread th;
    std::shared_ptr<std::atomic<bool>> done;
  };
  std::mutex mu;
  std::vector<Job> jobs;
  // A teardown holds the process's mmap lock (≈ 8 ms per GB of touched pages on the B200 hosts).  A job that follows at once sets
  // up its own mappings and threads in its first millisecond, and would do so behind that lock (12-26 ms measured at the start of
  // every bench step but the first).  The teardown therefore waits a moment before it starts, unless somebody is waiting for it.
  std::mutex go_mu;
  std::condition_variable go_cv;
  bool hurry = false;
  void reap() {  // waits for every teardown
    std::vector<Job> t;
    {
      std::lock_guard<std::mutex> g(mu);
      t.swap(jobs);
    }
    if (t.empty()) return;
    {
      std::lock_guard<std::mutex> g(go_mu);
      hurry = true;
    }
    go_cv.notify_all();
    for (auto &x : t)
      if (x.th.joinable()) x.th.join();
    std::lock_guard<std::mutex> g(go_mu);
    hurry = false;
  }
  void submit(char *map, uint64_t len) {
    std::lock_guard<std::mutex> g(mu);
    // threads that have finished are joined here, without waiting for the others, so that none piles up between reaps
    for (size_t i = 0; i < jobs.size();) {
      if (jobs[i].done->load()) {
        jobs[i].th.join();
        jobs.erase(jobs.begin() + (long)i);
      } else {
        i++;
      }
    }
    auto done = std::make_shared<std::atomic<bool>>(false);
    jobs.push_back(Job{std::thread([this, map, len, done] {
                         if (len > (256ull << 20)) {
                           std::unique_lock<std::mutex> lk(go_mu);
                           go_cv.wait_for(lk, std::chrono::milliseconds(60), [this] { return hurry; });
                         }
                         const uint64_t step = 64ull << 20;  // piece by piece: other threads get the lock in between
                         for (uint64_t off = 0; off < len; off += step) munmap(map + off, std::min<uint64_t>(step, len - off));
                         done->store(true);
                       }),
                       done});
  }
  ~Unmapper() { reap(); }
} g_unmapper;

// The read-only mapping of the most recently used BAM is kept for the next handle on the same file (same device, inode, size):
// setting up and tearing down an 11.5 GB mapping costs ~25 ms of the process's mmap lock per handle, which a caller that asks
// one BAM for one chromosome after the other (get_signal_for_chromosome) or runs job after job would pay every time.  A shared
// file mapping always shows the file's current bytes, so only a change of size or identity makes the cached mapping unusable.
struct MapCache {
  std::mutex mu;
  const uint8_t *map = nullptr;
  uint64_t len = 0;
  dev_t dev = 0;
  ino_t ino = 0;
  bool in_use = false;
  const uint8_t *take(const struct stat &st) {
    std::lock_guard<std::mutex> g(mu);
    if (!map || in_use || st.st_dev != dev || st.st_ino != ino || (uint64_t)st.st_size != len) return nullptr;
    in_use = true;
    return map;
  }
  // a handle is done with `m`: the cached mapping is released for the next handle, any other mapping takes the cache's place
  void give(const uint8_t *m, uint64_t n, dev_t d, ino_t i) {
    std::lock_guard<std::mutex> g(mu);
    if (m == map) {
      in_use = false;
      return;
    }
    if (in_use) {
      g_unmapper.submit((char *)const_cast<uint8_t *>(m), n);
      return;
    }
    if (map) g_unmapper.submit((char *)const_cast<uint8_t *>(map), len);
    map = m, len = n, dev = d, ino = i;
  }
} g_map_cache;
}  // namespace

// ---- Pileup::Impl: device state of one pileup ------------------------------------------------------------------------
struct BamWriteJob;  // bam_write.inl: attached to a pass by Pileup::write_bam

struct Pileup::Impl {
  DeviceCtx *ctx = nullptr;
  BamWriteJob *wjob = nullptr;
  int fd = -1;
  const uint8_t *map = nullptr;      // read-only mapping of the BAM (the readers copy out of the page cache through it)
  uint64_t map_len = 0;
  dev_t map_dev = 0;
  ino_t map_ino = 0;
  std::vector<uint64_t> cut_points;  // sorted BGZF block starts known from the index
  // geometry + filter tables on the device
  DevBuf d_ref_len, d_first_bin, d_first_chunk;
  DevBuf d_rg, d_tagv, d_bl_off, d_bl_starts, d_bl_stops, d_bc_table, d_bc_off, d_bc_pool;
  FilterDev fdev{};
  PileDev pdev{};
  GeomDev gdev{};
  // accumulators
  DevBuf d_diff, d_stats, d_probe, d_err;
  // finalize products
  DevBuf d_tile_sum, d_tile_heads, d_run_bin, d_run_val, d_counts, d_fs, d_multi_vals;
  uint64_t n_runs = 0, multi_rows = 0;
  uint32_t max_val = 0;
  uint64_t fin_bins = 0;
  // device text between format_bedgraph and write_bedgraph_pieces
  std::unique_ptr<Pileup::TextJob> text_job;
  // merged output of a sharded job: pages allocated in the background by the one rank that was asked to (preallocate_output)
  std::thread prealloc_thread;
  std::atomic<bool> prealloc_stop{false};
  std::atomic<uint64_t> prealloc_done{0};
  // resident copy of the shard's compressed bytes
  DevBuf d_resident;
  std::vector<Segment> res_segs;
  std::vector<BlockTable> res_tabs;
  ~Impl() {
    const double t0 = now_ms();
    prealloc_stop.store(true);
    if (prealloc_thread.joinable()) prealloc_thread.join();
    const double t1 = now_ms();
    if (map) g_map_cache.give(map, map_len, map_dev, map_ino);  // kept for the next handle on this file, else torn down in the background
    if (fd >= 0) close(fd);
    const double t2 = now_ms();
    text_job.reset();
    for (DevBuf *b : {&d_ref_len, &d_first_bin, &d_first_chunk, &d_rg, &d_tagv, &d_bl_off, &d_bl_starts, &d_bl_stops, &d_bc_table,
                      &d_bc_off, &d_bc_pool, &d_diff, &d_stats, &d_probe, &d_err, &d_tile_sum, &d_tile_heads, &d_run_bin, &d_run_val,
                      &d_counts, &d_fs, &d_multi_vals, &d_resident})
      b->release();
    if (env_u64("BND_DEBUG_TIMING", 0))
      fprintf(stderr, "[bnd] destroy: allocator thread %.2f ms, unmap + close %.2f ms, device buffers %.2f ms\n", t1 - t0, t2 - t1, now_ms() - t2);
  }
};

namespace {
template <class T>
void upload(DevBuf &b, const T *src, size_t n) {
  b.ensure(std::max<size_t>(n * sizeof(T), 16));
  if (n) CU(cudaMemcpy(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice));
}
}  // namespace

int device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void device_inflate(const uint8_t *bgzf, uint64_t n_bytes, std::vector<uint8_t> &out, std::vector<int32_t> &status, int device) {
  DeviceCtx &ctx = ctx_for(device);
  std::lock_guard<std::mutex> g(ctx.mu);
  BlockTable tab;
  scan_bgzf_blocks(bgzf, n_bytes, tab);
  const uint64_t nblk = tab.coff.size() - 1;
  out.assign(tab.uoff.back(), 0);
  status.assign(nblk, 0);
  if (nblk == 0) return;
  DevBuf comp, coff, uoff, infl, st, err, work;
  try {
    comp.ensure(tab.coff.back() + COMP_PAD);
    CU(cudaMemcpy(comp.p, bgzf, tab.coff.back(), cudaMemcpyHostToDevice));
    upload(coff, tab.coff.data(), tab.coff.size());
    upload(uoff, tab.uoff.data(), tab.uoff.size());
    infl.ensure(tab.uoff.back() + 64);
    st.ensure(nblk * 4);
    err.ensure(16);
    work.ensure(16);
    CU(cudaMemset(err.p, 0, 16));
    CU(cudaMemset(work.p, 0, 16));
    InflateArgs a{};
    a.comp = comp.as<uint8_t>();
    a.blk_coff = coff.as<uint64_t>();
    a.blk_uoff = uoff.as<uint64_t>();
    a.out = infl.as<uint8_t>();
    a.status = st.as<int32_t>();
    a.error_flag = err.as<int32_t>();
    a.work_counter = work.as<uint32_t>();
    a.crc = ctx.d_crc;
    a.n_blocks = (uint32_t)nblk;
    a.verify_crc = 1;
    CU(launch_inflate(a, ctx.sm_count, ctx.sets[0].stream));
    CU(cudaStreamSynchronize(ctx.sets[0].stream));
    if (!out.empty()) CU(cudaMemcpy(out.data(), infl.p, out.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(status.data(), st.p, nblk * 4, cudaMemcpyDeviceToHost));
  } catch (...) {
    for (DevBuf *b : {&comp, &coff, &uoff, &infl, &st, &err, &work}) b->release();
    throw;
  }
  for (DevBuf *b : {&comp, &coff, &uoff, &infl, &st, &err, &work}) b->release();
}

// ---- construction ---------------------------------------------------------------------------------------------------
Pileup::Pileup(const bnd_pileup_cfg *cfg, const bnd_filter_cfg *filter, bool drop_scaffold_chunks) : im_(new Impl()) {
  if (!cfg || !cfg->bam_path) fail("missing BAM path", ERR_INVALID);
  cfg_ = *cfg;
  bam_path_ = cfg->bam_path;
  cfg_.bam_path = nullptr;
  if (cfg_.bin_size == 0) fail("bin size must be greater than 0", ERR_INVALID);
  const double t_c0 = now_ms();
  im_->ctx = &ctx_for(cfg_.device);
  const double t_c1 = now_ms();
  open();
  const double t_c2 = now_ms();
  bnd_filter_cfg dflt;
  memset(&dflt, 0, sizeof dflt);
  dflt.max_length = 0xffffffffu;
  dflt.barcode_csv_column = -1;
  filt_ = make_filter_host(filter ? filter : &dflt, hdr_);
  geo_ = make_geometry(hdr_, bai_, cfg_.bin_size, drop_scaffold_chunks);
  if (env_u64("BND_DEBUG_TIMING", 0))
    fprintf(stderr, "[bnd] create: device context %.2f ms, open (index, header, cut points) %.2f ms, filter + geometry %.2f ms\n", t_c1 - t_c0, t_c2 - t_c1,
            now_ms() - t_c2);

  // device copies of geometry and filter tables
  Impl &im = *im_;
  const size_t n_ref = hdr_.names.size();
  upload(im.d_ref_len, hdr_.lens.data(), n_ref);
  upload(im.d_first_bin, geo_.ref_first_bin.data(), n_ref + 1);
  upload(im.d_first_chunk, geo_.ref_first_chunk.data(), n_ref + 1);
  GeomDev &g = im.gdev;
  g.n_ref = (int32_t)n_ref;
  g.ref_len = im.d_ref_len.as<uint64_t>();
  g.ref_first_bin = im.d_first_bin.as<uint64_t>();
  g.ref_first_chunk = im.d_first_chunk.as<uint64_t>();
  g.L = geo_.L;
  g.bin = geo_.bin;
  g.nbf = geo_.nbf;

  FilterDev &f = im.fdev;
  const bnd_filter_cfg &c = filt_.cfg;
  f.strand = c.strand;
  f.proper_pair = c.proper_pair != 0;
  f.min_mapq = c.min_mapq;
  f.min_length = c.min_length;
  f.max_length = c.max_length;
  f.has_min_frag = c.has_min_fragment_length != 0;
  f.min_frag = c.min_fragment_length;
  f.has_max_frag = c.has_max_fragment_length != 0;
  f.max_frag = c.max_fragment_length;
  f.ex_dup = c.exclude_duplicates != 0;
  f.ex_sec = c.exclude_secondary != 0;
  f.ex_sup = c.exclude_supplementary != 0;
  if (filt_.has_rg) {
    upload(im.d_rg, filt_.rg.data(), filt_.rg.size());
    f.has_rg = 1;
    f.rg_len = (int32_t)filt_.rg.size();
    f.rg = im.d_rg.as<char>();
  }
  if (filt_.has_tag && filt_.has_tag_value) {  // `--tag` alone is ignored (read_filter.rs:769-771)
    upload(im.d_tagv, filt_.tag_value.data(), filt_.tag_value.size());
    f.has_tag = 1;
    f.tag_name_ok = filt_.tag.size() == 2;
    f.tag0 = f.tag_name_ok ? filt_.tag[0] : 0;
    f.tag1 = f.tag_name_ok ? filt_.tag[1] : 0;
    f.tagv_len = (int32_t)filt_.tag_value.size();
    f.tagv = im.d_tagv.as<char>();
  }
  if (filt_.has_blacklist) {
    upload(im.d_bl_off, filt_.bl_off.data(), filt_.bl_off.size());
    upload(im.d_bl_starts, filt_.bl_starts.data(), filt_.bl_starts.size());
    upload(im.d_bl_stops, filt_.bl_stops.data(), filt_.bl_stops.size());
    f.has_blacklist = 1;
    f.bl_off = im.d_bl_off.as<uint64_t>();
    f.bl_starts = im.d_bl_starts.as<uint64_t>();
    f.bl_stops = im.d_bl_stops.as<uint64_t>();
  }
  if (filt_.has_barcodes) {
    upload(im.d_bc_table, filt_.bc_table.data(), filt_.bc_table.size());
    upload(im.d_bc_off, filt_.bc_off.data(), filt_.bc_off.size());
    upload(im.d_bc_pool, filt_.bc_pool.data(), filt_.bc_pool.size());
    f.has_barcodes = 1;
    f.bc_mask = (uint32_t)filt_.bc_table.size() - 1;
    f.bc_table = im.d_bc_table.as<uint32_t>();
    f.bc_off = im.d_bc_off.as<uint32_t>();
    f.bc_pool = im.d_bc_pool.as<char>();
  }
  PileDev &p = im.pdev;
  p.use_fragment = cfg_.use_fragment != 0;
  if (cfg_.has_shift) {
    for (int i = 0; i < 4; i++) p.shift[i] = cfg_.shift[i];
    p.shift_on = p.shift[0] || p.shift[1] || p.shift[2] || p.shift[3];
  }
  if (cfg_.has_truncate) {
    p.has_trunc = 1;
    p.t_has5 = cfg_.trunc_has5 != 0;
    p.t5 = cfg_.trunc5;
    p.t_has3 = cfg_.trunc_has3 != 0;
    p.t3 = cfg_.trunc3;
  }
  // this process's shard of the genomic chunks: contiguous chunk ranges balanced by compressed bytes
  const uint64_t nchunks = geo_.total_chunks();
  uint64_t lo = 0, hi = nchunks;
  const int world = cfg_.shard_world > 1 ? cfg_.shard_world : 1;
  const int rank = world > 1 ? cfg_.shard_rank : 0;
  if (rank < 0 || rank >= world) fail("shard_rank out of range", ERR_INVALID);
  if (world > 1 && nchunks > 0) {
    auto boundary = [&](int k) -> uint64_t {  // first chunk of rank k
      if (k <= 0) return 0;
      if (k >= world) return nchunks;
      if (nchunks > (1u << 22)) return nchunks * (uint64_t)k / (uint64_t)world;  // degenerate geometry: split by count
      const uint64_t first = hdr_.first_record_voff >> 16;
      const uint64_t target = first + (file_size_ - first) * (uint64_t)k / (uint64_t)world;
      uint64_t a = 0, b = nchunks;  // smallest chunk whose lower-bound offset reaches the target
      while (a < b) {
        uint64_t mid = (a + b) / 2;
        int r = geo_.ref_of_chunk(mid);
        uint64_t cs = (mid - geo_.ref_first_chunk[r]) * geo_.L + 1;
        uint64_t v = voff_lower_bound(bai_, r, cs);
        uint64_t co = v == ~0ull ? file_size_ : (v >> 16);
        if (co < target) a = mid + 1;
        else b = mid;
      }
      return a;
    };
    lo = boundary(rank);
    hi = boundary(rank + 1);
    if (hi < lo) hi = lo;
  }
  shard_ = range_for_chunks(lo, hi);
  shape_.chunk_len = geo_.L;
  shape_.n_chunks = nchunks;
}

Pileup::~Pileup() = default;

void Pileup::open() {
  Impl &im = *im_;
  im.fd = ::open(bam_path_.c_str(), O_RDONLY);
  if (im.fd < 0) fail("File not found: " + bam_path_);
  struct stat st;
  if (fstat(im.fd, &st) != 0) fail("cannot stat " + bam_path_);
  file_size_ = (uint64_t)st.st_size;
  if (S_ISREG(st.st_mode) && file_size_ > 0) {
    im.map_dev = st.st_dev, im.map_ino = st.st_ino;
    if (const uint8_t *cached = g_map_cache.take(st)) {
      im.map = cached;
      im.map_len = file_size_;
    } else {
      void *m = mmap(nullptr, file_size_, PROT_READ, MAP_SHARED, im.fd, 0);
      if (m != MAP_FAILED) {
        im.map = (const uint8_t *)m;
        im.map_len = file_size_;
      }
    }
  }
  const double t_o0 = now_ms();
  bai_ = read_bai(bam_path_ + ".bai");
  const double t_o1 = now_ms();
  // header: inflate leading BGZF blocks on the device until the header parses
  uint64_t want = 1u << 18;
  std::vector<uint8_t> raw, inflated;
  std::vector<int32_t> status;
  for (;;) {
    want = std::min<uint64_t>(want, file_size_);
    raw.resize(want);
    uint64_t got = 0;
    while (got < want) {
      ssize_t r = pread(im.fd, raw.data() + got, want - got, (off_t)got);
      if (r <= 0) fail("read error on " + bam_path_);
      got += (uint64_t)r;
    }
    BlockTable tab;
    scan_bgzf_blocks(raw.data(), want, tab);
    if (tab.coff.size() < 2) {
      if (want == file_size_) fail("not a BGZF file: " + bam_path_);
      want *= 4;
      continue;
    }
    device_inflate(raw.data(), tab.coff.back(), inflated, status, im.ctx->device);
    for (size_t i = 0; i < status.size(); i++)
      if (status[i] != 0) fail("BGZF inflate failed in the BAM header (block " + std::to_string(i) + ", status " + std::to_string(status[i]) + ")");
    uint64_t consumed = 0;
    if (parse_bam_header(inflated.data(), inflated.size(), hdr_, consumed)) {
      // virtual offset of the first record: block holding inflated byte `consumed`
      size_t b = (size_t)(std::upper_bound(tab.uoff.begin(), tab.uoff.end(), consumed) - tab.uoff.begin()) - 1;
      if (b >= tab.coff.size() - 1) {  // header ends exactly at the end of the inflated data
        b = tab.coff.size() - 1;
        hdr_.first_record_voff = tab.coff[b] << 16;
      } else {
        hdr_.first_record_voff = (tab.coff[b] << 16) | (consumed - tab.uoff[b]);
      }
      break;
    }
    if (want == file_size_) fail("truncated BAM header: " + bam_path_);
    want *= 4;
  }
  const double t_o2 = now_ms();
  // block starts known from the index: segment cut points
  std::vector<uint64_t> &cp = im.cut_points;
  // The linear index alone is dense (one entry per 16 kb window) and already in file order for a sorted BAM: take it
  // with repeats dropped.  Sparse or unordered indexes also contribute their bin chunks and are sorted.
  bool ordered = true;
  for (const BaiRef &R : bai_.refs)
    for (uint64_t v : R.linear) {
      const uint64_t c = v >> 16;
      if (!v || (!cp.empty() && cp.back() == c)) continue;
      if (!cp.empty() && c < cp.back()) ordered = false;
      cp.push_back(c);
    }
  if (!ordered || cp.size() < 1024) {
    for (const BaiRef &R : bai_.refs)
      for (const BaiChunk &c : R.chunks) {
        cp.push_back(c.beg >> 16);
        cp.push_back(c.end >> 16);
      }
    std::sort(cp.begin(), cp.end());
    cp.erase(std::unique(cp.begin(), cp.end()), cp.end());
  }
  while (!cp.empty() && cp.back() >= file_size_) cp.pop_back();
  if (env_u64("BND_DEBUG_TIMING", 0))
    fprintf(stderr, "[bnd] open: index %.2f ms, header %.2f ms, cut points %.2f ms (%zu)\n", t_o1 - t_o0, t_o2 - t_o1, now_ms() - t_o2, cp.size());
}

Pileup::Range Pileup::range_for_chunks(uint64_t lo, uint64_t hi) const {
  Range r;
  r.chunk_lo = lo;
  r.chunk_hi = hi;
  const uint64_t nchunks = geo_.total_chunks();
  if (lo >= hi) {
    r.byte_begin = r.byte_end = file_size_;
    return r;
  }
  const uint64_t first = hdr_.first_record_voff;
  uint64_t v0;
  if (lo == 0) {
    v0 = first;
  } else {
    int ref = geo_.ref_of_chunk(lo);
    uint64_t cs = (lo - geo_.ref_first_chunk[ref]) * geo_.L + 1;
    v0 = voff_lower_bound(bai_, ref, cs);
    if (v0 != ~0ull && v0 < first) v0 = first;
  }
  if (v0 == ~0ull) {
    r.byte_begin = r.byte_end = file_size_;
    return r;
  }
  r.byte_begin = v0 >> 16;
  r.entry_uoff = v0 & 0xffff;
  if (hi >= nchunks) {
    r.byte_end = file_size_;
    return r;
  }
  int ref_e = geo_.ref_of_chunk(hi - 1);
  uint64_t ce = std::min((hi - geo_.ref_first_chunk[ref_e]) * geo_.L, hdr_.lens[ref_e]);
  r.bounded = true;
  r.end_ref = ref_e;
  r.end_pos = ce;
  uint64_t v1 = voff_lower_bound(bai_, ref_e, ce + 1 + 2 * 16384);
  if (v1 == ~0ull) {
    r.byte_end = file_size_;
  } else {
    const std::vector<uint64_t> &cp = im_->cut_points;
    auto it = std::upper_bound(cp.begin(), cp.end(), v1 >> 16);
    r.byte_end = it == cp.end() ? file_size_ : *it;
  }
  if (r.byte_end < r.byte_begin) r.byte_end = r.byte_begin;
  return r;
}

void Pileup::reset_accumulators(const Range &r) {
  Impl &im = *im_;
  GeomDev &g = im.gdev;
  g.chunk_lo = r.chunk_lo;
  g.chunk_hi = r.chunk_hi;
  g.bin_lo = geo_.first_bin_of_chunk(r.chunk_lo);
  g.bin_hi = geo_.first_bin_of_chunk(r.chunk_hi);
  const uint64_t nb = g.bin_hi - g.bin_lo;
  im.d_diff.ensure((nb + 1) * sizeof(int32_t) + 64);
  CU(cudaMemset(im.d_diff.p, 0, (nb + 1) * sizeof(int32_t)));
  im.d_stats.ensure(ST_N_COUNTERS * 8);
  CU(cudaMemset(im.d_stats.p, 0, ST_N_COUNTERS * 8));
  im.d_probe.ensure(16);
  CU(cudaMemset(im.d_probe.p, 0, 16));
  im.d_err.ensure(16);
  CU(cudaMemset(im.d_err.p, 0, 16));
  CU(cudaDeviceSynchronize());  // the pass runs on non-blocking streams: order it after the clears
  memset(stats_, 0, sizeof stats_);
  total_stats_set_ = false;
  finalized_ = false;
  im.n_runs = 0;
  im.max_val = 0;
}

// ---- the hot path: segments of compressed bytes -> inflate -> record index -> filter/scatter ----------------------------
#define BND_BAMNADO_VERSION "0.9.0"  // CARGO_PKG_VERSION of the reference crate this path stands in for (bamnado/Cargo.toml:3)
#include "bam_write.inl"

void Pileup::run_pass(const Range &range, bool resident) {
  Impl &im = *im_;
  DeviceCtx &ctx = *im.ctx;
  std::lock_guard<std::mutex> lock(ctx.mu);
  CU(cudaSetDevice(ctx.device));
  const double t_wall0 = now_ms();
  const unsigned long long launches0 = launch_count();
  ctx.seg_bytes = std::max<uint64_t>(env_u64("BND_SEGMENT_BYTES", 96ull << 20), 1u << 16);  // re-read per pass (tests shrink it)
  reset_accumulators(range);
  tm_ = Timings();
  shape_.records = shape_.comp_bytes = shape_.inflated_bytes = shape_.blocks = 0;
  shape_.bins = im.gdev.bin_hi - im.gdev.bin_lo;

  // ---- plan segments: cut at block starts known from the index ----
  std::vector<Segment> segs;
  const std::vector<uint64_t> &cp = im.cut_points;
  // optional ramp of the segment sizes (BND_FIRST_SEGMENT_BYTES, doubling up to seg_bytes); measured slower on B200
  // than uniform segments because small launches leave the SMs latency-bound, so it is off by default
  uint64_t ramp = std::min<uint64_t>(ctx.seg_bytes, env_u64("BND_FIRST_SEGMENT_BYTES", ctx.seg_bytes));
  auto next_cut = [&](uint64_t cur, uint64_t limit) -> uint64_t {  // a block start in (cur, limit], as far as the target allows
    uint64_t target = cur + ramp;
    ramp = std::min<uint64_t>(ctx.seg_bytes, ramp * 2);
    if (target >= limit) return limit;
    auto it = std::upper_bound(cp.begin(), cp.end(), target);  // first > target
    if (it != cp.begin()) {
      uint64_t c = *(it - 1);
      if (c > cur) return std::min(c, limit);
    }
    if (it != cp.end()) return std::min(*it, limit);
    return limit;
  };
  if (resident) {
    segs = im.res_segs;
  } else {
    for (uint64_t cur = range.byte_begin; cur < range.byte_end;) {
      uint64_t nx = next_cut(cur, range.byte_end);
      segs.push_back(Segment{cur, nx});
      cur = nx;
    }
  }

  const int n_slots = (int)ctx.slots.size();
  std::unique_ptr<ReaderPool> pool;
  double dbg_wait_read = 0, dbg_wait_h2d = 0, dbg_sync = 0;
  auto submit = [&](size_t i) {
    StageSlot &slot = ctx.slots[i % n_slots];
    if (slot.h2d_pending) {
      const double tw0 = now_ms();
      CU(cudaEventSynchronize(slot.ev_h2d));
      dbg_wait_h2d += now_ms() - tw0;
      slot.h2d_pending = false;
    }
    slot.ensure(segs[i].end - segs[i].begin + 64);
    pool->submit(i, (int)(i % n_slots));
  };
  if (!resident) {
    pool.reset(new ReaderPool(im.fd, env_u64("BND_READ_MMAP", 1) ? im.map : nullptr, ctx, segs, std::max(1, ctx.n_readers)));
    for (size_t i = 0; i < segs.size() && i < (size_t)n_slots; i++) submit(i);
  }

  struct SegEvents {
    cudaEvent_t e[5];  // h2d begin, inflate begin, index begin, pileup begin, end
  };
  std::vector<SegEvents> evs;
  const bool verify_crc = env_u64("BND_VERIFY_CRC", 1) != 0;
  const int use_window = (int)env_u64("BND_USE_WINDOW", 1);
  const bool serialize = env_u64("BND_SERIALIZE", 0) != 0;  // one segment at a time: clean per-kernel event timings
  const bool dbg_k1_only = env_u64("BND_DEBUG_K1_ONLY", 0) != 0;  // timing experiments only: K2-K4 are skipped, results are meaningless
  for (auto &s : ctx.sets) s.carry_pending = false;

  size_t i = 0;
  unsigned long long probe_host[2] = {0, 0};
  for (;;) {
    for (; i < segs.size(); i++) {
      const Segment &sg = segs[i];
      SegSet &S = ctx.sets[i % ctx.n_sets];
      const BlockTable *tab;
      const uint8_t *host_src = nullptr;
      if (resident) {
        tab = &im.res_tabs[i];
      } else {
        const double tw0 = now_ms();
        pool->wait(i);
        dbg_wait_read += now_ms() - tw0;
        StageSlot &slot = ctx.slots[i % n_slots];
        tab = &slot.tab;
        host_src = slot.pin;
      }
      const uint64_t clen = sg.end - sg.begin;
      const uint64_t nblk = tab->coff.size() - 1;
      const uint64_t ulen = tab->uoff.back();
      const uint64_t data_end = ctx.carry_cap + ulen;
      const uint32_t n_tiles = (uint32_t)((data_end + REC_TILE - 1) / REC_TILE);
      if (im.wjob) im.wjob->before_segment(ulen + ctx.carry_cap);  // room in every output stream, else compress + write first
      // the segment that followed this set's previous use must have taken its carry before the buffers are reused
      if (S.carry_pending) {
        CU(cudaStreamWaitEvent(S.stream, S.ev_carry, 0));
        S.carry_pending = false;
      }
      // every buffer of the set is checked on its own: the pool hands out buffers of differing slack, so one of them fitting says
      // nothing about the others
      const std::pair<DevBuf *, size_t> need[] = {{&S.comp, resident ? 0 : (size_t)(clen + COMP_PAD)}, {&S.coff, (size_t)(nblk + 1) * 8}, {&S.uoff, (size_t)(nblk + 1) * 8},
                                                  {&S.infl, (size_t)(data_end + 64)}, {&S.t_entry, (size_t)n_tiles * 8}, {&S.t_exit, (size_t)n_tiles * 8},
                                                  {&S.t_count, (size_t)n_tiles * 4}, {&S.t_flag, (size_t)n_tiles + 16}, {&S.t_base, (size_t)n_tiles * 8},
                                                  {&S.t_slots, (size_t)n_tiles * REC_SLOTS * 2}, {&S.rec_off, (size_t)(ulen / 36 + 2) * 8}};
      bool grow = false;
      for (auto &nd : need) grow = grow || nd.second > nd.first->cap;
      if (grow) {
        for (auto &q : ctx.sets) CU(cudaStreamSynchronize(q.stream));  // rare: nothing may still use the old buffers
        for (auto &nd : need) nd.first->ensure(nd.second);
      }
      evs.emplace_back();
      SegEvents &E = evs.back();
      for (auto &e : E.e) CU(cudaEventCreate(&e));
      cudaStream_t st = S.stream;
      CU(cudaEventRecord(E.e[0], st));
      const uint8_t *d_comp;
      if (resident) {
        d_comp = im.d_resident.as<uint8_t>() + (sg.begin - im.res_segs.front().begin);
      } else {
        CU(cudaMemcpyAsync(S.comp.p, host_src, clen, cudaMemcpyHostToDevice, st));
        d_comp = S.comp.as<uint8_t>();
      }
      CU(cudaMemcpyAsync(S.coff.p, tab->coff.data(), (nblk + 1) * 8, cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(S.uoff.p, tab->uoff.data(), (nblk + 1) * 8, cudaMemcpyHostToDevice, st));
      if (!resident) {
        StageSlot &slot = ctx.slots[i % n_slots];
        CU(cudaEventRecord(slot.ev_h2d, st));
        slot.h2d_pending = true;
      }
      CU(cudaMemsetAsync(S.work, 0, 4, st));
      CU(cudaEventRecord(E.e[1], st));
      InflateArgs ia{};
      ia.comp = d_comp;
      ia.blk_coff = S.coff.as<uint64_t>();
      ia.blk_uoff = S.uoff.as<uint64_t>();
      ia.out = S.infl.as<uint8_t>() + ctx.carry_cap;
      ia.status = nullptr;
      ia.error_flag = im.d_err.as<int32_t>();
      ia.work_counter = S.work;
      ia.crc = ctx.d_crc;
      ia.n_blocks = (uint32_t)nblk;
      ia.verify_crc = verify_crc;
      CU(launch_inflate(ia, ctx.sm_count, st));
      CU(cudaEventRecord(E.e[2], st));
      if (i == 0) {
        SegState s0;
        memset(&s0, 0, sizeof s0);
        s0.entry_off = ctx.carry_cap + range.entry_uoff;
        CU(cudaMemcpyAsync(S.st, &s0, sizeof s0, cudaMemcpyHostToDevice, st));
      } else {
        SegSet &P = ctx.sets[(i - 1) % ctx.n_sets];
        CU(cudaStreamWaitEvent(st, P.ev_idx, 0));
        CU(launch_carry_copy(P.infl.as<uint8_t>(), S.infl.as<uint8_t>(), P.st, S.st, ctx.carry_cap, st));
        CU(cudaEventRecord(P.ev_carry, st));
        P.carry_pending = true;
      }
      RecArgs ra{};
      ra.buf = S.infl.as<uint8_t>();
      ra.tile0 = 0;
      ra.data_end = data_end;
      ra.n_tiles = n_tiles;
      ra.n_ref = (int32_t)hdr_.names.size();
      ra.tile_entry = S.t_entry.as<uint64_t>();
      ra.tile_exit = S.t_exit.as<uint64_t>();
      ra.tile_count = S.t_count.as<uint32_t>();
      ra.tile_flag = S.t_flag.as<uint8_t>();
      ra.tile_base = S.t_base.as<uint64_t>();
      ra.tile_slots = S.t_slots.as<uint16_t>();
      ra.rec_off = S.rec_off.as<uint64_t>();
      ra.st = S.st;
      ra.carry_cap = ctx.carry_cap;
      if (!dbg_k1_only) CU(launch_record_index(ra, st));
      CU(cudaEventRecord(S.ev_idx, st));
      CU(cudaEventRecord(E.e[3], st));
      PileArgs pa{};
      pa.buf = S.infl.as<uint8_t>();
      pa.rec_off = S.rec_off.as<uint64_t>();
      pa.st = S.st;
      pa.diff = im.d_diff.as<int32_t>();
      pa.stats = im.d_stats.as<unsigned long long>();
      pa.range_probe = im.d_probe.as<unsigned long long>();
      pa.f = im.fdev;
      pa.p = im.pdev;
      pa.g = im.gdev;
      pa.use_window = use_window;
      if (im.wjob) im.wjob->segment(S, (int)(i % ctx.n_sets), ulen + ctx.carry_cap);
      else if (!dbg_k1_only) CU(launch_pileup(pa, ctx.sm_count, st));
      CU(cudaEventRecord(E.e[4], st));
      shape_.comp_bytes += clen;
      shape_.inflated_bytes += ulen;
      shape_.blocks += nblk;
      if (serialize) CU(cudaStreamSynchronize(st));
      // refill the staging ring
      if (!resident && i + n_slots < segs.size()) submit(i + n_slots);
    }
    {
      const double tw0 = now_ms();
      for (auto &s : ctx.sets) CU(cudaStreamSynchronize(s.stream));
      dbg_sync += now_ms() - tw0;
    }
    // a bounded range ends where the index says it should; keep going while the last record seen is still inside it
    if (!range.bounded || segs.empty() || segs.back().end >= file_size_) break;
    CU(cudaMemcpy(probe_host, im.d_probe.p, 16, cudaMemcpyDeviceToHost));
    const unsigned long long end_key = ((unsigned long long)(uint32_t)(range.end_ref + 1) << 32) | (uint32_t)range.end_pos;
    if (probe_host[0] > end_key) break;
    if (resident) fail("the staged byte range ends before the shard's last record; use accumulate()");
    uint64_t cur = segs.back().end;
    segs.push_back(Segment{cur, next_cut(cur, file_size_)});
    submit(segs.size() - 1);
  }
  pool.reset();

  // ---- collect: errors, counters, timings ----
  int32_t err[4] = {0, 0, 0, 0};
  CU(cudaMemcpy(err, im.d_err.p, 16, cudaMemcpyDeviceToHost));
  if (err[0] != 0)
    fail("BGZF inflate failed on the device (status " + std::to_string(err[0] & 0xff) + ", block " + std::to_string((unsigned)err[0] >> 8) + " of its segment)");
  if (!segs.empty()) {
    SegState last;
    CU(cudaMemcpy(&last, ctx.sets[(segs.size() - 1) % ctx.n_sets].st, sizeof last, cudaMemcpyDeviceToHost));
    if (last.error == SEG_ERR_BAD_RECORD) fail("invalid BAM record (block_size < 32)");
    if (last.error == SEG_ERR_CARRY_TOO_LARGE) fail("a BAM record larger than BND_CARRY_BYTES straddles two segments");
    shape_.records = last.total_records;
    if (env_u64("BND_DEBUG_TIMING", 0))
      fprintf(stderr, "[bnd] record index: %u fix rounds, %u tiles re-walked over %zu segments\n", last.fix_rounds, last.rewalked, segs.size());
    if (!range.bounded && last.carry_len != 0) fail("truncated BAM record at end of file");
  }
  unsigned long long st[ST_N_COUNTERS];
  CU(cudaMemcpy(st, im.d_stats.p, sizeof st, cudaMemcpyDeviceToHost));
  for (int k = 0; k < ST_N_COUNTERS; k++) stats_[k] = st[k];
  for (auto &E : evs) {
    float ms;
    CU(cudaEventElapsedTime(&ms, E.e[0], E.e[1]));
    tm_.h2d_ms += ms;
    CU(cudaEventElapsedTime(&ms, E.e[1], E.e[2]));
    tm_.inflate_ms += ms;
    CU(cudaEventElapsedTime(&ms, E.e[2], E.e[3]));
    tm_.index_ms += ms;
    CU(cudaEventElapsedTime(&ms, E.e[3], E.e[4]));
    tm_.pileup_ms += ms;
    for (int k = 0; k < 5; k++) cudaEventDestroy(E.e[k]);
  }
  tm_.wall_ms = now_ms() - t_wall0;
  tm_.read_ms = dbg_wait_read;
  if (env_u64("BND_DEBUG_TIMING", 0))
    fprintf(stderr, "[bnd] pass: %zu segments, wall %.2f ms, waited on readers %.2f ms, on staging reuse %.2f ms, final sync %.2f ms\n", segs.size(),
            tm_.wall_ms, dbg_wait_read, dbg_wait_h2d, dbg_sync);
  shape_.launches = launch_count() - launches0;
  shape_.segments = segs.size();
  // the accumulators now hold `range`; only a pass over the shard's own range feeds the whole-object outputs
  accumulated_ = range.chunk_lo == shard_.chunk_lo && range.chunk_hi == shard_.chunk_hi && range.byte_begin == shard_.byte_begin;
}

void Pileup::accumulate() { run_pass(shard_, false); }

void Pileup::stage() {
  Impl &im = *im_;
  DeviceCtx &ctx = *im.ctx;
  std::lock_guard<std::mutex> lock(ctx.mu);
  CU(cudaSetDevice(ctx.device));
  im.res_segs.clear();
  im.res_tabs.clear();
  Range r = shard_;
  if (r.bounded && r.byte_end < file_size_) {  // margin for records of the last chunk that sit past the planned end
    const std::vector<uint64_t> &cps = im.cut_points;
    auto it = std::upper_bound(cps.begin(), cps.end(), r.byte_end + ctx.seg_bytes);
    r.byte_end = it == cps.end() ? file_size_ : *it;
  }
  const uint64_t total = r.byte_end - r.byte_begin;
  im.d_resident.ensure(total + COMP_PAD);
  if (total == 0) return;
  // one pass over the file through a pinned bounce buffer; segment tables are kept on the host
  const std::vector<uint64_t> &cp = im.cut_points;
  std::vector<uint8_t> tmp;
  for (uint64_t cur = r.byte_begin; cur < r.byte_end;) {
    uint64_t target = cur + ctx.seg_bytes, nx = r.byte_end;
    if (target < r.byte_end) {
      auto it = std::upper_bound(cp.begin(), cp.end(), target);
      if (it != cp.begin() && *(it - 1) > cur) nx = std::min(*(it - 1), r.byte_end);
      else if (it != cp.end()) nx = std::min(*it, r.byte_end);
    }
    const uint64_t len = nx - cur;
    tmp.resize(len);
    uint64_t got = 0;
    while (got < len) {
      ssize_t k = pread(im.fd, tmp.data() + got, len - got, (off_t)(cur + got));
      if (k <= 0) fail("read error on " + bam_path_);
      got += (uint64_t)k;
    }
    BlockTable tab;
    scan_bgzf_blocks(tmp.data(), len, tab);
    if (tab.coff.back() != len) fail("BAM segment does not end on a BGZF block boundary (corrupt file or index)");
    CU(cudaMemcpy(im.d_resident.as<uint8_t>() + (cur - r.byte_begin), tmp.data(), len, cudaMemcpyHostToDevice));
    im.res_segs.push_back(Segment{cur, nx});
    im.res_tabs.push_back(std::move(tab));
    cur = nx;
  }
}

void Pileup::accumulate_resident() {
  if (im_->res_segs.empty() && shard_.byte_end > shard_.byte_begin) fail("accumulate_resident called before stage", ERR_INVALID);
  run_pass(shard_, true);
}

void Pileup::set_total_stats(const uint64_t s[ST_N_COUNTERS]) {
  for (int k = 0; k < ST_N_COUNTERS; k++) stats_[k] = s[k];
  total_stats_set_ = true;
}

void Pileup::stats(uint64_t out[ST_N_COUNTERS]) const {
  for (int k = 0; k < ST_N_COUNTERS; k++) out[k] = stats_[k];
}

double Pileup::norm_factor() const {
  // NormalizationMethod::scale_factor (signal_normalization.rs:50-74) over n_reads_after_filtering (read_filter.rs:229-243)
  if (cfg_.shard_world > 1 && cfg_.norm_method != 0 && !total_stats_set_)
    fail("sharded pile-up: CPM/RPKM needs the job-wide filter counters (bnd_pileup_set_total_stats) before normalising", ERR_INVALID);
  uint64_t failed = 0;
  for (int k = 1; k < ST_N_COUNTERS; k++) failed += stats_[k];
  return bnd_scale_factor(cfg_.norm_method, cfg_.scale_factor, cfg_.bin_size, stats_[0] - failed);
}

// ---- K5: scan + collapse ------------------------------------------------------------------------------------------------
uint64_t Pileup::finalize_range(bool dense_counts, bool emit_runs) {
  Impl &im = *im_;
  DeviceCtx &ctx = *im.ctx;
  std::lock_guard<std::mutex> lock(ctx.mu);
  CU(cudaSetDevice(ctx.device));
  cudaStream_t st = ctx.sets[0].stream;
  const GeomDev &g = im.gdev;
  const uint64_t nb = g.bin_hi - g.bin_lo;
  const uint64_t n_tiles64 = (nb + FIN_TILE - 1) / FIN_TILE;
  if (n_tiles64 > 0xffffffffull) fail("too many bins for one shard");
  const uint32_t n_tiles = (uint32_t)n_tiles64;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  im.d_tile_sum.ensure((size_t)n_tiles * 4 + 16);
  im.d_tile_heads.ensure((size_t)n_tiles * 8 + 16);
  im.d_fs.ensure(sizeof(FinState));
  CU(cudaMemsetAsync(im.d_fs.p, 0, sizeof(FinState), st));
  FinArgs a{};
  a.diff = im.d_diff.as<int32_t>();
  a.n_bins = nb;
  a.g = g;
  a.collapse = cfg_.collapse != 0;
  a.tile_sum = im.d_tile_sum.as<int32_t>();
  a.tile_heads = im.d_tile_heads.as<uint64_t>();
  a.n_tiles = n_tiles;
  a.fs = im.d_fs.as<FinState>();
  CU(cudaEventRecord(e0, st));
  CU(launch_bins_reduce(a, ctx.sm_count, st));
  CU(launch_tile_scan(a, st));
  FinState fs;
  CU(cudaMemcpyAsync(&fs, im.d_fs.p, sizeof fs, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  im.n_runs = emit_runs ? fs.n_runs : 0;
  if (emit_runs) {
    im.d_run_bin.ensure((size_t)im.n_runs * 8 + 16);
    im.d_run_val.ensure((size_t)im.n_runs * 4 + 16);
    a.run_bin = im.d_run_bin.as<uint64_t>();
    a.run_val = im.d_run_val.as<uint32_t>();
  }
  if (dense_counts) {
    im.d_counts.ensure((size_t)nb * 4 + 16);
    a.counts = im.d_counts.as<uint32_t>();
  }
  CU(launch_bins_emit(a, ctx.sm_count, st));
  CU(cudaMemcpyAsync(&fs, im.d_fs.p, sizeof fs, cudaMemcpyDeviceToHost, st));
  CU(cudaEventRecord(e1, st));
  CU(cudaStreamSynchronize(st));
  im.max_val = fs.max_val;
  im.fin_bins = nb;
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  tm_.finalize_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  finalized_ = true;
  return im.n_runs;
}

uint64_t Pileup::finalize() {
  if (!accumulated_) fail("finalize called before accumulate", ERR_INVALID);
  return finalize_range(false);
}

void Pileup::ensure_finalized() {
  if (!accumulated_) accumulate();
  if (!finalized_) finalize_range(false);
}

void Pileup::fetch_rows(std::vector<RunRow> &out) {
  Impl &im = *im_;
  DeviceCtx &ctx = *im.ctx;
  std::lock_guard<std::mutex> lock(ctx.mu);
  CU(cudaSetDevice(ctx.device));
  out.resize(im.n_runs);
  if (im.n_runs == 0) return;
  DevBuf rows;
  rows.ensure(im.n_runs * sizeof(RunRow));
  RunArgs a{};
  a.run_bin = im.d_run_bin.as<uint64_t>();
  a.run_val = im.d_run_val.as<uint32_t>();
  a.n_runs = im.n_runs;
  a.g = im.gdev;
  a.rows = rows.as<RunRow>();
  cudaStream_t st = ctx.sets[0].stream;
  try {
    CU(launch_expand_runs(a, ctx.sm_count, st));
    CU(cudaMemcpyAsync(out.data(), rows.p, im.n_runs * sizeof(RunRow), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  } catch (...) {
    rows.release();
    throw;
  }
  rows.release();
}

void Pileup::runs(std::vector<RunRow> &out) {
  ensure_finalized();
  fetch_rows(out);
}

void Pileup::chromosome(const std::string &chrom, std::vector<RunRow> &out) {
  auto it = hdr_.name_to_id.find(chrom);
  if (it == hdr_.name_to_id.end() || !geo_.in_stats[it->second]) fail("Chromosome " + chrom + " not found", ERR_NOT_FOUND);
  const int r = it->second;
  Range range = range_for_chunks(geo_.ref_first_chunk[r], geo_.ref_first_chunk[r + 1]);
  run_pass(range, false);
  finalize_range(false);
  fetch_rows(out);
  // the bin array, counters and rows now describe this contig only: a later whole-object call starts over
  // (pileup_chromosome does not change what to_bedgraph writes, coverage_analysis.rs:337-430)
  accumulated_ = finalized_ = false;
}

void Pileup::signal(const std::string &chrom, std::vector<float> &out) {
  auto it = hdr_.name_to_id.find(chrom);
  if (it == hdr_.name_to_id.end()) fail("Chromosome " + chrom + " not found in BAM file.", ERR_NOT_FOUND);
  const int r = it->second;
  if (!geo_.in_stats[r]) fail("Chromosome " + chrom + " not found");
  Impl &im = *im_;
  // the Python entry point passes no shift / truncate and always piles up raw counts (lib.rs:275-287)
  PileDev saved = im.pdev;
  im.pdev.shift_on = 0;
  im.pdev.has_trunc = 0;
  Range range = range_for_chunks(geo_.ref_first_chunk[r], geo_.ref_first_chunk[r + 1]);
  try {
    run_pass(range, false);
  } catch (...) {
    im.pdev = saved;
    throw;
  }
  im.pdev = saved;
  finalize_range(true);
  accumulated_ = finalized_ = false;  // single-contig state, see chromosome()
  DeviceCtx &ctx = *im.ctx;
  std::lock_guard<std::mutex> lock(ctx.mu);
  const uint64_t n = hdr_.lens[r];
  out.assign(n, 0.f);
  if (n == 0) return;
  DevBuf d_out;
  d_out.ensure(n * 4);
  SignalArgs a{};
  a.counts = im.d_counts.as<uint32_t>();
  a.g = im.gdev;
  a.ref = r;
  a.chrom_size = n;
  a.scale = cfg_.scale_factor;
  a.out = d_out.as<float>();
  cudaStream_t st = ctx.sets[0].stream;
  try {
    CU(launch_dense_signal(a, ctx.sm_count, st));
    CU(cudaMemcpyAsync(out.data(), d_out.p, n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  } catch (...) {
    d_out.release();
    throw;
  }
  d_out.release();
}

// ---- BamStats::is_paired_end ----------------------------------------------------------------------------------------------
bool Pileup::is_paired_end() {
  Impl &im = *im_;
  const uint64_t begin = hdr_.first_record_voff >> 16;
  uint64_t want = std::min<uint64_t>(file_size_ - begin, 4u << 20);
  std::vector<uint8_t> raw(want), inflated;
  std::vector<int32_t> status;
  uint64_t got = 0;
  while (got < want) {
    ssize_t r = pread(im.fd, raw.data() + got, want - got, (off_t)(begin + got));
    if (r <= 0) fail("Failed to open BAM file to check paired-end status");
    got += (uint64_t)r;
  }
  BlockTable tab;
  scan_bgzf_blocks(raw.data(), want, tab, 64);
  device_inflate(raw.data(), tab.coff.back(), inflated, status, im.ctx->device);
  uint64_t p = hdr_.first_record_voff & 0xffff;
  for (int i = 0; i <= 1000; i++) {
    if (p + 36 > inflated.size()) break;
    uint32_t bs = (uint32_t)inflated[p] | ((uint32_t)inflated[p + 1] << 8) | ((uint32_t)inflated[p + 2] << 16) | ((uint32_t)inflated[p + 3] << 24);
    if (bs < 32) break;
    uint32_t flag = (uint32_t)inflated[p + 18] | ((uint32_t)inflated[p + 19] << 8);
    if (flag & 0x1) return true;
    p += 4ull + bs;
  }
  return false;
}

// ---- MultiBamPileup ---------------------------------------------------------------------------------------------------------
namespace {
// Shared run heads of the table described by `a` (its own head test, or the precomputed bit mask) -> run_bin[n_rows] and the
// counts of every BAM of `a` at those heads -> vals[n_bams][n_rows].  Returns n_rows.
uint64_t multi_emit_table(DeviceCtx &ctx, MultiArgs &a, DevBuf &run_bin, DevBuf &vals, cudaStream_t st) {
  const uint64_t n_tiles64 = (a.n_bins + 255) / 256;
  if (n_tiles64 > 0xffffffffull) fail("too many bins for one shard");
  a.n_tiles = (uint32_t)n_tiles64;
  if (a.n_bins == 0) return 0;
  DevBuf d_tiles, d_total;
  unsigned long long n_runs = 0;
  try {
    d_tiles.ensure((size_t)a.n_tiles * 8 + 16);
    d_total.ensure(16);
    a.tile_heads = d_tiles.as<uint64_t>();
    CU(launch_multi_count(a, ctx.sm_count, st));
    CU(launch_u64_scan(a.tile_heads, a.n_tiles, d_total.as<unsigned long long>(), st));
    CU(cudaMemcpyAsync(&n_runs, d_total.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    run_bin.ensure(n_runs * 8 + 16);
    vals.ensure(n_runs * 4 * (size_t)a.n_bams + 16);
    a.run_bin = run_bin.as<uint64_t>();
    a.out_vals = vals.as<uint32_t>();
    a.n_runs_cap = n_runs;
    CU(launch_multi_emit(a, ctx.sm_count, st));
    CU(cudaStreamSynchronize(st));
  } catch (...) {
    d_tiles.release();
    d_total.release();
    throw;
  }
  d_tiles.release();
  d_total.release();
  return n_runs;
}
}  // namespace

// MultiBamPileup::to_tsv with every BAM on this device (coverage_analysis.rs:783-864): per-BAM raw counts (one u32 per bin stays
// in HBM for every BAM), shared run heads, and the TSV text itself are produced on the device; the host writes the header line
// and moves the per-reference pieces into (chrom byte string, start) order like collapse_equal_bins sorts them.
void Pileup::multi_to_tsv(std::vector<Pileup *> &bams, const std::string &path) {
  if (bams.empty()) fail("At least one BAM pileup is required", ERR_INVALID);
  if ((int)bams.size() > MULTI_MAX_BAMS) fail("too many BAM files for one multi-bam-coverage call", ERR_INVALID);
  Pileup &first = *bams[0];
  for (Pileup *b : bams) {
    if (b->cfg_.bin_size != first.cfg_.bin_size) fail("All BAM pileups must have the same bin size", ERR_INVALID);
    if (b->cfg_.collapse) fail("All BAM pileups must have collapse set to false", ERR_INVALID);
    if (b->hdr_.names != first.hdr_.names || b->hdr_.lens != first.hdr_.lens || b->geo_.L != first.geo_.L || b->geo_.in_stats != first.geo_.in_stats)
      fail("multi-bam-coverage on the B200 path needs BAMs with the same reference dictionary and chunk length");
    if (b->im_->ctx != first.im_->ctx) fail("all BAMs of one multi-bam-coverage call must use the same device", ERR_INVALID);
  }
  MultiArgs a{};
  a.n_bams = (int)bams.size();
  for (size_t i = 0; i < bams.size(); i++) {
    Pileup &b = *bams[i];
    b.multi_prepare(false);
    a.counts[i] = b.im_->d_counts.as<uint32_t>();
  }
  std::string header = "chrom\tstart\tend";
  for (Pileup *b : bams) header += "\t" + b->bam_path_;
  header.push_back('\n');
  Impl &im = *first.im_;
  DeviceCtx &ctx = *im.ctx;
  std::vector<uint64_t> mine;
  {
    std::lock_guard<std::mutex> lock(ctx.mu);
    CU(cudaSetDevice(ctx.device));
    a.g = im.gdev;
    a.n_bins = a.g.bin_hi - a.g.bin_lo;
    im.multi_rows = multi_emit_table(ctx, a, im.d_run_bin, im.d_multi_vals, ctx.sets[0].stream);
  }
  first.multi_format_tsv(im.d_multi_vals.as<uint32_t>(), im.multi_rows, (int)bams.size(), nullptr, 0, im.multi_rows, mine, false);
  first.multi_write_tsv_pieces(path, header.c_str(), mine.data(), 1, 0);
}

// bin_counts::count_bins (bamnado/src/bin_counts.rs:100-289).  Every sample runs the ordinary pile-up without collapsing
// (`bin_intervals(.., false)`, :240) and without shift/truncate (`IntervalMaker::new(.., None, None)`, :219-227); its dense
// counts are scattered on the device into column `s` of one matrix whose rows follow the shared grid: chromosomes sorted
// by name (:150-154), bins [1 + i * bin, min((i + 1) * bin, len)] (:158-170).  The samples are processed one after the other
// like the reference's outer loop (:178); the matrix stays in HBM until the last one is done.
void Pileup::count_bins(const bnd_pileup_cfg *cfgs, const bnd_filter_cfg *filters, int n_samples, BinCounts &out) {
  if (n_samples < 1 || !cfgs) fail("At least one BAM file is required for bin counting", ERR_INVALID);
  const uint64_t bin = cfgs[0].bin_size;
  const bool drop_scaffold = cfgs[0].ignore_scaffold != 0;
  const bool use_fragment = cfgs[0].use_fragment != 0;
  out = BinCounts();
  out.bin_size = bin;
  out.library_sizes.assign((size_t)n_samples, 0);
  std::map<std::string, uint64_t> grid_sizes;  // chromsizes_ref_name() of the first sample
  std::map<std::string, uint64_t> row0_of;
  DevBuf d_matrix, d_row0;
  DeviceCtx *ctx0 = nullptr;
  auto release = [&] {
    d_matrix.release();
    d_row0.release();
  };
  try {
    for (int s = 0; s < n_samples; s++) {
      bnd_pileup_cfg c = cfgs[s];
      if (c.bin_size != bin) fail("all samples must use the same bin size", ERR_INVALID);
      c.collapse = 0;
      c.norm_method = 0;  // raw counts
      c.scale_factor = 1.0f;
      c.use_fragment = use_fragment;
      c.ignore_scaffold = 0;  // handled by the chunk list, not at output time
      c.has_shift = c.has_truncate = 0;
      c.shard_rank = 0;
      c.shard_world = 1;
      c.device = cfgs[0].device;
      Pileup p(&c, filters ? &filters[s] : nullptr, drop_scaffold);
      std::map<std::string, uint64_t> sizes;
      for (size_t r = 0; r < p.hdr_.names.size(); r++)
        if (p.geo_.in_stats[r]) sizes[p.hdr_.names[r]] = p.hdr_.lens[r];
      if (s == 0) {
        grid_sizes = sizes;
        ctx0 = p.im_->ctx;
        uint64_t rows = 0;
        for (auto &kv : grid_sizes) {  // std::map iterates in byte-string order = the sorted chromosome names
          if (drop_scaffold && is_scaffold(kv.first)) continue;
          out.chroms.push_back(kv.first);
          out.chrom_len.push_back(kv.second);
          out.chrom_first_row.push_back(rows);
          row0_of[kv.first] = rows;
          rows += (kv.second + bin - 1) / bin;  // `while start <= length` (:161)
        }
        out.chrom_first_row.push_back(rows);
        out.n_bins = rows;
        if (rows == 0) fail("No bins produced; check --bin-size and --ignore-scaffolds");
        std::lock_guard<std::mutex> lock(ctx0->mu);
        CU(cudaSetDevice(ctx0->device));
        d_matrix.ensure((size_t)rows * (size_t)n_samples * 8 + 16);
        CU(cudaMemset(d_matrix.p, 0, (size_t)rows * (size_t)n_samples * 8));
      } else if (sizes != grid_sizes) {
        fail("Chromosome name/length mismatch between BAM files: " + p.bam_path_ + " does not match " + std::string(cfgs[0].bam_path) +
             ". bam-normalize requires all input BAM files to share an identical set of reference sequences so that a common bin grid can be used.");
      }
      p.multi_prepare();  // pile-up + dense raw counts on the device
      uint64_t st[ST_N_COUNTERS];
      p.stats(st);
      uint64_t failed = 0;
      for (int k = 1; k < ST_N_COUNTERS; k++) failed += st[k];
      out.library_sizes[(size_t)s] = st[ST_N_TOTAL] - failed;  // n_reads_after_filtering (read_filter.rs:229-243)
      std::vector<uint64_t> row0(p.hdr_.names.size(), ~0ull);
      for (size_t r = 0; r < row0.size(); r++) {
        auto it = row0_of.find(p.hdr_.names[r]);
        if (it != row0_of.end() && p.geo_.in_stats[r]) row0[r] = it->second;
      }
      Impl &im = *p.im_;
      DeviceCtx &ctx = *im.ctx;
      std::lock_guard<std::mutex> lock(ctx.mu);
      CU(cudaSetDevice(ctx.device));
      cudaStream_t strm = ctx.sets[0].stream;
      d_row0.ensure(row0.size() * 8 + 16);
      CU(cudaMemcpyAsync(d_row0.p, row0.data(), row0.size() * 8, cudaMemcpyHostToDevice, strm));
      BinMatrixArgs a{};
      a.counts = im.d_counts.as<uint32_t>();
      a.g = im.gdev;
      a.grid_row0 = d_row0.as<uint64_t>();
      a.matrix = d_matrix.as<unsigned long long>();
      a.bins_per_chunk_grid = p.geo_.L / bin;
      a.n_samples = (uint32_t)n_samples;
      a.sample = (uint32_t)s;
      CU(launch_bin_matrix(a, ctx.sm_count, strm));
      CU(cudaStreamSynchronize(strm));
    }
    out.counts.resize((size_t)out.n_bins * (size_t)n_samples);
    {
      std::lock_guard<std::mutex> lock(ctx0->mu);
      CU(cudaSetDevice(ctx0->device));
      CU(cudaMemcpy(out.counts.data(), d_matrix.p, out.counts.size() * 8, cudaMemcpyDeviceToHost));
    }
  } catch (...) {
    release();
    throw;
  }
  release();
}

uint64_t Pileup::multi_prepare(bool resident) {
  if (cfg_.collapse) fail("All BAM pileups must have collapse set to false", ERR_INVALID);
  if (resident && im_->res_segs.empty() && shard_.byte_end > shard_.byte_begin) fail("multi_prepare_resident called before stage", ERR_INVALID);
  run_pass(shard_, resident);
  finalize_range(true, false);
  // only the dense counts are needed from here on: the difference array (as large as the counts; 12.4 GB for hg38 at bin 1) goes
  // back to the buffer pool so that the next BAM of this rank reuses it
  {
    std::lock_guard<std::mutex> lock(im_->ctx->mu);
    im_->d_diff.release();
  }
  accumulated_ = false;  // the accumulators are gone; a later whole-object call runs its own pass
  return im_->gdev.bin_hi - im_->gdev.bin_lo;
}

// ---- one BAM per rank --------------------------------------------------------------------------------------------
// 1. multi_prepare: pile-up + dense raw counts;  2. multi_change_bits: "my count changes here", one BIT per bin;  3. the caller
// gathers the masks of all ranks (NCCL all-gather over NVLink; hg38 at bin 1: 387 MB per rank) and multi_or_bits folds them into
// the shared run heads;  4. multi_compact_column: this rank's counts at those heads stay on the device;  5. the caller exchanges
// column slices (all-to-all: every rank receives rows [k n / world, (k + 1) n / world) of every column);  6. multi_format_tsv
// writes the text of that slice on the device;  7. after an all-gather of the per-reference byte counts every rank fills its
// pieces of the one TSV (multi_write_tsv_pieces).
namespace {
void check_multi_group(Pileup *const *ps, int n) {
  if (n < 1 || !ps) fail("At least one BAM pileup is required", ERR_INVALID);
  if (n > MULTI_MAX_BAMS) fail("too many BAM files for one multi-bam-coverage call", ERR_INVALID);
}
}  // namespace

void Pileup::multi_change_bits(Pileup *const *ps, int n, uint32_t *dev_bits) {
  check_multi_group(ps, n);
  Impl &im = *ps[0]->im_;
  DeviceCtx &ctx = *im.ctx;
  MultiArgs a{};
  a.n_bams = n;
  for (int i = 0; i < n; i++) {
    Impl &q = *ps[i]->im_;
    if (!ps[i]->finalized_ || !q.d_counts.p) fail("multi_change_bits called before multi_prepare", ERR_INVALID);
    if (q.ctx != im.ctx || q.gdev.bin_lo != im.gdev.bin_lo || q.gdev.bin_hi != im.gdev.bin_hi) fail("the BAMs of one rank must share device and bin grid", ERR_INVALID);
    a.counts[i] = q.d_counts.as<uint32_t>();
  }
  std::lock_guard<std::mutex> lock(ctx.mu);
  CU(cudaSetDevice(ctx.device));
  a.g = im.gdev;
  a.n_bins = a.g.bin_hi - a.g.bin_lo;
  a.bits_out = dev_bits;
  CU(launch_multi_bits(a, ctx.sm_count, ctx.sets[0].stream));
  CU(cudaStreamSynchronize(ctx.sets[0].stream));
}

void Pileup::multi_or_bits(int device, uint32_t *dst, const uint32_t *src, uint32_t n_src, uint64_t n_words) {
  DeviceCtx &ctx = ctx_for(device);
  std::lock_guard<std::mutex> lock(ctx.mu);
  CU(launch_multi_or(dst, src, n_src, n_words, ctx.sm_count, ctx.sets[0].stream));
  CU(cudaStreamSynchronize(ctx.sets[0].stream));
}

// counts of the n BAMs of this rank at the shared row heads: vals[n][n_rows], kept (with the head bins) in ps[0]
uint64_t Pileup::multi_compact_columns(Pileup *const *ps, int n, const uint32_t *dev_bits, const uint32_t **dev_vals) {
  check_multi_group(ps, n);
  Impl &im = *ps[0]->im_;
  DeviceCtx &ctx = *im.ctx;
  MultiArgs a{};
  a.n_bams = n;
  for (int i = 0; i < n; i++) {
    Impl &q = *ps[i]->im_;
    if (!ps[i]->finalized_ || !q.d_counts.p) fail("multi_compact_columns called before multi_prepare", ERR_INVALID);
    if (q.ctx != im.ctx) fail("the BAMs of one rank must share the device", ERR_INVALID);
    a.counts[i] = q.d_counts.as<uint32_t>();
  }
  std::lock_guard<std::mutex> lock(ctx.mu);
  CU(cudaSetDevice(ctx.device));
  a.g = im.gdev;
  a.n_bins = a.g.bin_hi - a.g.bin_lo;
  a.bits = dev_bits;
  im.multi_rows = multi_emit_table(ctx, a, im.d_run_bin, im.d_multi_vals, ctx.sets[0].stream);
  if (dev_vals) *dev_vals = im.d_multi_vals.as<uint32_t>();
  return im.multi_rows;
}

// TSV text of rows [row_lo, row_hi) of the shared table; dev_cols[c * stride + (row - row_lo)] is BAM c's count of that row.
void Pileup::multi_format_tsv(const uint32_t *dev_cols, uint64_t stride, int n_cols, const int32_t *col_order, uint64_t row_lo, uint64_t row_hi,
                              std::vector<uint64_t> &bytes_per_ref, bool sizes_only) {
  Impl &im = *im_;
  DeviceCtx &ctx = *im.ctx;
  const size_t n_ref = hdr_.names.size();
  if (row_lo > row_hi || row_hi > im.multi_rows) fail("row slice outside the multi-BAM table", ERR_INVALID);
  if (n_cols < 1 || n_cols > MULTI_MAX_BAMS) fail("bad column count", ERR_INVALID);
  std::lock_guard<std::mutex> lock(ctx.mu);
  CU(cudaSetDevice(ctx.device));
  cudaStream_t st = ctx.sets[0].stream;
  im.text_job.reset(new TextJob());
  TextJob &job = *im.text_job;
  job.span.assign(n_ref, {0, 0});
  bytes_per_ref.assign(n_ref, 0);
  const uint64_t n_rows = row_hi - row_lo;
  if (n_rows == 0) return;
  try {
    std::string names;
    std::vector<uint32_t> name_off(n_ref + 1, 0);
    for (size_t r = 0; r < n_ref; r++) {
      names += hdr_.names[r];
      name_off[r + 1] = (uint32_t)names.size();
    }
    upload(job.d_names, names.data(), names.size());
    upload(job.d_name_off, name_off.data(), name_off.size());
    const uint64_t n_tiles64 = (n_rows + TEXT_THREADS - 1) / TEXT_THREADS;
    if (n_tiles64 > 0xffffffffull) fail("too many rows for one shard");
    job.d_row_len.ensure(n_rows * 2 + 16);
    job.d_tile_off.ensure(n_tiles64 * 8 + 16);
    job.d_ref_off.ensure((n_ref + 1) * 8);
    job.d_skip.ensure(16);  // total
    CU(cudaMemsetAsync(job.d_ref_off.p, 0xff, (n_ref + 1) * 8, st));
    MultiTextArgs a{};
    a.run_bin = im.d_run_bin.as<uint64_t>();
    a.n_runs = im.multi_rows;
    a.row_lo = row_lo;
    a.n_rows = n_rows;
    a.vals = dev_cols;
    a.stride = stride;
    a.n_cols = n_cols;
    for (int c = 0; c < n_cols; c++) {
      const int pos = col_order ? col_order[c] : c;
      if (pos < 0 || pos >= n_cols) fail("bad column order", ERR_INVALID);
      a.col_pos[c] = (uint8_t)pos;
    }
    a.g = im.gdev;
    a.names = job.d_names.as<char>();
    a.name_off = job.d_name_off.as<uint32_t>();
    a.row_len = job.d_row_len.as<uint16_t>();
    a.tile_off = job.d_tile_off.as<uint64_t>();
    a.n_tiles = (uint32_t)n_tiles64;
    a.ref_text_off = job.d_ref_off.as<uint64_t>();
    CU(launch_multi_text_len(a, ctx.sm_count, st));
    CU(launch_u64_scan(a.tile_off, a.n_tiles, job.d_skip.as<unsigned long long>(), st));
    unsigned long long total = 0;
    CU(cudaMemcpyAsync(&total, job.d_skip.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    job.nbytes = total;
    if (!sizes_only) {
      job.d_text.ensure(job.nbytes + 16);
      a.text = job.d_text.as<char>();
    }
    CU(launch_multi_text_write(a, ctx.sm_count, st));
    std::vector<uint64_t> ref_off(n_ref + 1);
    CU(cudaMemcpyAsync(ref_off.data(), job.d_ref_off.p, (n_ref + 1) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    uint64_t next = job.nbytes;
    for (size_t r = n_ref; r-- > 0;)
      if (ref_off[r] != ~0ull) {
        job.span[r] = {ref_off[r], next};
        next = ref_off[r];
      }
    for (size_t r = 0; r < n_ref; r++) bytes_per_ref[r] = job.span[r].second - job.span[r].first;
  } catch (...) {
    im.text_job.reset();
    throw;
  }
  if (sizes_only) im.text_job.reset();
}

// Every rank fills its pieces of the one TSV: header line (written by rank 0), then the references in byte-string order and
// inside a reference the row slices in rank order.  bytes_all is [world][n_refs].
void Pileup::multi_write_tsv_pieces(const std::string &path, const char *header, const uint64_t *bytes_all, int world, int rank) {
  Impl &im = *im_;
  if (!im.text_job) fail("multi_write_tsv_pieces called before multi_format_tsv", ERR_INVALID);
  if (world < 1 || rank < 0 || rank >= world || !bytes_all) fail("bad shard table", ERR_INVALID);
  g_unmapper.reap();
  std::lock_guard<std::mutex> lock(im.ctx->mu);
  std::unique_ptr<TextJob> job = std::move(im.text_job);
  const uint64_t base = header ? strlen(header) : 0;
  const uint64_t total = place_pieces(*job, bytes_all, world, rank, base);
  int fd = ::open(path.c_str(), O_RDWR | O_CREAT, 0644);
  if (fd < 0) fail("cannot create " + path);
  struct stat st;
  if (fstat(fd, &st) != 0 || ((uint64_t)st.st_size != total && ftruncate(fd, (off_t)total) != 0)) {
    close(fd);
    fail("cannot size " + path);
  }
  if (rank == 0 && base) {
    uint64_t done = 0;
    while (done < base) {
      ssize_t w = pwrite(fd, header + done, base - done, (off_t)done);
      if (w <= 0) {
        close(fd);
        fail("write error on " + path);
      }
      done += (uint64_t)w;
    }
  }
  write_pieces(*job, fd, path, total, 0, false, nullptr, 0, world > 1);
}

// ---- collapse-bedgraph ------------------------------------------------------------------------------------------------------
namespace {
// host bytes -> device through the pinned staging slots: `threads` copy every 64 MiB piece into a slot (non-temporal stores), the
// copy engine takes it from there while the next piece is being filled
void upload_bytes(DeviceCtx &ctx, const uint8_t *src, uint64_t n, uint8_t *d_dst, cudaStream_t st) {
  const uint64_t CH = 64ull << 20;
  const int n_slots = (int)ctx.slots.size();
  const int threads = (int)std::min<uint64_t>(16, std::max<uint64_t>(2, host_cores_share() * 3 / 4));
  for (auto &sl : ctx.slots) sl.h2d_pending = false;
  uint64_t k = 0;
  for (uint64_t off = 0; off < n; off += CH, k++) {
    StageSlot &slot = ctx.slots[k % n_slots];
    if (slot.h2d_pending) {
      CU(cudaEventSynchronize(slot.ev_h2d));
      slot.h2d_pending = false;
    }
    const uint64_t len = std::min(CH, n - off);
    slot.ensure(len + 64);
    std::vector<std::thread> th;
    const uint64_t per = (len + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
      const uint64_t a0 = std::min(len, per * t), a1 = std::min(len, a0 + per);
      if (a1 > a0) th.emplace_back([&, a0, a1] { stream_copy(slot.pin + a0, src + off + a0, a1 - a0); });
    }
    for (auto &t : th) t.join();
    CU(cudaMemcpyAsync(d_dst + off, slot.pin, len, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(slot.ev_h2d, st));
    slot.h2d_pending = true;
  }
  CU(cudaStreamSynchronize(st));
  for (auto &sl : ctx.slots) sl.h2d_pending = false;
}

// str::parse::<f64> for the spellings the device leaves to the host (more than 19 significant digits, large exponents, inf / nan)
bool parse_f64_host(const char *p, size_t n, double &out) {
  if (n == 0 || n > 4000) return false;
  for (size_t i = 0; i < n; i++) {
    const char c = p[i];
    if (c == 'x' || c == 'X' || c == '(' || c == 'p' || c == 'P') return false;  // strtod's hex floats and nan(...) are not Rust floats
  }
  char buf[4001];
  memcpy(buf, p, n);
  buf[n] = 0;
  char *end = nullptr;
  out = strtod(buf, &end);
  return end == buf + n;
}

// collapse_adjacent_bins over device columns sorted or not (bedgraph_utils.rs:5-45): head flags, scan, compaction
void collapse_device_columns(DeviceCtx &ctx, cudaStream_t st, DevBuf &d_chrom, DevBuf &d_start, DevBuf &d_end, DevBuf &d_score, uint64_t n, BdgRows &out) {
  DevBuf d_perm, d_flag, d_tiles, d_total, o_chrom, o_start, o_end, o_score, g_chrom, g_start, g_end, g_score;
  auto release = [&] {
    for (DevBuf *b : {&d_perm, &d_flag, &d_tiles, &d_total, &o_chrom, &o_start, &o_end, &o_score, &g_chrom, &g_start, &g_end, &g_score}) b->release();
  };
  try {
    d_flag.ensure(16);
    CU(cudaMemsetAsync(d_flag.p, 0, 16, st));
    BdgArgs a{};
    a.chrom = d_chrom.as<uint32_t>();
    a.start = d_start.as<long long>();
    a.end = d_end.as<long long>();
    a.score = d_score.as<double>();
    a.n = n;
    const uint64_t n_tiles64 = (n + 255) / 256;
    if (n_tiles64 > 0xffffffffull) fail("too many bedGraph rows");
    a.n_tiles = (uint32_t)n_tiles64;
    a.unsorted = d_flag.as<unsigned int>();
    CU(launch_bdg_sorted_check(a, ctx.sm_count, st));
    unsigned int unsorted = 0;
    CU(cudaMemcpyAsync(&unsorted, d_flag.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (unsorted) {
      // sort (chrom, start): permutation from the device radix sort (library), columns gathered on the device
      d_perm.ensure(n * 4);
      CU(sort_rows_by_chrom_start(a.chrom, a.start, n, d_perm.as<uint32_t>(), st));
      g_chrom.ensure(n * 4);
      g_start.ensure(n * 8);
      g_end.ensure(n * 8);
      g_score.ensure(n * 8);
      CU(launch_bdg_gather_rows(d_perm.as<uint32_t>(), n, a.chrom, a.start, a.end, a.score, g_chrom.as<uint32_t>(), g_start.as<long long>(), g_end.as<long long>(),
                                g_score.as<double>(), ctx.sm_count, st));
      a.chrom = g_chrom.as<uint32_t>();
      a.start = g_start.as<long long>();
      a.end = g_end.as<long long>();
      a.score = g_score.as<double>();
    }
    d_tiles.ensure((size_t)a.n_tiles * 8 + 16);
    d_total.ensure(16);
    a.tile_heads = d_tiles.as<uint64_t>();
    CU(launch_bdg_count(a, ctx.sm_count, st));
    CU(launch_u64_scan(a.tile_heads, a.n_tiles, d_total.as<unsigned long long>(), st));
    unsigned long long m = 0;
    CU(cudaMemcpyAsync(&m, d_total.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::vector<long long> init_min(m, INT64_MAX), init_max(m, INT64_MIN);
    o_chrom.ensure(m * 4 + 16);
    o_score.ensure(m * 8 + 16);
    upload(o_start, init_min.data(), m);
    upload(o_end, init_max.data(), m);
    a.out_chrom = o_chrom.as<uint32_t>();
    a.out_start = o_start.as<long long>();
    a.out_end = o_end.as<long long>();
    a.out_score = o_score.as<double>();
    CU(launch_bdg_emit(a, ctx.sm_count, st));
    out.chrom.resize(m), out.start.resize(m), out.end.resize(m), out.score.resize(m);
    if (m) {
      CU(cudaMemcpyAsync(out.chrom.data(), o_chrom.p, m * 4, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(out.start.data(), o_start.p, m * 8, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(out.end.data(), o_end.p, m * 8, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(out.score.data(), o_score.p, m * 8, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
  } catch (...) {
    release();
    throw;
  }
  release();
}
}  // namespace

// The whole of `collapse-bedgraph` between the file read and the text write (src/cli.rs:1149-1205): the raw bytes go to HBM,
// lines / fields / numbers are parsed there, chromosome names become ranks (the host only sees the handful of distinct names),
// and the rows are collapsed.  `names` returns the chromosome names in byte-string order = the meaning of out.chrom.
void collapse_bedgraph_text(const uint8_t *text, uint64_t n, std::vector<std::string> &names, BdgRows &out, int device) {
  out = BdgRows();
  names.clear();
  if (n == 0) return;
  DeviceCtx &ctx = ctx_for(device);
  std::lock_guard<std::mutex> lock(ctx.mu);
  cudaStream_t st = ctx.sets[0].stream;
  DevBuf d_text, d_tiles, d_total, d_line_off, d_state, d_start, d_end, d_score, d_name_off, d_name_len, d_slow, d_cnt, d_vals;
  DevBuf r_start, r_end, r_score, r_name_off, r_name_len, d_flag, d_seg_off, d_seg_len, d_seg_rank, d_chrom;
  auto release = [&] {
    for (DevBuf *b : {&d_text, &d_tiles, &d_total, &d_line_off, &d_state, &d_start, &d_end, &d_score, &d_name_off, &d_name_len, &d_slow, &d_cnt, &d_vals, &r_start,
                      &r_end, &r_score, &r_name_off, &r_name_len, &d_flag, &d_seg_off, &d_seg_len, &d_seg_rank, &d_chrom})
      b->release();
  };
  const bool dbg = env_u64("BND_DEBUG_TIMING", 0) != 0;
  const double t0 = now_ms();
  try {
    d_text.ensure(n + 64);
    const double t_alloc = now_ms();
    upload_bytes(ctx, text, n, d_text.as<uint8_t>(), st);
    const double t1 = now_ms();
    // ---- lines ----
    const uint64_t n_tiles64 = (n + BDG_TILE - 1) / BDG_TILE;
    if (n_tiles64 > 0xffffffffull) fail("bedGraph too large");
    BdgParseArgs pa{};
    pa.text = d_text.as<uint8_t>();
    pa.n_bytes = n;
    pa.n_tiles = (uint32_t)n_tiles64;
    d_tiles.ensure(n_tiles64 * 8 + 16);
    d_total.ensure(16);
    pa.tile_count = d_tiles.as<uint64_t>();
    CU(launch_bdg_newline_count(pa, ctx.sm_count, st));
    CU(launch_u64_scan(pa.tile_count, pa.n_tiles, d_total.as<unsigned long long>(), st));
    unsigned long long n_nl = 0;
    CU(cudaMemcpyAsync(&n_nl, d_total.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const bool ends_nl = text[n - 1] == '\n';
    const uint64_t n_lines = n_nl + (ends_nl ? 0 : 1);
    pa.n_lines = n_lines;
    d_line_off.ensure((n_lines + 1) * 8);
    pa.line_off = d_line_off.as<uint64_t>();
    CU(launch_bdg_line_starts(pa, ctx.sm_count, st));
    const uint64_t sentinel = ends_nl ? n : n + 1;
    CU(cudaMemcpyAsync(pa.line_off + n_lines, &sentinel, 8, cudaMemcpyHostToDevice, st));
    // ---- fields and numbers ----
    d_state.ensure(n_lines + 16);
    d_start.ensure(n_lines * 8);
    d_end.ensure(n_lines * 8);
    d_score.ensure(n_lines * 8);
    d_name_off.ensure(n_lines * 8);
    d_name_len.ensure(n_lines * 4);
    d_cnt.ensure(16);
    pa.state = d_state.as<uint8_t>();
    pa.start = d_start.as<long long>();
    pa.end = d_end.as<long long>();
    pa.score = d_score.as<double>();
    pa.name_off = d_name_off.as<uint64_t>();
    pa.name_len = d_name_len.as<uint32_t>();
    pa.n_slow = d_cnt.as<unsigned long long>();
    pa.bad_line = d_cnt.as<unsigned long long>() + 1;
    unsigned long long cnt[2] = {0, 0};
    pa.slow_cap = 1u << 20;
    for (int attempt = 0; attempt < 2; attempt++) {
      d_slow.ensure(pa.slow_cap * sizeof(BdgSlow));
      pa.slow = d_slow.as<BdgSlow>();
      const unsigned long long init[2] = {0ull, ~0ull};
      CU(cudaMemcpyAsync(d_cnt.p, init, 16, cudaMemcpyHostToDevice, st));
      CU(launch_bdg_parse_lines(pa, ctx.sm_count, st));
      CU(cudaMemcpyAsync(cnt, d_cnt.p, 16, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      if (cnt[0] <= pa.slow_cap) break;
      pa.slow_cap = cnt[0];  // more scores for the host than the list holds: size it and parse again
    }
    auto line_text = [&](uint64_t li) {  // for error messages: the line as the reference would print it
      std::vector<uint64_t> lo(2);
      CU(cudaMemcpy(lo.data(), pa.line_off + li, 16, cudaMemcpyDeviceToHost));
      uint64_t e = std::min<uint64_t>(lo[1] - 1, n);
      return std::string((const char *)text + lo[0], (const char *)text + std::min(e, lo[0] + 200));
    };
    if (cnt[1] != ~0ull) fail("invalid bedGraph line: " + line_text(cnt[1]));
    const uint64_t n_slow = cnt[0];
    if (n_slow) {
      std::vector<BdgSlow> slow(n_slow);
      CU(cudaMemcpy(slow.data(), d_slow.p, n_slow * sizeof(BdgSlow), cudaMemcpyDeviceToHost));
      std::vector<double> vals(n_slow);
      std::atomic<uint64_t> bad{~0ull};
      const int threads = (int)std::min<uint64_t>(std::max<uint64_t>(1, n_slow / 4096), std::max(2u, host_cores_share()));
      std::vector<std::thread> th;
      for (int t = 0; t < threads; t++)
        th.emplace_back([&, t] {
          for (uint64_t k = (uint64_t)t; k < n_slow; k += (uint64_t)threads)
            if (!parse_f64_host((const char *)text + slow[k].off, slow[k].len, vals[k])) {
              uint64_t cur = bad.load();
              while (slow[k].line < cur && !bad.compare_exchange_weak(cur, slow[k].line)) {
              }
            }
        });
      for (auto &t : th) t.join();
      if (bad.load() != ~0ull) fail("invalid bedGraph line: " + line_text(bad.load()));
      upload(d_vals, vals.data(), n_slow);
      CU(launch_bdg_scatter_slow(d_slow.as<BdgSlow>(), d_vals.as<double>(), n_slow, pa.score, ctx.sm_count, st));
    }
    const double t2 = now_ms();
    // ---- rows: drop the skipped lines ----
    BdgRowArgs ra{};
    ra.state = pa.state;
    ra.n_lines = n_lines;
    const uint64_t lt64 = (n_lines + COL_TILE - 1) / COL_TILE;
    if (lt64 > 0xffffffffull) fail("too many bedGraph lines");
    ra.n_tiles = (uint32_t)lt64;
    d_tiles.ensure(lt64 * 8 + 16);
    ra.tile_count = d_tiles.as<uint64_t>();
    CU(launch_bdg_row_count(ra, ctx.sm_count, st));
    CU(launch_u64_scan(ra.tile_count, ra.n_tiles, d_total.as<unsigned long long>(), st));
    unsigned long long n_rows = 0;
    CU(cudaMemcpyAsync(&n_rows, d_total.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (n_rows == 0) {
      release();
      return;
    }
    DevBuf *c_start = &d_start, *c_end = &d_end, *c_score = &d_score, *c_name_off = &d_name_off, *c_name_len = &d_name_len;
    if (n_rows != n_lines) {
      r_start.ensure(n_rows * 8);
      r_end.ensure(n_rows * 8);
      r_score.ensure(n_rows * 8);
      r_name_off.ensure(n_rows * 8);
      r_name_len.ensure(n_rows * 4);
      ra.start = pa.start, ra.end = pa.end, ra.score = pa.score, ra.name_off = pa.name_off, ra.name_len = pa.name_len;
      ra.o_start = r_start.as<long long>(), ra.o_end = r_end.as<long long>(), ra.o_score = r_score.as<double>();
      ra.o_name_off = r_name_off.as<uint64_t>(), ra.o_name_len = r_name_len.as<uint32_t>();
      CU(launch_bdg_compact_rows(ra, ctx.sm_count, st));
      c_start = &r_start, c_end = &r_end, c_score = &r_score, c_name_off = &r_name_off, c_name_len = &r_name_len;
    }
    // ---- chromosome names -> ranks in byte-string order (polars sorts the string column bytewise) ----
    BdgNameArgs na{};
    na.text = pa.text;
    na.name_off = c_name_off->as<uint64_t>();
    na.name_len = c_name_len->as<uint32_t>();
    na.n = n_rows;
    const uint64_t rt64 = (n_rows + COL_TILE - 1) / COL_TILE;
    na.n_tiles = (uint32_t)rt64;
    d_flag.ensure(n_rows + 16);
    na.flag = d_flag.as<uint8_t>();
    na.tile_count = d_tiles.as<uint64_t>();
    CU(launch_bdg_name_heads(na, ctx.sm_count, st));
    CU(launch_u64_scan(na.tile_count, na.n_tiles, d_total.as<unsigned long long>(), st));
    unsigned long long n_seg = 0;
    CU(cudaMemcpyAsync(&n_seg, d_total.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    d_seg_off.ensure(n_seg * 8);
    d_seg_len.ensure(n_seg * 4);
    na.seg_name_off = d_seg_off.as<uint64_t>();
    na.seg_name_len = d_seg_len.as<uint32_t>();
    CU(launch_bdg_name_segments(na, ctx.sm_count, st));
    std::vector<uint64_t> seg_off(n_seg);
    std::vector<uint32_t> seg_len(n_seg), seg_rank(n_seg);
    CU(cudaMemcpyAsync(seg_off.data(), d_seg_off.p, n_seg * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(seg_len.data(), d_seg_len.p, n_seg * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    {
      std::map<std::string, uint32_t> rank_of;
      std::vector<std::map<std::string, uint32_t>::iterator> it_of(n_seg, rank_of.end());
      for (uint64_t k = 0; k < n_seg; k++) it_of[k] = rank_of.emplace(std::string((const char *)text + seg_off[k], seg_len[k]), 0u).first;
      uint32_t r = 0;
      for (auto &kv : rank_of) {
        kv.second = r++;
        names.push_back(kv.first);
      }
      for (uint64_t k = 0; k < n_seg; k++) seg_rank[k] = it_of[k]->second;
    }
    upload(d_seg_rank, seg_rank.data(), n_seg);
    na.seg_rank = d_seg_rank.as<uint32_t>();
    d_chrom.ensure(n_rows * 4);
    na.chrom = d_chrom.as<uint32_t>();
    CU(launch_bdg_assign_chrom(na, ctx.sm_count, st));
    const double t3 = now_ms();
    // the text and the per-line scratch are no longer needed: make room for the sort / the output
    for (DevBuf *b : {&d_text, &d_line_off, &d_state, &d_slow, &d_vals, &d_flag, &d_seg_off, &d_seg_len}) b->release();
    collapse_device_columns(ctx, st, d_chrom, *c_start, *c_end, *c_score, n_rows, out);
    if (dbg)
      fprintf(stderr, "[bnd] collapse-bedgraph: %llu bytes, %llu lines, %llu rows, %llu scores parsed on the host, %llu name segments; device buffer %.2f ms, upload %.2f ms, lines + parse %.2f ms, "
                      "rows + names %.2f ms, collapse %.2f ms -> %zu rows\n",
              (unsigned long long)n, (unsigned long long)n_lines, (unsigned long long)n_rows, (unsigned long long)n_slow, (unsigned long long)n_seg, t_alloc - t0, t1 - t_alloc,
              t2 - t1, t3 - t2, now_ms() - t3, out.chrom.size());
  } catch (...) {
    release();
    throw;
  }
  release();
}

// ---- bedGraph text ------------------------------------------------------------------------------------------------------
// The rows are formatted on the device in header order; write_bedgraph's (chrom byte string, start) order is a
// permutation of whole per-reference pieces, applied while the bytes leave the pinned bounce buffers.
namespace {


// copies [a, b) of the device text, as seen through `pieces`, to its destinations via `sink(src, len, out_off)`
template <class Sink>
void scatter_chunk(const std::vector<TextPiece> &pieces, const uint8_t *src, uint64_t a, uint64_t b, Sink sink) {
  // pieces are sorted by dev_begin
  size_t lo = 0, hi = pieces.size();
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    if (pieces[mid].dev_end <= a) lo = mid + 1;
    else hi = mid;
  }
  for (size_t k = lo; k < pieces.size() && pieces[k].dev_begin < b; k++) {
    uint64_t x = std::max(a, pieces[k].dev_begin), y = std::min(b, pieces[k].dev_end);
    if (y > x) sink(src + (x - a), y - x, pieces[k].out_off + (x - pieces[k].dev_begin));
  }
}
}  // namespace


// Formats every row on the device; returns the device text and how its per-reference pieces map to the output order.
void Pileup::format_text(TextJob &job) {
  Impl &im = *im_;
  DeviceCtx &ctx = *im.ctx;
  const size_t n_ref = hdr_.names.size();
  job.nbytes = 0;
  job.pieces.clear();
  job.span.assign(n_ref, {0, 0});
  if (im.n_runs == 0) return;
  // score text per distinct raw count: (count as f64) * factor, printed like a polars Float64 column
  const double factor = norm_factor();
  const uint64_t n_scores = (uint64_t)im.max_val + 1;
  if (n_scores > (1ull << 26)) fail("bin counts too large for the score table");
  std::vector<char> score_text(n_scores * SCORE_W, 0);
  {
    const int nt = (int)std::max<uint64_t>(1, std::min<uint64_t>(std::thread::hardware_concurrency(), n_scores / 4096 + 1));
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
      th.emplace_back([&, t] {
        for (uint64_t v = (uint64_t)t; v < n_scores; v += (uint64_t)nt) {
          std::string s = format_f64((double)(uint32_t)v * factor);
          if (s.size() > SCORE_W - 1) s.resize(SCORE_W - 1);
          score_text[v * SCORE_W] = (char)s.size();
          memcpy(&score_text[v * SCORE_W + 1], s.data(), s.size());
        }
      });
    for (auto &t : th) t.join();
  }
  std::string names;
  std::vector<uint32_t> name_off(n_ref + 1, 0);
  std::vector<uint8_t> skip(n_ref, 0);
  size_t max_name = 0;
  for (size_t r = 0; r < n_ref; r++) {
    names += hdr_.names[r];
    name_off[r + 1] = (uint32_t)names.size();
    max_name = std::max(max_name, hdr_.names[r].size());
    skip[r] = cfg_.ignore_scaffold && is_scaffold(hdr_.names[r]);
  }
  if (max_name > 60000) fail("reference name too long");
  CU(cudaSetDevice(ctx.device));
  cudaStream_t st = ctx.sets[0].stream;
  upload(job.d_names, names.data(), names.size());
  upload(job.d_name_off, name_off.data(), name_off.size());
  upload(job.d_skip, skip.data(), skip.size());
  upload(job.d_scores, score_text.data(), score_text.size());
  const uint64_t n_tiles64 = (im.n_runs + TEXT_THREADS - 1) / TEXT_THREADS;
  if (n_tiles64 > 0xffffffffull) fail("too many rows for one shard");
  job.d_row_len.ensure(im.n_runs * 2 + 16);
  job.d_tile_off.ensure(n_tiles64 * 8 + 16);
  job.d_ref_off.ensure((n_ref + 1) * 8);
  CU(cudaMemsetAsync(job.d_ref_off.p, 0xff, (n_ref + 1) * 8, st));
  TextArgs a{};
  a.run_bin = im.d_run_bin.as<uint64_t>();
  a.run_val = im.d_run_val.as<uint32_t>();
  a.n_runs = im.n_runs;
  a.g = im.gdev;
  a.names = job.d_names.as<char>();
  a.name_off = job.d_name_off.as<uint32_t>();
  a.ref_skip = job.d_skip.as<uint8_t>();
  a.score_text = job.d_scores.as<char>();
  a.n_scores = (uint32_t)n_scores;
  a.row_len = job.d_row_len.as<uint16_t>();
  a.tile_off = job.d_tile_off.as<uint64_t>();
  a.n_tiles = (uint32_t)n_tiles64;
  a.fs = im.d_fs.as<FinState>();
  a.ref_text_off = job.d_ref_off.as<uint64_t>();
  CU(launch_text_len(a, ctx.sm_count, st));
  CU(launch_text_scan(a, st));
  FinState fs;
  CU(cudaMemcpyAsync(&fs, im.d_fs.p, sizeof fs, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  job.nbytes = fs.text_bytes;
  job.d_text.ensure(job.nbytes + 16);
  a.text = job.d_text.as<char>();
  CU(launch_text_write(a, ctx.sm_count, st));
  std::vector<uint64_t> ref_off(n_ref + 1);
  CU(cudaMemcpyAsync(ref_off.data(), job.d_ref_off.p, (n_ref + 1) * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  // device text is in header order; where each reference's rows go is decided by place_pieces()
  uint64_t next = job.nbytes;
  for (size_t r = n_ref; r-- > 0;)
    if (ref_off[r] != ~0ull) {
      job.span[r] = {ref_off[r], next};
      next = ref_off[r];
    }
}

// write_bedgraph sorts by (chrom as a byte string, start) (coverage_analysis.rs:63): the output is the references in name
// order, and inside one reference the shards in rank order (a shard is a contiguous range of genomic chunks).
// bytes_all[k * n_ref + r] = text bytes shard k holds for reference r.  Returns the size of the whole file.
uint64_t Pileup::place_pieces(TextJob &job, const uint64_t *bytes_all, int world, int rank, uint64_t base) const {
  const size_t n_ref = hdr_.names.size();
  std::vector<int> order(n_ref);
  for (size_t r = 0; r < n_ref; r++) order[r] = (int)r;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return hdr_.names[x] < hdr_.names[y]; });
  job.pieces.clear();
  uint64_t out_off = base;
  for (int r : order)
    for (int k = 0; k < world; k++) {
      const uint64_t n = bytes_all[(size_t)k * n_ref + (size_t)r];
      if (k == rank) {
        if (n != job.span[r].second - job.span[r].first) fail("bedGraph piece table does not match this shard's text", ERR_INVALID);
        if (n) job.pieces.push_back(TextPiece{job.span[r].first, job.span[r].second, out_off});
      }
      out_off += n;
    }
  std::sort(job.pieces.begin(), job.pieces.end(), [](const TextPiece &x, const TextPiece &y) { return x.dev_begin < y.dev_begin; });
  return out_off;
}

// Streams the device text through the pinned bounce buffers; `sink(src, len, out_off)` may be called concurrently.
template <class Sink>
static void drain_text(DeviceCtx &ctx, const uint8_t *d_text, uint64_t nbytes, const std::vector<TextPiece> &pieces, int n_threads, Sink sink) {
  if (nbytes == 0) return;
  const uint64_t CH = std::min<uint64_t>(env_u64("BND_TEXT_CHUNK_BYTES", 64ull << 20), nbytes);
  ctx.bounce.ensure(CH);
  cudaStream_t st = ctx.sets[0].stream;
  const uint64_t n_chunks = (nbytes + CH - 1) / CH;
  // worker pool consuming (chunk, sub-range) tasks
  struct Task {
    const uint8_t *src;
    uint64_t a, b;
  };
  std::mutex mu;
  std::condition_variable cv, cv_done;
  std::deque<Task> q;
  int outstanding[PinnedBounce::N] = {0, 0};
  bool stop = false;
  std::string err;
  auto worker = [&] {
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> g(mu);
        cv.wait(g, [&] { return stop || !q.empty(); });
        if (q.empty()) return;
        t = q.front();
        q.pop_front();
      }
      try {
        scatter_chunk(pieces, t.src, t.a, t.b, sink);
      } catch (const std::exception &e) {
        std::lock_guard<std::mutex> g(mu);
        if (err.empty()) err = e.what();
      }
      {
        std::lock_guard<std::mutex> g(mu);
        for (int i = 0; i < PinnedBounce::N; i++)
          if (t.src >= ctx.bounce.buf[i] && t.src < ctx.bounce.buf[i] + ctx.bounce.cap) outstanding[i]--;
      }
      cv_done.notify_all();
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < std::max(1, n_threads); i++) th.emplace_back(worker);
  auto finish = [&] {
    {
      std::lock_guard<std::mutex> g(mu);
      stop = true;
    }
    cv.notify_all();
    for (auto &t : th) t.join();
  };
  try {
    for (uint64_t k = 0; k < n_chunks + 1; k++) {
      if (k < n_chunks) {
        const int b = (int)(k % PinnedBounce::N);
        {  // the buffer must have been fully consumed
          std::unique_lock<std::mutex> g(mu);
          cv_done.wait(g, [&] { return outstanding[b] == 0; });
        }
        const uint64_t a0 = k * CH, a1 = std::min(nbytes, a0 + CH);
        CU(cudaMemcpyAsync(ctx.bounce.buf[b], d_text + a0, a1 - a0, cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(ctx.bounce.ev[b], st));
      }
      if (k > 0) {  // hand the previous chunk to the workers while this one is in flight
        const uint64_t j = k - 1;
        const int b = (int)(j % PinnedBounce::N);
        CU(cudaEventSynchronize(ctx.bounce.ev[b]));
        const uint64_t a0 = j * CH, a1 = std::min(nbytes, a0 + CH);
        const uint64_t SUB = 4ull << 20;
        std::lock_guard<std::mutex> g(mu);
        for (uint64_t x = a0; x < a1; x += SUB) {
          q.push_back(Task{ctx.bounce.buf[b] + (x - a0), x, std::min(a1, x + SUB)});
          outstanding[b]++;
        }
        cv.notify_all();
      }
    }
    {
      std::unique_lock<std::mutex> g(mu);
      cv_done.wait(g, [&] { return outstanding[0] == 0 && outstanding[1] == 0; });
    }
  } catch (...) {
    finish();
    throw;
  }
  finish();
  if (!err.empty()) fail(err);
}

void Pileup::bedgraph_text(std::string &out) {
  ensure_finalized();
  DeviceCtx &ctx = *im_->ctx;
  const double t0 = now_ms();
  std::lock_guard<std::mutex> lock(ctx.mu);
  TextJob job;
  format_text(job);
  std::vector<uint64_t> mine(hdr_.names.size());
  for (size_t r = 0; r < mine.size(); r++) mine[r] = job.span[r].second - job.span[r].first;
  place_pieces(job, mine.data(), 1, 0);
  out.assign(job.nbytes, '\0');
  char *dst = out.empty() ? nullptr : &out[0];
  drain_text(ctx, job.d_text.as<uint8_t>(), job.nbytes, job.pieces, 4,
             [dst](const uint8_t *src, uint64_t len, uint64_t off) { memcpy(dst + off, src, len); });
  tm_.text_ms = now_ms() - t0;
}

// Fills this shard's pieces into the output file `fd` (size `total`; the first `prealloc` bytes already have pages) and
// closes it.  Regular files: the writer threads memcpy into a shared mapping (parallel page-cache fills; pwrite would
// serialise on the inode lock); anything unmappable falls back to positional writes.
void Pileup::write_pieces(TextJob &job, int fd, const std::string &path, uint64_t total, uint64_t prealloc, bool truncate_at_end, char *premap,
                          uint64_t premap_len, bool sparse_ok) {
  DeviceCtx &ctx = *im_->ctx;
  char *map = nullptr;
  uint64_t map_len = total;
  const double t_w0 = now_ms();
  double t_backed = t_w0, t_mapped = t_w0;
  try {
    // every byte the writers will touch must be backed before it is mapped: a store into a hole of a sparse file raises
    // SIGBUS when the file system is full, which would take the host process down instead of returning an error
    bool backed = !job.pieces.empty();
    // Sharded jobs: every rank allocating its own pieces queues on the file's inode lock (tmpfs: 10-14 GB/s for all ranks
    // together, ~1 us per page even where pages exist), while pages allocated by the writers' own page faults scale with the
    // ranks (27 GB/s over eight ranks).  When the file system has room for twice the file, the holes are left to the faults;
    // a file system that is nearly full takes the slow, checked path.
    bool roomy = false;
    if (sparse_ok) {
      struct statvfs vfs;
      if (fstatvfs(fd, &vfs) == 0) roomy = (double)vfs.f_bavail * (double)vfs.f_frsize >= 2.0 * (double)total;
    }
    for (const TextPiece &pc : job.pieces) {
      if (roomy) break;
      const uint64_t a0 = pc.out_off, a1 = pc.out_off + (pc.dev_end - pc.dev_begin);
      if (a1 > prealloc && posix_fallocate(fd, (off_t)std::max(a0, prealloc), (off_t)(a1 - std::max(a0, prealloc))) != 0) backed = false;
    }
    t_backed = now_ms();
    if (backed && total && premap && total <= premap_len) {
      map = premap;  // the mapping whose pages were touched while the pile-up ran
      map_len = premap_len;
      premap = nullptr;
    } else if (backed && total) {
      void *m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (m != MAP_FAILED) map = (char *)m;
    }
    if (premap) {
      munmap(premap, premap_len);
      premap = nullptr;
    }
    // the writers spend most of their time in page faults, so more of them than cores pays: 8 -> 24 threads on the
    // 16-core B200 hosts takes the drain of 2.1 GB from 155 to 125 ms (profiles/r1_sweeps.md)
    const unsigned hw = host_cores_share();
    const int writers = (int)env_u64("BND_WRITERS", std::min(32u, std::max(4u, hw + hw / 2)));
    t_mapped = now_ms();
    const bool nt = env_u64("BND_WRITE_NT", 1) != 0;  // the text is not read again by this process: keep it out of the caches
    if (map)
      drain_text(ctx, job.d_text.as<uint8_t>(), job.nbytes, job.pieces, writers, [map, nt](const uint8_t *src, uint64_t len, uint64_t off) {
        if (nt && len >= 4096) stream_copy((uint8_t *)map + off, src, len);
        else memcpy(map + off, src, len);
      });
    else
      drain_text(ctx, job.d_text.as<uint8_t>(), job.nbytes, job.pieces, writers, [fd](const uint8_t *src, uint64_t len, uint64_t off) {
        uint64_t done = 0;
        while (done < len) {
          ssize_t w = pwrite(fd, src + done, len - done, (off_t)(off + done));
          if (w <= 0) fail("write error on the bedGraph output");
          done += (uint64_t)w;
        }
      });
  } catch (...) {
    if (map) munmap(map, map_len);
    if (premap) munmap(premap, premap_len);
    close(fd);
    throw;
  }
  // The file's pages are complete; tearing down the page-table entries of the mapping (26 ms for 2 GB) only concerns this
  // process's address space, so it is left to a background thread and the file is closed right away.
  const double t_drained = now_ms();
  if (map) g_unmapper.submit(map, map_len);
  if (truncate_at_end && ftruncate(fd, (off_t)total) != 0) {
    close(fd);
    fail("write error on " + path);
  }
  if (close(fd) != 0) fail("write error on " + path);
  if (env_u64("BND_DEBUG_TIMING", 0)) fprintf(stderr, "[bnd] write_pieces: %zu pieces, backing %.2f ms, map %.2f ms, drain %.2f ms, truncate + close %.2f ms\n", job.pieces.size(), t_backed - t_w0,
                                             t_mapped - t_backed, t_drained - t_mapped, now_ms() - t_drained);
}

void Pileup::to_bedgraph(const std::string &path) {
  DeviceCtx &ctx = *im_->ctx;
  const double t00 = now_ms();
  int fd = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) fail("cannot create " + path);
  // When the pile-up still has to run, the output file's pages are allocated in the background meanwhile
  // (one fallocate of an upper bound of the text size: rows <= min(bins, 2 reads + chunks), <= name + 46 bytes each).
  // Filling pre-allocated page-cache pages runs at memory speed (16 GB/s with 8 threads on the B200 hosts) instead of the
  // 6 GB/s of faulting them in one by one; the surplus is cut off by the final ftruncate.
  std::atomic<uint64_t> prealloc_done{0}, touched{0};
  std::atomic<bool> prealloc_stop{false};
  std::thread prealloc_thread;
  std::vector<std::thread> touchers;
  char *premap = nullptr;
  uint64_t premap_len = 0;
  double t_setup = t00;
  auto stop_background = [&] {
    prealloc_stop.store(true);
    if (prealloc_thread.joinable()) prealloc_thread.join();
    for (auto &t : touchers) t.join();
    touchers.clear();
  };
  try {
    if (!accumulated_ && env_u64("BND_PREALLOC", 1)) {
      // 72 % of the upper bound (rows rarely reach it and are shorter than the longest possible: 2.1 GB of 2.9 GB at the
      // BASELINE size), and never more than the pile-up can hide: it moves ~30 GB/s of
      // BAM, fallocate ~14 GB/s of pages.  The thread allocates in 64 MiB steps and stops when the pile-up is done.
      uint64_t want = estimate_text_bytes(false);
      want = std::min<uint64_t>(want, (shard_.byte_end - shard_.byte_begin) * 6 / 10);
      want &= ~(uint64_t)4095;
      if (want > env_u64("BND_PREALLOC_MIN", 8ull << 20)) {
        // The same range is mapped now and its pages are touched behind the allocation (one store per page: the page-table entry
        // exists and is writable when the text arrives), so that the writers later run at copy speed instead of taking one page
        // fault per 4 KiB.  Plain stores take the per-VMA lock only; MADV_POPULATE_WRITE held the process's mmap lock for the
        // whole range and stalled every CUDA driver call behind it (profiles/r1_sweeps.md).
        if (env_u64("BND_PRETOUCH", 1)) {
          void *m = mmap(nullptr, want, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
          if (m != MAP_FAILED) premap = (char *)m, premap_len = want;
        }
        prealloc_thread = std::thread([fd, want, &prealloc_done, &prealloc_stop] {
          const uint64_t step = 64ull << 20;
          uint64_t off = 0;
          while (off < want && !prealloc_stop.load(std::memory_order_relaxed)) {
            const uint64_t n = std::min(step, want - off);
            if (posix_fallocate(fd, (off_t)off, (off_t)n) != 0) break;
            off += n;
            prealloc_done.store(off, std::memory_order_release);
          }
          prealloc_done.store(off);
        });
        if (premap) {
          const int n_touch = (int)env_u64("BND_TOUCHERS", std::max(1u, std::min(4u, host_cores_share() / 4)));  // 2 / 4 threads touch 1.4 / 2.1 GB of pages behind a 300 ms pile-up
          for (int t = 0; t < n_touch; t++)
            touchers.emplace_back([premap, want, t, n_touch, &prealloc_done, &prealloc_stop, &touched] {
              const uint64_t step = 8ull << 20;  // touchers interleave 8 MiB blocks behind the allocation front
              for (uint64_t blk = (uint64_t)t; blk * step < want; blk += (uint64_t)n_touch) {
                const uint64_t b0 = blk * step, b1 = std::min(want, b0 + step);
                while (prealloc_done.load(std::memory_order_acquire) < b1) {
                  if (prealloc_stop.load(std::memory_order_relaxed)) return;
                  std::this_thread::sleep_for(std::chrono::microseconds(200));
                }
                if (prealloc_stop.load(std::memory_order_relaxed)) return;
                for (uint64_t p = b0; p < b1; p += 4096) ((volatile char *)premap)[p] = 0;
                touched.fetch_add(b1 - b0, std::memory_order_relaxed);
              }
            });
        }
      }
    }
    t_setup = now_ms();
    ensure_finalized();
  } catch (...) {
    stop_background();
    if (premap) munmap(premap, premap_len);
    close(fd);
    throw;
  }
  const double t_fin = now_ms();
  stop_background();
  const double t_bg = now_ms();
  g_unmapper.reap();  // the previous job's output mapping was torn down while this pile-up ran
  const uint64_t prealloc = prealloc_done.load();
  const double t0 = now_ms();
  if (env_u64("BND_DEBUG_TIMING", 0))
    fprintf(stderr, "[bnd] to_bedgraph: open + background start %.2f ms, accumulate + finalize %.2f ms, background stop %.2f ms, unmap reap %.2f ms\n", t_setup - t00,
            t_fin - t_setup, t_bg - t_fin, t0 - t_bg);
  std::lock_guard<std::mutex> lock(ctx.mu);
  TextJob job;
  uint64_t out_bytes = 0;
  double t1 = t0;
  try {
    format_text(job);
    t1 = now_ms();
    std::vector<uint64_t> mine(hdr_.names.size());
    for (size_t r = 0; r < mine.size(); r++) mine[r] = job.span[r].second - job.span[r].first;
    out_bytes = place_pieces(job, mine.data(), 1, 0);
  } catch (...) {
    if (premap) munmap(premap, premap_len);
    close(fd);
    throw;
  }
  write_pieces(job, fd, path, out_bytes, prealloc, true, premap, premap_len);
  tm_.text_ms = now_ms() - t0;
  if (env_u64("BND_DEBUG_TIMING", 0))
    fprintf(stderr, "[bnd] to_bedgraph: pile-up %.2f ms (prealloc %llu MB, %llu MB touched), format %.2f ms, drain + close %.2f ms, %llu bytes\n", t0 - t00,
            (unsigned long long)(prealloc >> 20), (unsigned long long)(touched.load() >> 20), t1 - t0, now_ms() - t1, (unsigned long long)job.nbytes);
}

// ---- to_bigwig (coverage_analysis.rs:642-759) ------------------------------------------------------------------------------------
// Rows: optional scaffold drop, BAM header order, start-1 / end-1, score = (count as f64 * factor) as f32.  Everything that
// touches the rows runs on the device (bigwig.cuh): items, 1 024-item sections, one zlib stream per section, the summary
// statistics and the zoom levels; the host lays the file out (header, chromosome tree, R-trees) and the section blobs leave
// through the same pinned bounce buffers and writer threads as bedGraph text.
namespace {
struct BwLevel {  // compressed sections of the data level or of one zoom level, back to back on the device
  DevBuf blob;
  uint64_t bytes = 0, n_units = 0;
  uint32_t reduction = 0;
  std::vector<bwl::Section> secs;  // chrom, start, end, offset (inside the blob), size
  ~BwLevel() { blob.release(); }
};
// Everything behind the items: 1 024-item sections, one zlib stream per section, summary statistics, zoom levels, file layout
// and the drain of the section blobs into the file.  Shared by BamPileup::to_bigwig and the BigWig-to-BigWig tools
// (bigwig_compare.cpp).  `chroms`: (name, dense id, length) in id order; chrom_item_first[c] = first item of chromosome c.
void encode_bigwig_file(DeviceCtx &ctx, cudaStream_t st, const std::string &path, const std::vector<std::tuple<std::string, uint32_t, uint32_t>> &chroms,
                        const std::vector<uint64_t> &chrom_item_first, DevBuf &d_items, double t0) {
  const uint32_t n_chrom = (uint32_t)chroms.size();
  const uint64_t n_items = chrom_item_first[n_chrom];
  const bool compress = env_u64("BND_BIGWIG_COMPRESS", 1) != 0;
  DevBuf d_secs, d_staging, d_sizes, d_bounds, d_stats, d_offsets;
  DevBuf z_item_first, z_win_first, z_win_first2, z_recs, z_recs2, z_flags, z_flags2, z_tiles, z_total, z_out, z_chrom_first;
  auto release = [&] {
    for (DevBuf *b : {&d_secs, &d_staging, &d_sizes, &d_bounds, &d_stats, &d_offsets, &z_item_first, &z_win_first, &z_win_first2, &z_recs, &z_recs2, &z_flags,
                      &z_flags2, &z_tiles, &z_total, &z_out, &z_chrom_first})
      b->release();
  };
  BwLevel data;
  std::vector<std::unique_ptr<BwLevel>> zooms;
  uint64_t covered = 0;
  double mn = 0, mx = 0, sum = 0, sumsq = 0;
  uint32_t max_raw = 0;
  int fd = -1;
  char *map = nullptr;
  uint64_t file_bytes = 0;
  const double t_items = now_ms();
  try {
    // ---- one level = section table -> zlib streams -> packed blob ----
    auto encode = [&](BwLevel &lv, const uint8_t *units, const std::vector<uint64_t> &first_unit_of_chrom, uint32_t unit_bytes, uint32_t header_bytes,
                      const uint32_t lags[4], bool want_stats) {
      std::vector<BwSection> table;
      for (uint32_t c = 0; c < n_chrom; c++)
        for (uint64_t u = first_unit_of_chrom[c]; u < first_unit_of_chrom[c + 1]; u += BW_ITEMS)
          table.push_back(BwSection{u, (uint32_t)std::min<uint64_t>(BW_ITEMS, first_unit_of_chrom[c + 1] - u), c});
      const uint32_t ns = (uint32_t)table.size();
      lv.n_units = first_unit_of_chrom[n_chrom];
      lv.secs.clear();
      lv.bytes = 0;
      if (ns == 0) return;
      const uint32_t raw_max = header_bytes + unit_bytes * BW_ITEMS;
      const uint32_t stride = ((compress ? bw_comp_bound(raw_max) : raw_max) + 15u) & ~15u;
      upload(d_secs, table.data(), table.size());
      d_staging.ensure((size_t)ns * stride + 16);
      d_sizes.ensure((size_t)ns * 4);
      d_bounds.ensure((size_t)ns * 8);
      if (want_stats) d_stats.ensure((size_t)ns * 40);
      BwDeflateArgs da{};
      da.units = units;
      da.sections = d_secs.as<BwSection>();
      da.n_sections = ns;
      da.unit_bytes = unit_bytes;
      da.header_bytes = header_bytes;
      da.stride = stride;
      for (int k = 0; k < 4; k++) da.dist[k] = lags[k];
      da.compress = compress;
      da.staging = d_staging.as<uint8_t>();
      da.comp_size = d_sizes.as<uint32_t>();
      da.bounds = d_bounds.as<uint32_t>();
      da.stats = want_stats ? d_stats.as<double>() : nullptr;
      CU(launch_bw_deflate(da, ctx.sm_count, st));
      std::vector<uint32_t> sizes(ns), bounds((size_t)ns * 2);
      CU(cudaMemcpyAsync(sizes.data(), d_sizes.p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(bounds.data(), d_bounds.p, (size_t)ns * 8, cudaMemcpyDeviceToHost, st));
      std::vector<double> stats;
      if (want_stats) {
        stats.resize((size_t)ns * 5);
        CU(cudaMemcpyAsync(stats.data(), d_stats.p, (size_t)ns * 40, cudaMemcpyDeviceToHost, st));
      }
      CU(cudaStreamSynchronize(st));
      std::vector<uint64_t> offs(ns);
      lv.secs.resize(ns);
      for (uint32_t k = 0; k < ns; k++) {
        offs[k] = lv.bytes;
        lv.secs[k] = bwl::Section{table[k].chrom, bounds[2 * k], bounds[2 * k + 1], lv.bytes, sizes[k]};
        lv.bytes += sizes[k];
        max_raw = std::max(max_raw, header_bytes + table[k].count * unit_bytes);
      }
      if (want_stats)
        for (uint32_t k = 0; k < ns; k++) {  // section order: reproducible totals
          const double *q = &stats[(size_t)k * 5];
          if (k == 0) mn = q[1], mx = q[2];
          covered += (uint64_t)q[0];
          mn = std::min(mn, q[1]), mx = std::max(mx, q[2]);
          sum += q[3], sumsq += q[4];
        }
      upload(d_offsets, offs.data(), offs.size());
      lv.blob.ensure(lv.bytes + 16);
      BwCompactArgs ca{};
      ca.staging = d_staging.as<uint8_t>();
      ca.comp_size = d_sizes.as<uint32_t>();
      ca.offset = d_offsets.as<uint64_t>();
      ca.n_sections = ns;
      ca.stride = stride;
      ca.blob = lv.blob.as<uint8_t>();
      CU(launch_bw_compact(ca, ctx.sm_count, st));
      CU(cudaStreamSynchronize(st));  // d_secs / d_staging are reused by the next level
    };
    const uint32_t item_lags[4] = {4, 8, 12, 24};  // start = previous end (lag 8), fields of the previous items (12, 24), zero / equal halves (4)
    encode(data, d_items.as<uint8_t>(), chrom_item_first, BW_ITEM_BYTES, BW_SEC_HEADER, item_lags, true);
    const double t_data = now_ms();

    // ---- zoom levels: first reduction ~10x the mean item width, then 4x per level, until a level is small.  The first level
    // is summarised from the items, every further one from the dense window slots of the level below. ----
    const double t_zoom0 = now_ms();
    if (n_items) {
      upload(z_item_first, chrom_item_first.data(), chrom_item_first.size());
      uint64_t reduction = std::max<uint64_t>(10, covered / n_items * 10);
      DevBuf *recs_cur = &z_recs, *recs_prev = &z_recs2, *flags_cur = &z_flags, *flags_prev = &z_flags2, *wf_cur = &z_win_first, *wf_prev = &z_win_first2;
      for (uint32_t z = 0; z < bwl::MAX_ZOOM && reduction <= 0x7fffffffull; z++, reduction *= 4) {
        std::vector<uint64_t> win_first(n_chrom + 1, 0);
        for (uint32_t c = 0; c < n_chrom; c++) win_first[c + 1] = win_first[c] + ((uint64_t)std::get<2>(chroms[c]) + reduction - 1) / reduction;
        const uint64_t n_win = win_first[n_chrom];
        const uint64_t n_tiles64 = (n_win + BW_THREADS - 1) / BW_THREADS;
        if (n_tiles64 > 0xffffffffull) fail("too many zoom windows");
        upload(*wf_cur, win_first.data(), win_first.size());
        recs_cur->ensure(std::max<uint64_t>(n_win, 1) * BW_ZOOM_BYTES);
        flags_cur->ensure(n_win + 16);
        z_tiles.ensure(n_tiles64 * 8 + 16);
        z_total.ensure(16);
        z_chrom_first.ensure((n_chrom + 1) * 8);
        BwZoomArgs za{};
        za.items = d_items.as<BwItem>();
        za.item_first = z_item_first.as<uint64_t>();
        za.win_first = wf_cur->as<uint64_t>();
        za.n_chrom = n_chrom;
        za.reduction = (uint32_t)reduction;
        za.n_windows = n_win;
        za.recs = recs_cur->as<uint8_t>();
        za.flags = flags_cur->as<uint8_t>();
        za.tile_count = z_tiles.as<uint64_t>();
        za.n_tiles = (uint32_t)n_tiles64;
        za.chrom_first_rec = z_chrom_first.as<uint64_t>();
        if (z > 0) {
          za.prev_recs = recs_prev->as<uint8_t>();
          za.prev_flags = flags_prev->as<uint8_t>();
          za.prev_win_first = wf_prev->as<uint64_t>();
        }
        CU(launch_bw_zoom_windows(za, ctx.sm_count, st));
        CU(launch_u64_scan(za.tile_count, za.n_tiles, z_total.as<unsigned long long>(), st));
        unsigned long long n_recs = 0;
        CU(cudaMemcpyAsync(&n_recs, z_total.p, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (z > 0 && n_recs < 64) break;  // coarse enough: a viewer reads the previous level in one go
        z_out.ensure(std::max<uint64_t>(n_recs, 1) * BW_ZOOM_BYTES);
        za.out = z_out.as<uint8_t>();
        CU(launch_bw_zoom_compact(za, n_recs, ctx.sm_count, st));
        std::vector<uint64_t> chrom_first_rec(n_chrom + 1);
        CU(cudaMemcpyAsync(chrom_first_rec.data(), z_chrom_first.p, (n_chrom + 1) * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        zooms.emplace_back(new BwLevel());
        BwLevel &lv = *zooms.back();
        lv.reduction = (uint32_t)reduction;
        const uint32_t zoom_lags[4] = {4, 28, 32, 0};  // min = max (4), start = previous end (28), same field of the previous record (32)
        encode(lv, z_out.as<uint8_t>(), chrom_first_rec, BW_ZOOM_BYTES, 0, zoom_lags, false);
        if (n_recs < 64) break;
        std::swap(recs_cur, recs_prev);
        std::swap(flags_cur, flags_prev);
        std::swap(wf_cur, wf_prev);
      }
    }
    const double t_zoom1 = now_ms();

    // ---- file layout ----
    bwl::Buf head;  // header (64) + zoom headers + total summary, then the chromosome tree
    head.pad_to(64 + bwl::MAX_ZOOM * 24 + 40);
    const uint64_t chrom_tree_off = head.d.size();
    bwl::write_chrom_tree(head, chroms);
    const uint64_t data_off = head.d.size();
    head.u64(data.secs.size());
    const uint64_t data_blob_off = head.d.size();
    for (auto &sc : data.secs) sc.offset += data_blob_off;
    const uint64_t index_off = data_blob_off + data.bytes;
    bwl::Buf index;
    bwl::write_rtree(index, data.secs, index_off, index_off);
    struct ZoomPlace {
      uint64_t data_off, blob_off, index_off;
      std::vector<uint8_t> index;
    };
    std::vector<ZoomPlace> zp(zooms.size());
    uint64_t cur = index_off + index.d.size();
    for (size_t z = 0; z < zooms.size(); z++) {
      BwLevel &lv = *zooms[z];
      zp[z].data_off = cur;
      zp[z].blob_off = cur + 4;  // u32 record count in front of a level's sections, as UCSC writers put it
      for (auto &sc : lv.secs) sc.offset += zp[z].blob_off;
      zp[z].index_off = zp[z].blob_off + lv.bytes;
      bwl::Buf tmp;
      bwl::write_rtree(tmp, lv.secs, zp[z].index_off, zp[z].index_off);
      zp[z].index.swap(tmp.d);
      cur = zp[z].index_off + zp[z].index.size();
    }
    file_bytes = cur + 4;  // trailing magic
    {
      bwl::Buf hd;
      hd.u32(0x888FFC26);
      hd.u16(4);
      hd.u16((uint16_t)zooms.size());
      hd.u64(chrom_tree_off);
      hd.u64(data_off);
      hd.u64(index_off);
      hd.u16(0), hd.u16(0);
      hd.u64(0);
      hd.u64(64 + bwl::MAX_ZOOM * 24);  // total summary offset: after the (reserved) zoom headers
      hd.u32(compress ? max_raw : 0);   // uncompressBufSize: size of the largest inflated section; 0 = sections are stored raw
      hd.u64(0);
      for (uint32_t z = 0; z < bwl::MAX_ZOOM; z++) {
        const bool on = z < zooms.size();
        hd.u32(on ? zooms[z]->reduction : 0), hd.u32(0);
        hd.u64(on ? zp[z].data_off : 0), hd.u64(on ? zp[z].index_off : 0);
      }
      hd.u64(covered), hd.f64(mn), hd.f64(mx), hd.f64(sum), hd.f64(sumsq);
      memcpy(head.d.data(), hd.d.data(), hd.d.size());
    }
    fd = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) fail("cannot create " + path);
    if (posix_fallocate(fd, 0, (off_t)file_bytes) != 0) fail("cannot allocate " + path);
    void *m = mmap(nullptr, file_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) fail("cannot map " + path);
    map = (char *)m;
    memcpy(map, head.d.data(), head.d.size());
    memcpy(map + index_off, index.d.data(), index.d.size());
    for (size_t z = 0; z < zooms.size(); z++) {
      const uint32_t n = (uint32_t)std::min<uint64_t>(zooms[z]->n_units, 0xffffffffull);
      memcpy(map + zp[z].data_off, &n, 4);
      memcpy(map + zp[z].index_off, zp[z].index.data(), zp[z].index.size());
    }
    const uint32_t magic = 0x888FFC26;
    memcpy(map + file_bytes - 4, &magic, 4);
    const unsigned hw = host_cores_share();
    const int writers = (int)env_u64("BND_WRITERS", std::min(32u, std::max(4u, hw)));
    auto drain_level = [&](BwLevel &lv, uint64_t off) {
      if (!lv.bytes) return;
      std::vector<TextPiece> one{TextPiece{0, lv.bytes, off}};
      char *base = map;
      drain_text(ctx, lv.blob.as<uint8_t>(), lv.bytes, one, writers, [base](const uint8_t *src, uint64_t len, uint64_t o) { memcpy(base + o, src, len); });
    };
    const double t_layout = now_ms();
    drain_level(data, data_blob_off);
    for (size_t z = 0; z < zooms.size(); z++) drain_level(*zooms[z], zp[z].blob_off);
    if (env_u64("BND_DEBUG_TIMING", 0))
      fprintf(stderr, "[bnd] to_bigwig: items %.2f ms, data sections %.2f ms, zoom levels %.2f ms, layout + file %.2f ms, drain %.2f ms\n", t_items - t0, t_data - t_items,
              t_zoom1 - t_zoom0, t_layout - t_zoom1, now_ms() - t_layout);
  } catch (...) {
    release();
    if (map) munmap(map, file_bytes);
    if (fd >= 0) close(fd);
    throw;
  }
  release();
  g_unmapper.submit(map, file_bytes);
  if (close(fd) != 0) fail("write error on " + path);
  if (env_u64("BND_DEBUG_TIMING", 0))
    fprintf(stderr, "[bnd] to_bigwig: %llu sections, %llu bytes of data sections, %zu zoom levels, %.2f ms after the pile-up\n",
            (unsigned long long)data.secs.size(), (unsigned long long)data.bytes, zooms.size(), now_ms() - t0);
}


}  // namespace

void Pileup::to_bigwig(const std::string &path) {
  ensure_finalized();
  g_unmapper.reap();
  Impl &im = *im_;
  DeviceCtx &ctx = *im.ctx;
  const double t0 = now_ms();
  const double factor = norm_factor();
  std::lock_guard<std::mutex> lock(ctx.mu);
  CU(cudaSetDevice(ctx.device));
  cudaStream_t st = ctx.sets[0].stream;
  const size_t n_ref = hdr_.names.size();

  // chromosome ids are dense (0 .. kept - 1, header order): readers such as libBigWig index arrays of `itemCount`
  // entries by id, so the BAM reference index cannot be used once references are dropped
  std::vector<uint8_t> skip(n_ref, 0);
  std::vector<uint32_t> dense(n_ref, 0);
  std::vector<std::tuple<std::string, uint32_t, uint32_t>> chroms;
  std::vector<int> ref_of_dense;
  for (size_t r = 0; r < n_ref; r++) {
    skip[r] = cfg_.ignore_scaffold && is_scaffold(hdr_.names[r]);
    if (hdr_.lens[r] > 0xffffffffull) fail("chromosome too long for BigWig");
    if (!skip[r]) {
      dense[r] = (uint32_t)chroms.size();
      chroms.emplace_back(hdr_.names[r], dense[r], (uint32_t)hdr_.lens[r]);
      ref_of_dense.push_back((int)r);
    }
  }
  const uint32_t n_chrom = (uint32_t)chroms.size();

  DevBuf d_row_first, d_item_first, d_skip, d_items, d_flag;
  auto release = [&] {
    for (DevBuf *b : {&d_row_first, &d_item_first, &d_skip, &d_items, &d_flag}) b->release();
  };
  try {
    // ---- items of the kept references ----
    std::vector<uint64_t> row_first(n_ref + 1, 0), item_first(n_ref + 1, 0);
    d_row_first.ensure((n_ref + 1) * 8);
    if (im.n_runs) {
      CU(launch_bw_row_first(im.d_run_bin.as<uint64_t>(), im.n_runs, im.gdev, d_row_first.as<uint64_t>(), st));
      CU(cudaMemcpyAsync(row_first.data(), d_row_first.p, (n_ref + 1) * 8, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
    }
    for (size_t r = 0; r < n_ref; r++) item_first[r + 1] = item_first[r] + (skip[r] ? 0 : row_first[r + 1] - row_first[r]);
    const uint64_t n_items = item_first[n_ref];
    std::vector<uint64_t> chrom_item_first(n_chrom + 1, 0);  // per dense chromosome
    for (uint32_t c = 0; c < n_chrom; c++) chrom_item_first[c] = item_first[ref_of_dense[c]];
    chrom_item_first[n_chrom] = n_items;
    upload(d_item_first, item_first.data(), item_first.size());
    upload(d_skip, skip.data(), skip.size());
    d_items.ensure(std::max<uint64_t>(n_items, 1) * sizeof(BwItem));
    d_flag.ensure(16);
    CU(cudaMemsetAsync(d_flag.p, 0, 16, st));
    if (im.n_runs) {
      BwItemArgs ia{};
      ia.run_bin = im.d_run_bin.as<uint64_t>();
      ia.run_val = im.d_run_val.as<uint32_t>();
      ia.n_runs = im.n_runs;
      ia.g = im.gdev;
      ia.row_first = d_row_first.as<uint64_t>();
      ia.item_first = d_item_first.as<uint64_t>();
      ia.ref_skip = d_skip.as<uint8_t>();
      ia.factor = factor;
      ia.items = d_items.as<BwItem>();
      ia.range_error = d_flag.as<unsigned int>();
      CU(launch_bw_items(ia, ctx.sm_count, st));
      unsigned int bad = 0;
      CU(cudaMemcpyAsync(&bad, d_flag.p, 4, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      if (bad) fail("interval out of range for BigWig");
    }
    encode_bigwig_file(ctx, st, path, chroms, chrom_item_first, d_items, t0);
  } catch (...) {
    release();
    throw;
  }
  release();
  tm_.text_ms = now_ms() - t0;
}

// Sharded jobs write ONE bedGraph (write_bedgraph, coverage_analysis.rs:43-75): every rank formats its rows on the device and
// reports the bytes it holds per reference; the caller gathers those tables from all ranks (n_ref x u64 each, the only
// exchange besides the 13 filter counters) and every rank then fills its own pieces of the same file.
// Upper estimate of the bedGraph text of the chunks [lo, hi): rows <= min(bins, 2 reads + chunks), <= name + 46 bytes each,
// scaled by what real outputs reach (2.1 of 2.9 GB at the BASELINE size).
uint64_t Pileup::estimate_text_bytes(bool whole_job) const {
  const uint64_t bins = whole_job ? geo_.first_bin_of_chunk(geo_.total_chunks()) : shard_bins_upper();
  const uint64_t rows_cap = std::min<uint64_t>(bins, 2 * (geo_.n_mapped + geo_.n_unmapped) + geo_.total_chunks() + 16);
  double bytes = 0, bins_total = 0;
  for (size_t r = 0; r < hdr_.names.size(); r++) {
    const double nb = (double)(geo_.ref_first_bin[r + 1] - geo_.ref_first_bin[r]);
    bytes += nb * (double)(hdr_.names[r].size() + 46);
    bins_total += nb;
  }
  const double per_row = bins_total > 0 ? bytes / bins_total : 64.0;
  return (uint64_t)((double)rows_cap * per_row * 0.72) & ~(uint64_t)4095;
}

// Sharded jobs: ONE rank calls this before its pile-up.  The merged file is created and its estimated pages are allocated by a
// background thread while the pile-ups run; without it every rank allocates its own pieces when it writes them, and those
// posix_fallocate calls queue up on the file's inode lock (2.1 GB at 14 GB/s: 150 ms that no rank can hide).
void Pileup::preallocate_output(const std::string &path) {
  Impl &im = *im_;
  if (im.prealloc_thread.joinable()) {
    im.prealloc_stop.store(true);
    im.prealloc_thread.join();
  }
  im.prealloc_stop.store(false);
  const uint64_t want = estimate_text_bytes(true);
  const int fd = ::open(path.c_str(), O_RDWR | O_CREAT, 0644);
  if (fd < 0) fail("cannot create " + path);
  im.prealloc_done.store(0);
  im.prealloc_thread = std::thread([fd, want, stop = &im.prealloc_stop, done = &im.prealloc_done] {
    const uint64_t step = 64ull << 20;
    for (uint64_t off = 0; off < want && !stop->load(std::memory_order_relaxed); off += step) {
      if (posix_fallocate(fd, (off_t)off, (off_t)std::min(step, want - off)) != 0) break;
      done->store(std::min(want, off + step));
    }
    close(fd);
  });
}

// Bytes of the merged output this rank has allocated so far; stops the background allocation (called after the pile-up, before
// the ranks exchange their tables, so that every rank learns how much of the file needs no allocation when it writes).
uint64_t Pileup::preallocated_bytes() {
  Impl &im = *im_;
  if (im.prealloc_thread.joinable()) {
    im.prealloc_stop.store(true);
    im.prealloc_thread.join();
  }
  return im.prealloc_done.load();
}

void Pileup::format_bedgraph(std::vector<uint64_t> &bytes_per_ref) {
  ensure_finalized();
  Impl &im = *im_;
  std::lock_guard<std::mutex> lock(im.ctx->mu);
  im.text_job.reset(new TextJob());
  try {
    format_text(*im.text_job);
  } catch (...) {
    im.text_job.reset();
    throw;
  }
  bytes_per_ref.assign(hdr_.names.size(), 0);
  for (size_t r = 0; r < bytes_per_ref.size(); r++) bytes_per_ref[r] = im.text_job->span[r].second - im.text_job->span[r].first;
}

void Pileup::write_bedgraph_pieces(const std::string &path, const uint64_t *bytes_all, int world, int rank, uint64_t preallocated) {
  Impl &im = *im_;
  if (!im.text_job) fail("write_bedgraph_pieces called before format_bedgraph", ERR_INVALID);
  if (world < 1 || rank < 0 || rank >= world || !bytes_all) fail("bad shard table", ERR_INVALID);
  g_unmapper.reap();
  const double t0 = now_ms();
  std::lock_guard<std::mutex> lock(im.ctx->mu);
  std::unique_ptr<TextJob> job = std::move(im.text_job);
  const uint64_t total = place_pieces(*job, bytes_all, world, rank);
  // every rank opens (or creates) the same file and sets the same size; pieces never overlap, so no rank has to wait for
  // another one before it fills its own
  preallocated_bytes();  // a rank that still allocates in the background stops here (callers normally did that before the exchange)
  int fd = ::open(path.c_str(), O_RDWR | O_CREAT, 0644);
  if (fd < 0) fail("cannot create " + path);
  struct stat st;
  if (fstat(fd, &st) != 0 || ((uint64_t)st.st_size != total && ftruncate(fd, (off_t)total) != 0)) {
    close(fd);
    fail("cannot size " + path);
  }
  const double t_open = now_ms();
  // [0, preallocated) already has pages (allocated behind the pile-ups by the rank that was asked to): only what lies beyond is
  // allocated here, piece by piece - on tmpfs such calls cost ~1 us per page even for existing pages and queue on the inode lock
  write_pieces(*job, fd, path, total, std::min(preallocated, total), false, nullptr, 0, true);
  tm_.text_ms = now_ms() - t0;
  if (env_u64("BND_DEBUG_TIMING", 0)) fprintf(stderr, "[bnd] write_bedgraph_pieces: reap + place + open + size %.2f ms, pieces %.2f ms\n", t_open - t0, now_ms() - t_open);
}

// K-deflate exposed on its own (bnd_bgzf_compress): bytes in host memory -> BGZF blocks + EOF marker, compressed on the device
void device_bgzf_compress(const uint8_t *data, uint64_t n, std::vector<uint8_t> &out, int device) {
  DeviceCtx &ctx = ctx_for(device);
  std::lock_guard<std::mutex> lock(ctx.mu);
  cudaStream_t st = ctx.sets[0].stream;
  BamWriteJob job;
  job.ctx = &ctx;
  job.store_only = env_u64("BND_BAM_STORE_ONLY", 0) != 0;
  ClassStream s;
  s.mem = &out;
  out.clear();
  DevBuf d;
  try {
    if (n) {
      d.ensure(n + 64);
      upload_bytes(ctx, data, n, d.as<uint8_t>(), st);
      job.compress_and_write(s, d.as<uint8_t>(), n, nullptr, st);
    }
  } catch (...) {
    d.release();
    throw;
  }
  d.release();
  out.insert(out.end(), BGZF_EOF, BGZF_EOF + sizeof BGZF_EOF);
}

// ---- split / modify / split-exogenous ----------------------------------------------------------------------------------------
void Pileup::write_bam(int mode, int tn5_shift, const std::string &exogenous_prefix, const std::vector<std::string> &paths, const std::string &command_line,
                       uint64_t *counters, uint64_t *file_bytes) {
  Impl &im = *im_;
  DeviceCtx &ctx = *im.ctx;
  if (cfg_.shard_world > 1) fail("BAM writers run unsharded", ERR_INVALID);
  if (mode != BAMW_SPLIT && mode != BAMW_MODIFY && mode != BAMW_SPLIT_EXOGENOUS) fail("unknown BAM writer mode", ERR_INVALID);
  const double t0 = now_ms();
  BamWriteJob job;
  {
    std::lock_guard<std::mutex> lock(ctx.mu);
    CU(cudaSetDevice(ctx.device));
    job.begin(ctx, hdr_, im.fdev, nullptr, mode, tn5_shift, exogenous_prefix, paths, command_line);
  }
  im.wjob = &job;
  try {
    run_pass(shard_, false);
  } catch (...) {
    im.wjob = nullptr;
    accumulated_ = false;
    throw;
  }
  im.wjob = nullptr;
  accumulated_ = false;  // the pass fed the writers, not the bins
  unsigned long long cnt[BW_N_COUNTERS];
  {
    std::lock_guard<std::mutex> lock(ctx.mu);
    CU(cudaSetDevice(ctx.device));
    CU(cudaMemcpy(cnt, job.d_counts.p, sizeof cnt, cudaMemcpyDeviceToHost));
    if (cnt[BW_N_ERRORS]) {
      // the reference propagates these with `?` and stops (read_filter.rs: missing reference id / malformed aux data;
      // bam_modifier.rs:186 "Missing mate alignment start")
      for (auto &c : job.cls)
        if (c.fd >= 0) close(c.fd), c.fd = -1;
      fail(std::to_string(cnt[BW_N_ERRORS]) + " record(s) made the read filter or the Tn5 shift fail (no reference id, malformed aux data or no mate start)");
    }
    job.flush(true);
  }
  for (int k = 0; k < BW_N_COUNTERS; k++)
    if (counters) counters[k] = cnt[k];
  for (int k = 0; k < BAMW_MAX_CLASSES; k++)
    if (file_bytes) file_bytes[k] = job.cls[k].file_off;
  if (env_u64("BND_DEBUG_TIMING", 0)) {
    uint64_t raw = 0, out = 0, blocks = 0;
    for (auto &c : job.cls) raw += c.raw_bytes, out += c.file_off, blocks += c.blocks;
    fprintf(stderr, "[bnd] write_bam (mode %d): %.2f ms; %llu records kept of %llu, %llu raw bytes -> %llu bytes in %llu BGZF blocks; %llu flushes, compress %.2f ms, D2H + write %.2f ms\n",
            mode, now_ms() - t0, (unsigned long long)(cnt[BW_N_WRITTEN] + cnt[BW_N_ENDOGENOUS] + cnt[BW_N_EXOGENOUS] + cnt[BW_N_UNMAPPED] + cnt[BW_N_QCFAIL] + cnt[BW_N_DUPLICATE] + cnt[BW_N_SECONDARY]),
            cnt[BW_N_TOTAL], (unsigned long long)raw, (unsigned long long)out, (unsigned long long)blocks, (unsigned long long)job.flushes, job.compress_ms, job.write_ms);
  }
}

std::string path_with_extension(const std::string &prefix, const std::string &ext) {
  const size_t slash = prefix.find_last_of('/');
  const size_t name0 = slash == std::string::npos ? 0 : slash + 1;
  const size_t dot = prefix.find_last_of('.');
  std::string stem = prefix;
  if (dot != std::string::npos && dot > name0) stem = prefix.substr(0, dot);  // a leading dot is part of the name
  return stem + "." + ext;
}

void bam_write_tool(const char *bam_path, const bnd_filter_cfg *filter, int mode, int tn5_shift, const char *exogenous_prefix,
                    const std::vector<std::string> &paths, const char *command_line, int device, uint64_t *counters, uint64_t *file_bytes) {
  if (!bam_path) fail("missing BAM path", ERR_INVALID);
  bnd_pileup_cfg cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.bam_path = bam_path;
  cfg.bin_size = 1000;  // the bins are not used; a coarse grid keeps the idle accumulators small
  cfg.scale_factor = 1.0f;
  cfg.collapse = 1;
  cfg.device = device;
  Pileup p(&cfg, filter);
  p.write_bam(mode, tn5_shift, exogenous_prefix ? exogenous_prefix : "", paths, command_line ? command_line : "", counters, file_bytes);
}

#include "bigwig_tools.inl"

}  // namespace bnd
