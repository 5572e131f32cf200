// host_io_probe.cu — measures the host-side legs that bound the file -> bedGraph job on a GPU box (not part of the product):
//   read leg   page cache -> device: pread into pinned staging (N threads) vs memcpy from a mapping vs DMA straight out of a
//              cudaHostRegister'ed mapping of the file (time of the registration itself included)
//   write leg  device text -> output file: memcpy into a fresh shared mapping (N threads) with and without pre-allocated /
//              pre-populated pages, pwrite, and D2H straight into a registered mapping
// Build: nvcc -O2 -std=c++17 -o scripts/_build/host_io_probe scripts/host_io_probe.cu ; run: host_io_probe [GiB] [dir]
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CU(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x);                \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class F>
static double par(int n, F f) {
  double t0 = now();
  std::vector<std::thread> th;
  for (int i = 0; i < n; i++) th.emplace_back([&, i] { f(i); });
  for (auto &t : th) t.join();
  return now() - t0;
}

int main(int argc, char **argv) {
  const double gib = argc > 1 ? atof(argv[1]) : 2.0;
  const std::string dir = argc > 2 ? argv[2] : "/dev/shm";
  const size_t N = (size_t)(gib * (1ull << 30)) & ~(size_t)((2u << 20) - 1);
  const int cores = (int)std::thread::hardware_concurrency();
  printf("host cores %d, %.2f GiB in %s\n", cores, N / double(1ull << 30), dir.c_str());
  const std::string in_path = dir + "/hio_probe_in.bin";
  {  // input file, written once (page cache warm afterwards)
    int fd = open(in_path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    std::vector<char> buf(8 << 20);
    for (size_t i = 0; i < buf.size(); i++) buf[i] = (char)(i * 131 + (i >> 9));
    for (size_t off = 0; off < N; off += buf.size()) {
      if (pwrite(fd, buf.data(), buf.size(), (off_t)off) != (ssize_t)buf.size()) return 2;
    }
    close(fd);
  }
  char *dev = nullptr;
  CU(cudaMalloc(&dev, N));
  char *pin = nullptr;
  double t0 = now();
  CU(cudaHostAlloc(&pin, N, cudaHostAllocDefault));
  printf("cudaHostAlloc %.2f GiB: %.1f ms\n", gib, (now() - t0) * 1e3);
  cudaStream_t st;
  CU(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  auto h2d = [&](const char *src, const char *what) {
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
      CU(cudaEventRecord(e0, st));
      CU(cudaMemcpyAsync(dev, src, N, cudaMemcpyHostToDevice, st));
      CU(cudaEventRecord(e1, st));
      CU(cudaEventSynchronize(e1));
      float ms;
      CU(cudaEventElapsedTime(&ms, e0, e1));
      best = std::min(best, ms);
    }
    printf("H2D from %-28s %.1f GB/s\n", what, N / best / 1e6);
  };
  int fd = open(in_path.c_str(), O_RDONLY);
  // ---- read leg ----
  for (int nt : {1, 4, 8, 12, 16, 24, 32}) {
    if (nt > 2 * cores) break;
    const size_t piece = 8u << 20;
    std::atomic<size_t> next{0};
    double dt = par(nt, [&](int) {
      for (;;) {
        size_t off = next.fetch_add(piece);
        if (off >= N) break;
        size_t len = std::min(piece, N - off), got = 0;
        while (got < len) got += (size_t)pread(fd, pin + off + got, len - got, (off_t)(off + got));
      }
    });
    printf("pread -> pinned, %2d threads: %.1f GB/s\n", nt, N / dt / 1e9);
  }
  h2d(pin, "cudaHostAlloc memory");
  {
    t0 = now();
    char *m = (char *)mmap(nullptr, N, PROT_READ, MAP_SHARED, fd, 0);
    printf("mmap(input) %.2f ms\n", (now() - t0) * 1e3);
    for (int nt : {8, 16}) {
      const size_t piece = 8u << 20;
      std::atomic<size_t> next{0};
      double dt = par(nt, [&](int) {
        for (;;) {
          size_t off = next.fetch_add(piece);
          if (off >= N) break;
          memcpy(pin + off, m + off, std::min(piece, N - off));
        }
      });
      printf("memcpy mapping -> pinned, %2d threads (%s): %.1f GB/s\n", nt, nt == 8 ? "first touch of the mapping" : "mapped", N / dt / 1e9);
    }
    munmap(m, N);
    for (unsigned flags : {(unsigned)cudaHostRegisterReadOnly, (unsigned)cudaHostRegisterDefault}) {
      m = (char *)mmap(nullptr, N, PROT_READ, MAP_SHARED, fd, 0);
      t0 = now();
      cudaError_t e = cudaHostRegister(m, N, flags);
      double dt = now() - t0;
      printf("cudaHostRegister(read-only mapping, flags %u): %s, %.1f ms (%.1f GB/s)\n", flags, cudaGetErrorString(e), dt * 1e3, N / dt / 1e9);
      if (e == cudaSuccess) {
        h2d(m, "registered page-cache mapping");
        t0 = now();
        CU(cudaHostUnregister(m));
        printf("cudaHostUnregister: %.1f ms\n", (now() - t0) * 1e3);
        // in pieces of 64 MiB by 4 threads (what a pipelined reader would do)
        const size_t piece = 64u << 20;
        std::atomic<size_t> next{0};
        std::atomic<int> bad{0};
        double dtp = par(4, [&](int) {
          cudaSetDevice(0);
          for (;;) {
            size_t off = next.fetch_add(piece);
            if (off >= N) break;
            if (cudaHostRegister(m + off, std::min(piece, N - off), flags) != cudaSuccess) bad++;
          }
        });
        printf("  same in 64 MiB pieces by 4 threads: %.1f ms (%.1f GB/s), %d failures\n", dtp * 1e3, N / dtp / 1e9, bad.load());
        for (size_t off = 0; off < N; off += piece) cudaHostUnregister(m + off);
      } else {
        cudaGetLastError();
      }
      munmap(m, N);
    }
  }
  close(fd);
  // ---- write leg ----
  CU(cudaMemset(dev, 'x', N));
  const std::string out_path = dir + "/hio_probe_out.bin";
  auto fresh = [&](bool falloc) {
    unlink(out_path.c_str());
    int f = open(out_path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (falloc) {
      double t = now();
      if (posix_fallocate(f, 0, (off_t)N) != 0) printf("fallocate failed\n");
      printf("  posix_fallocate: %.1f ms (%.1f GB/s)\n", (now() - t) * 1e3, N / (now() - t) / 1e9);
    } else if (ftruncate(f, (off_t)N) != 0) {
      printf("ftruncate failed\n");
    }
    return f;
  };
  for (int falloc = 0; falloc < 2; falloc++)
    for (int nt : {8, 16, 24, 32}) {
      int f = fresh(falloc);
      char *m = (char *)mmap(nullptr, N, PROT_READ | PROT_WRITE, MAP_SHARED, f, 0);
      const size_t piece = 4u << 20;
      std::atomic<size_t> next{0};
      double dt = par(nt, [&](int) {
        for (;;) {
          size_t off = next.fetch_add(piece);
          if (off >= N) break;
          memcpy(m + off, pin + off, std::min(piece, N - off));
        }
      });
      printf("memcpy pinned -> output mapping (%s), %2d threads: %.1f GB/s\n", falloc ? "fallocated" : "sparse", nt, N / dt / 1e9);
      munmap(m, N);
      close(f);
    }
  for (int nt : {8, 24}) {  // per-task windows mapped with MAP_POPULATE (page tables filled in one go, no per-page faults)
    int f = fresh(true);
    const size_t piece = 4u << 20;
    std::atomic<size_t> next{0};
    double dt = par(nt, [&](int) {
      for (;;) {
        size_t off = next.fetch_add(piece);
        if (off >= N) break;
        size_t len = std::min(piece, N - off);
        char *w = (char *)mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, f, (off_t)off);
        memcpy(w, pin + off, len);
        munmap(w, len);
      }
    });
    printf("per-task MAP_POPULATE windows (fallocated), %2d threads: %.1f GB/s\n", nt, N / dt / 1e9);
    close(f);
  }
  for (int nt : {1, 8, 24}) {
    int f = fresh(true);
    const size_t piece = 4u << 20;
    std::atomic<size_t> next{0};
    double dt = par(nt, [&](int) {
      for (;;) {
        size_t off = next.fetch_add(piece);
        if (off >= N) break;
        size_t len = std::min(piece, N - off), done = 0;
        while (done < len) done += (size_t)pwrite(f, pin + off + done, len - done, (off_t)(off + done));
      }
    });
    printf("pwrite pinned -> output (fallocated), %2d threads: %.1f GB/s\n", nt, N / dt / 1e9);
    close(f);
  }
  {  // D2H straight into the registered output mapping
    int f = fresh(true);
    char *m = (char *)mmap(nullptr, N, PROT_READ | PROT_WRITE, MAP_SHARED, f, 0);
    t0 = now();
    cudaError_t e = cudaHostRegister(m, N, cudaHostRegisterDefault);
    double dt = now() - t0;
    printf("cudaHostRegister(output mapping, fallocated): %s, %.1f ms (%.1f GB/s)\n", cudaGetErrorString(e), dt * 1e3, N / dt / 1e9);
    if (e == cudaSuccess) {
      CU(cudaEventRecord(e0, st));
      CU(cudaMemcpyAsync(m, dev, N, cudaMemcpyDeviceToHost, st));
      CU(cudaEventRecord(e1, st));
      CU(cudaEventSynchronize(e1));
      float ms;
      CU(cudaEventElapsedTime(&ms, e0, e1));
      printf("D2H into the registered output mapping: %.1f GB/s\n", N / ms / 1e6);
      t0 = now();
      CU(cudaHostUnregister(m));
      printf("cudaHostUnregister: %.1f ms\n", (now() - t0) * 1e3);
      bool ok = m[12345] == 'x' && m[N - 1] == 'x';
      printf("file content %s\n", ok ? "ok" : "WRONG");
    } else {
      cudaGetLastError();
    }
    munmap(m, N);
    close(f);
  }
  {
    CU(cudaEventRecord(e0, st));
    CU(cudaMemcpyAsync(pin, dev, N, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(e1, st));
    CU(cudaEventSynchronize(e1));
    float ms;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    printf("D2H into cudaHostAlloc memory: %.1f GB/s\n", N / ms / 1e6);
  }
  unlink(out_path.c_str());
  unlink(in_path.c_str());
  return 0;
}
