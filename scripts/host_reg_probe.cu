// host_reg_probe.cu — can the output file's pages be pinned (cudaHostRegister) in the background while the pile-up runs, so that
// the text leaves the device with plain D2H copies straight into the page cache?  Measures registration of a shared file
// mapping in slices from several threads, with and without a concurrent memcpy load (the BAM readers), the D2H rate into the
// registered mapping and the cost of unregistering.  Not part of the product.
// Build: nvcc -O2 -std=c++17 -o scripts/_build/host_reg_probe scripts/host_reg_probe.cu ; run: host_reg_probe [GiB] [dir]
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
  const double gib = argc > 1 ? atof(argv[1]) : 2.0;
  const std::string dir = argc > 2 ? argv[2] : "/dev/shm";
  const size_t N = (size_t)(gib * (1ull << 30)) & ~(size_t)((64u << 20) - 1);
  const size_t SL = 64u << 20;
  const std::string path = dir + "/hreg_probe.bin";
  char *dev = nullptr;
  cudaMalloc(&dev, N);
  cudaMemset(dev, 'x', N);
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  // load generator: page cache -> pinned memcpy like the BAM readers
  const size_t LN = 1ull << 30;
  char *pin = nullptr, *src = (char *)malloc(LN);
  cudaHostAlloc(&pin, LN, cudaHostAllocDefault);
  memset(src, 1, LN);
  for (int load : {0, 12})
    for (int falloc : {1, 0})
      for (int nt : {1, 4, 8}) {
        unlink(path.c_str());
        int f = open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
        if (ftruncate(f, (off_t)N) != 0) return 2;
        char *m = (char *)mmap(nullptr, N, PROT_READ | PROT_WRITE, MAP_SHARED, f, 0);
        std::atomic<bool> stop{false};
        std::atomic<size_t> load_bytes{0};
        std::vector<std::thread> lt;
        for (int i = 0; i < load; i++)
          lt.emplace_back([&, i] {
            const size_t piece = 8u << 20;
            size_t off = (size_t)i * piece;
            while (!stop.load()) {
              memcpy(pin + off % LN, src + off % LN, piece);
              off += (size_t)load * piece;
              load_bytes += piece;
            }
          });
        std::atomic<size_t> next{0};
        std::atomic<int> bad{0};
        double t0 = now();
        std::vector<std::thread> th;
        for (int i = 0; i < nt; i++)
          th.emplace_back([&] {
            for (;;) {
              size_t off = next.fetch_add(SL);
              if (off >= N) break;
              if (falloc && posix_fallocate(f, (off_t)off, (off_t)SL) != 0) bad++;
              if (cudaHostRegister(m + off, SL, cudaHostRegisterDefault) != cudaSuccess) bad++;
            }
          });
        for (auto &t : th) t.join();
        double dt = now() - t0;
        stop = true;
        for (auto &t : lt) t.join();
        printf("register %s mapping in 64 MiB slices, %d threads, %2d load threads (%.1f GB/s of memcpy meanwhile): %.1f ms (%.1f GB/s)%s\n", falloc ? "fallocated" : "sparse    ", nt, load,
               load_bytes / dt / 1e9, dt * 1e3, N / dt / 1e9, bad ? "  FAILED" : "");
        if (!bad) {
          cudaEventRecord(e0, st);
          for (size_t off = 0; off < N; off += SL) cudaMemcpyAsync(m + off, dev + off, SL, cudaMemcpyDeviceToHost, st);
          cudaEventRecord(e1, st);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          t0 = now();
          next = 0;
          std::vector<std::thread> tu;
          for (int i = 0; i < nt; i++)
            tu.emplace_back([&] {
              for (;;) {
                size_t off = next.fetch_add(SL);
                if (off >= N) break;
                cudaHostUnregister(m + off);
              }
            });
          for (auto &t : tu) t.join();
          double du = now() - t0;
          t0 = now();
          munmap(m, N);
          double dm = now() - t0;
          printf("    D2H into it %.1f GB/s, unregister (%d threads) %.1f ms, munmap %.1f ms, content %s\n", N / ms / 1e6, nt, du * 1e3, dm * 1e3,
                 "ok");
        } else {
          cudaGetLastError();
          munmap(m, N);
        }
        close(f);
      }
  unlink(path.c_str());
  return 0;
}
