// host_copy_probe.cu — which host copy moves page-cache bytes into pinned staging (BAM readers) and pinned text into the output
// file's page cache (bedGraph writers) fastest on a GPU box: pread, glibc memcpy out of / into a shared mapping, or an AVX2 loop
// with non-temporal stores (no read-for-ownership of the destination lines)?  Not part of the product.
// Build: nvcc -O2 -std=c++17 -Xcompiler -mavx2 -o scripts/_build/host_copy_probe scripts/host_copy_probe.cu ; run: host_copy_probe [GiB] [dir]
#include <cuda_runtime.h>
#include <fcntl.h>
#include <immintrin.h>
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

static void nt_copy(char *dst, const char *src, size_t n) {
  while (n && ((uintptr_t)dst & 31)) *dst++ = *src++, n--;
  size_t v = n / 128;
  for (size_t i = 0; i < v; i++) {
    __m256i a = _mm256_loadu_si256((const __m256i *)(src)), b = _mm256_loadu_si256((const __m256i *)(src + 32));
    __m256i c = _mm256_loadu_si256((const __m256i *)(src + 64)), d = _mm256_loadu_si256((const __m256i *)(src + 96));
    _mm256_stream_si256((__m256i *)(dst), a);
    _mm256_stream_si256((__m256i *)(dst + 32), b);
    _mm256_stream_si256((__m256i *)(dst + 64), c);
    _mm256_stream_si256((__m256i *)(dst + 96), d);
    src += 128, dst += 128;
  }
  n -= v * 128;
  _mm_sfence();
  memcpy(dst, src, n);
}

template <class F>
static double par(int n, F f) {
  double t0 = now();
  std::vector<std::thread> th;
  for (int i = 0; i < n; i++) th.emplace_back([&, i] { f(i); });
  for (auto &t : th) t.join();
  return now() - t0;
}

int main(int argc, char **argv) {
  const double gib = argc > 1 ? atof(argv[1]) : 4.0;
  const std::string dir = argc > 2 ? argv[2] : "/dev/shm";
  const size_t N = (size_t)(gib * (1ull << 30)) & ~(size_t)((8u << 20) - 1);
  const std::string in_path = dir + "/hcp_in.bin", out_path = dir + "/hcp_out.bin";
  {
    int fd = open(in_path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    std::vector<char> buf(8 << 20);
    for (size_t i = 0; i < buf.size(); i++) buf[i] = (char)(i * 131 + (i >> 9));
    for (size_t off = 0; off < N; off += buf.size())
      if (pwrite(fd, buf.data(), buf.size(), (off_t)off) != (ssize_t)buf.size()) return 2;
    close(fd);
  }
  char *pin = nullptr;
  if (cudaHostAlloc(&pin, N, cudaHostAllocDefault) != cudaSuccess) return 3;
  memset(pin, 0, N);
  int fd = open(in_path.c_str(), O_RDONLY);
  char *m = (char *)mmap(nullptr, N, PROT_READ, MAP_SHARED, fd, 0);
  const size_t piece = 8u << 20;
  for (int rep = 0; rep < 2; rep++)
    for (int nt : {8, 12, 16, 24, 32}) {
      std::atomic<size_t> next{0};
      double dt = par(nt, [&](int) {
        for (;;) {
          size_t off = next.fetch_add(piece);
          if (off >= N) break;
          size_t got = 0;
          while (got < piece) got += (size_t)pread(fd, pin + off + got, piece - got, (off_t)(off + got));
        }
      });
      next = 0;
      double dm = par(nt, [&](int) {
        for (;;) {
          size_t off = next.fetch_add(piece);
          if (off >= N) break;
          memcpy(pin + off, m + off, piece);
        }
      });
      next = 0;
      double dn = par(nt, [&](int) {
        for (;;) {
          size_t off = next.fetch_add(piece);
          if (off >= N) break;
          nt_copy(pin + off, m + off, piece);
        }
      });
      printf("read leg, %2d threads: pread %.1f GB/s, memcpy from mapping %.1f GB/s, non-temporal copy from mapping %.1f GB/s\n", nt, N / dt / 1e9, N / dm / 1e9,
             N / dn / 1e9);
    }
  bool ok = memcmp(pin, m, N) == 0;
  printf("read leg content %s\n", ok ? "ok" : "WRONG");
  // write leg: pinned -> fallocated output mapping, pieces at odd offsets like bedGraph text
  for (int mode = 0; mode < 2; mode++)
    for (int nt : {16, 24, 32}) {
      unlink(out_path.c_str());
      int fo = open(out_path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
      if (posix_fallocate(fo, 0, (off_t)N) != 0) return 4;
      char *w = (char *)mmap(nullptr, N, PROT_READ | PROT_WRITE, MAP_SHARED, fo, 0);
      const size_t wp = (4u << 20) - 37;  // unaligned piece size
      std::atomic<size_t> next{0};
      double dt = par(nt, [&](int) {
        for (;;) {
          size_t off = next.fetch_add(wp);
          if (off >= N) break;
          size_t len = std::min(wp, N - off);
          if (mode == 0) memcpy(w + off, pin + off, len);
          else nt_copy(w + off, pin + off, len);
        }
      });
      bool good = memcmp(w, pin, N) == 0;
      printf("write leg (%s), %2d threads: %.1f GB/s, content %s\n", mode ? "non-temporal" : "memcpy", nt, N / dt / 1e9, good ? "ok" : "WRONG");
      munmap(w, N);
      close(fo);
    }
  unlink(out_path.c_str());
  unlink(in_path.c_str());
  return 0;
}
