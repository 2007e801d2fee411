// Micro-benchmark: is cudaHostRegister of pageable / mmap'd ranges fast enough to replace staging copies?
// nvcc -O2 -o /tmp/hostreg_bench tools/hostreg_bench.cu && /tmp/hostreg_bench
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <chrono>
#include <thread>
#include <vector>

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void run(const char* what, char* base, size_t total, size_t chunk, int threads, unsigned flags, char* dev) {
  std::vector<std::thread> th;
  std::vector<int> fails(threads, 0);
  const size_t n = total / chunk;
  double t0 = now();
  for (int t = 0; t < threads; ++t)
    th.emplace_back([&, t] {
      cudaSetDevice(0);
      cudaStream_t s;
      cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
      for (size_t k = t; k < n; k += threads) {
        char* p = base + k * chunk;
        if (cudaHostRegister(p, chunk, flags) != cudaSuccess) { fails[t]++; cudaGetLastError(); continue; }
        cudaMemcpyAsync(dev + k * chunk, p, chunk, cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);
        cudaHostUnregister(p);
      }
      cudaStreamDestroy(s);
    });
  for (auto& x : th) x.join();
  double dt = now() - t0;
  int f = 0;
  for (int x : fails) f += x;
  printf("%-28s chunk %3zu MiB threads %2d: %6.2f GB/s  (register failures %d)\n", what, chunk >> 20, threads, total / dt / 1e9, f);
}

int main() {
  const size_t total = 1ull << 30;
  char* dev;
  cudaMalloc(&dev, total);
  // pageable heap
  char* heap = (char*)aligned_alloc(4096, total);
  memset(heap, 1, total);
  // file-backed read-only mmap
  const char* path = "/tmp/hostreg_bench.bin";
  int fd = open(path, O_CREAT | O_RDWR | O_TRUNC, 0600);
  for (size_t off = 0; off < total; off += (64u << 20)) { if (write(fd, heap, 64u << 20) < 0) return 1; }
  fsync(fd);
  char* mm = (char*)mmap(nullptr, total, PROT_READ, MAP_SHARED, fd, 0);
  char* mmp = (char*)mmap(nullptr, total, PROT_READ, MAP_PRIVATE, fd, 0);
  volatile char sink = 0;
  for (size_t i = 0; i < total; i += 4096) { sink += mm[i]; sink += mmp[i]; }
  for (size_t chunk : {size_t(8) << 20, size_t(32) << 20, size_t(128) << 20})
    for (int threads : {1, 4, 8}) {
      run("heap, default flags", heap, total, chunk, threads, cudaHostRegisterDefault, dev);
      run("mmap shared RO, ReadOnly", mm, total, chunk, threads, cudaHostRegisterReadOnly, dev);
      run("mmap private RO, ReadOnly", mmp, total, chunk, threads, cudaHostRegisterReadOnly, dev);
    }
  // whole-range registration once
  double t0 = now();
  cudaError_t e = cudaHostRegister(heap, total, cudaHostRegisterDefault);
  double t1 = now();
  printf("register 1 GiB heap at once: %s, %.1f ms (%.2f GB/s)\n", cudaGetErrorString(e), (t1 - t0) * 1e3, total / (t1 - t0) / 1e9);
  if (e == cudaSuccess) { t0 = now(); cudaHostUnregister(heap); printf("unregister: %.1f ms\n", (now() - t0) * 1e3); }
  t0 = now();
  e = cudaHostRegister(mm, total, cudaHostRegisterReadOnly);
  t1 = now();
  printf("register 1 GiB mmap RO at once: %s, %.1f ms (%.2f GB/s)\n", cudaGetErrorString(e), (t1 - t0) * 1e3, total / (t1 - t0) / 1e9);
  unlink(path);
  FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
  char line[128] = "";
  if (f && fgets(line, sizeof line, f)) printf("THP: %s", line);
  return 0;
}
