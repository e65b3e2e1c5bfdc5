// How fast can 4 GB of page-locked host memory be had?  cudaHostAlloc against
// mmap + parallel first touch (+ transparent huge pages) + cudaHostRegister, and the D2H rate into each.
// nvcc -O2 -o pin_alloc pin_alloc.cu -lpthread ; ./pin_alloc [GB]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <sys/mman.h>
#include <cuda_runtime.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void touch(char* p, size_t bytes, int threads) {
  std::vector<std::thread> pool;
  const size_t per = (bytes / threads + 4095) / 4096 * 4096;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([=] {
      const size_t a = (size_t)t * per, b = a + per < bytes ? a + per : bytes;
      for (size_t i = a; i < b; i += 4096) p[i] = 0;
    });
  for (auto& t : pool) t.join();
}

static double d2h_rate(void* host, void* dev, size_t bytes) {
  cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost);
  const double t0 = now();
  cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost);
  return bytes / (now() - t0) * 1e-9;
}

int main(int argc, char** argv) {
  const size_t bytes = (size_t)(argc > 1 ? atof(argv[1]) : 4.0) * (1ull << 30);
  cudaFree(0);
  void* dev = nullptr;
  cudaMalloc(&dev, bytes);
  cudaMemset(dev, 1, bytes);
  cudaDeviceSynchronize();
  const int threads = (int)std::thread::hardware_concurrency();
  {
    double t0 = now();
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    double t1 = now();
    printf("cudaHostAlloc: %.3f s (%s), D2H %.1f GB/s\n", t1 - t0, cudaGetErrorString(e), d2h_rate(p, dev, bytes));
    t0 = now();
    cudaFreeHost(p);
    printf("  cudaFreeHost: %.3f s\n", now() - t0);
  }
  for (int huge = 0; huge < 2; ++huge) {
    double t0 = now();
    char* p = (char*)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (huge) madvise(p, bytes, MADV_HUGEPAGE);
    touch(p, bytes, threads);
    double t1 = now();
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    double t2 = now();
    printf("mmap%s + touch(%d threads): %.3f s, cudaHostRegister: %.3f s (%s), D2H %.1f GB/s\n", huge ? " + MADV_HUGEPAGE" : "",
           threads, t1 - t0, t2 - t1, cudaGetErrorString(e), d2h_rate(p, dev, bytes));
    t0 = now();
    cudaHostUnregister(p);
    munmap(p, bytes);
    printf("  unregister + munmap: %.3f s\n", now() - t0);
  }
  {
    double t0 = now();
    char* p = (char*)malloc(bytes);
    double r = d2h_rate(p, dev, bytes);   // first copy inside faults the pages in
    printf("pageable malloc: first+second D2H took %.3f s total, steady D2H %.1f GB/s\n", now() - t0, r);
    free(p);
  }
  FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
  if (f) { char buf[128] = {0}; fgets(buf, 127, f); printf("THP: %s", buf); fclose(f); }
  return 0;
}
