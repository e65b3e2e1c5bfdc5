#include <cstdio>
#include <cuda_runtime.h>
// dependent-chain latency and throughput of DFMA / DADD / DMUL / F2F on this GPU
template <int CHAINS>
__global__ void k_dfma(double* out, double a, double b, int iters, long long* cyc) {
  double x[CHAINS];
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = __fma_rn(x[c], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_cvt(double* out, const float* in, int iters, long long* cyc) {
  float v = in[threadIdx.x]; double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { acc = (double)v + acc; v = (float)acc; }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; float* in;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
  const int iters = 4096;
  long long h;
#define RUN(CH, BLOCKS, THREADS) { k_dfma<CH><<<BLOCKS, THREADS>>>(out, 1.0000001, 1e-9, iters, cyc); cudaDeviceSynchronize(); \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("chains/thread=%d warps/SM=%d: %.2f cycles per DFMA-per-chain, %.2f cycles per warp-DFMA issued per SMSP\n", CH, THREADS/32, (double)h/iters, (double)h/iters/CH/((THREADS/32+3)/4)); }
  RUN(1, 1, 32) RUN(2, 1, 32) RUN(4, 1, 32) RUN(8, 1, 32) RUN(16, 1, 32)
  RUN(1, 1, 128) RUN(4, 1, 128) RUN(8, 1, 128) RUN(4, 1, 512) RUN(8, 1, 512)
  k_cvt<<<1, 32>>>(out, in, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("F2F.F64.F32 + DADD + F2F.F32.F64 dependent loop: %.2f cycles per iteration\n", (double)h / iters);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
