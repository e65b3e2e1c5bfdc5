// Microbenchmark: what DFMA rate does one B200 SM sustain, as a function of independent
// chains per thread (ILP) and resident warps per SM?  Build: nvcc -gencode
// arch=compute_100a,code=sm_100a -O3 -o dfma_peak dfma_peak.cu ; prints DFMA per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, double a, double b, int iters) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) acc[i] = __fma_rn(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

template <int ILP>
void run(int warps_per_sm, int sms, double mhz) {
  const int threads = 128;
  const int blocks_per_sm = warps_per_sm * 32 / threads;
  const int iters = 20000;
  double* d;
  cudaMalloc(&d, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dfma_kernel<ILP><<<sms * blocks_per_sm, threads>>>(d, 1.0000001, 1e-9, 100);
  cudaEventRecord(e0);
  dfma_kernel<ILP><<<sms * blocks_per_sm, threads>>>(d, 1.0000001, 1e-9, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double dfma = (double)sms * blocks_per_sm * threads * (double)iters * 8 * ILP;
  const double per_clk_sm = dfma / (ms * 1e-3) / (mhz * 1e6) / sms;
  printf("ILP %d warps/SM %2d : %.1f DFMA/clk/SM  (%.2f ms)\n", ILP, warps_per_sm, per_clk_sm, ms);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  printf("%s, %d SMs, %.0f MHz (nominal max; rates assume it)\n", p.name, sms, mhz);
  for (int w : {4, 8, 12, 16, 32}) {
    run<1>(w, sms, mhz);
    run<2>(w, sms, mhz);
    run<4>(w, sms, mhz);
    run<8>(w, sms, mhz);
  }
  return 0;
}
