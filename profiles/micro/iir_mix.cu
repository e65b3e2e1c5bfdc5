// Microbenchmark: the instruction mix of one recursive-Gaussian step -- per chain and sample
// one shared load + float->double conversion, 4 feed-forward + 4 feedback + 1 subtract FP64
// operations with the real dependency structure (feedback chain 5 deep), one 8-byte shared
// store -- as a function of independent chains per thread and resident warps per SM.
// Answers: how close to the FP64 peak can ANY schedule of this mix get at 2-3 warps per SM
// sub-partition?  Prints DP ops per clock per SM (peak 64).
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void __launch_bounds__(128) iir_kernel(const float* __restrict__ in, double* out, int iters,
                                                  double n0, double n1, double n2, double n3, double d1,
                                                  double d2, double d3, double d4) {
  __shared__ float xs[16][128];
  __shared__ double ys[CH][4][128];
  const int t = threadIdx.x;
  for (int i = 0; i < 16; ++i) xs[i][t] = in[i * 128 + t];
  __syncthreads();
  double h0[CH], h1[CH], h2[CH], h3[CH], x1[CH], x2[CH], x3[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) { h0[c] = h1[c] = h2[c] = h3[c] = 0.1 * c; x1[c] = x2[c] = x3[c] = 0.0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const double x0 = (double)xs[(j + c) & 15][t];
        double n = x0 * n0;
        n = __fma_rn(x1[c], n1, n);
        n = __fma_rn(x2[c], n2, n);
        n = __fma_rn(x3[c], n3, n);
        double d = h0[c] * d1;
        d = __fma_rn(h1[c], d2, d);
        d = __fma_rn(h2[c], d3, d);
        d = __fma_rn(h3[c], d4, d);
        const double y = n - d;
        x3[c] = x2[c]; x2[c] = x1[c]; x1[c] = x0;
        h3[c] = h2[c]; h2[c] = h1[c]; h1[c] = h0[c]; h0[c] = y;
        ys[c][j & 3][t] = y;
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += h0[c] + ys[c][3][t];
  if (s == 12345.678) out[0] = s;
}

template <int CH>
void run(int blocks_per_sm, int sms, double mhz, const float* din) {
  const int iters = 400;
  double* d;
  cudaMalloc(&d, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaFuncSetAttribute(iir_kernel<CH>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  iir_kernel<CH><<<sms * blocks_per_sm, 128>>>(din, d, 4, .5, .1, .05, .01, -.9, .3, -.05, .002);
  cudaEventRecord(e0);
  iir_kernel<CH><<<sms * blocks_per_sm, 128>>>(din, d, iters, .5, .1, .05, .01, -.9, .3, -.05, .002);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double dp = (double)sms * blocks_per_sm * 128 * (double)iters * 16 * CH * 9;
  printf("chains/thread %d  warps/SM %2d : %.1f DP/clk/SM (%.0f %% of 64)  %s\n", CH, blocks_per_sm * 4,
         dp / (ms * 1e-3) / (mhz * 1e6) / sms, 100 * dp / (ms * 1e-3) / (mhz * 1e6) / sms / 64,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* din;
  cudaMalloc(&din, 64 * 128 * 4);
  cudaMemset(din, 0, 64 * 128 * 4);
  for (int b : {1, 2, 3, 4, 5, 6, 8}) {
    run<1>(b, p.multiProcessorCount, khz / 1000.0, din);
    run<2>(b, p.multiProcessorCount, khz / 1000.0, din);
    run<4>(b, p.multiProcessorCount, khz / 1000.0, din);
  }
  return 0;
}
