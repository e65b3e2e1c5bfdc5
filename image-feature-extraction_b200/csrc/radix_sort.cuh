// Device radix sort of float32 keys (ascending, the order std::sort gives on NaN-free data;
// NaNs go to the ends by bit pattern) and stream compaction of per-voxel feature samples.
//
// Used by the step BEFORE binning: equalized bin edges are quantile walks over the sorted
// in-mask feature samples of every (scale, feature) row
// (tools/DetermineHistogramBinEdges_MultiScaleEigenvalueFeatures.cxx:171-296,
// include/ife/Statistics/DetermineEdgesForEqualizedHistogram.h:21-139).
//
// Sort: least-significant-digit radix sort, 4-bit digits, 8 passes, each pass
//   count   one block per tile of 2048 keys -> counts[digit][tile]
//   scan    exclusive prefix over counts (digit-major = the order of the output)
//   scatter the same tiles again; within a tile the keys are taken in 8 coalesced rounds of 256,
//           and a key's rank among equal digits is (earlier rounds) + (earlier warps of this
//           round) + (lower lanes of its warp, __match_any_sync) -- stable by construction.
// HBM-bound: 8 x (2 reads + 1 write) x 4 B per key.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ife {

constexpr int kRsThreads = 256, kRsRounds = 8, kRsTile = kRsThreads * kRsRounds, kRsBins = 16;

__device__ __forceinline__ uint32_t rs_key_of(float f) {   // order-preserving float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float rs_float_of(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void rs_to_keys_kernel(uint32_t* data, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    data[i] = rs_key_of(__uint_as_float(data[i]));
}
__global__ void rs_from_keys_kernel(uint32_t* data, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    data[i] = __float_as_uint(rs_float_of(data[i]));
}

__global__ void __launch_bounds__(kRsThreads)
rs_count_kernel(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t* __restrict__ counts, unsigned n_tiles) {
  __shared__ uint32_t s_cnt[kRsBins];
  if (threadIdx.x < kRsBins) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * kRsTile;
#pragma unroll
  for (int r = 0; r < kRsRounds; ++r) {
    const size_t i = base + (size_t)r * kRsThreads + threadIdx.x;
    const bool on = i < n;
    const unsigned bin = on ? (keys[i] >> shift) & (kRsBins - 1) : 0u;
    const unsigned act = __ballot_sync(0xffffffffu, on);
    if (on) {
      const unsigned peers = __match_any_sync(act, bin);
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_cnt[bin], (uint32_t)__popc(peers));
    }
  }
  __syncthreads();
  if (threadIdx.x < kRsBins) counts[(size_t)threadIdx.x * n_tiles + blockIdx.x] = s_cnt[threadIdx.x];
}

// exclusive prefix sum of `total` counters, one block (the array is a few hundred thousand entries)
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t* __restrict__ counts, size_t total) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (size_t i0 = 0; i0 < total; i0 += 1024) {
    const size_t i = i0 + threadIdx.x;
    const uint32_t v = i < total ? counts[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const uint32_t before = s_carry + (warp ? s_warp[warp - 1] : 0u) + x - v;
    if (i < total) counts[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = before + v;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, int shift,
                  const uint32_t* __restrict__ offsets, unsigned n_tiles) {
  constexpr int W = kRsThreads / 32;
  __shared__ uint32_t s_base[kRsBins];        // where the next key of each digit goes
  __shared__ uint32_t s_warp[W][kRsBins];     // this round: keys of each digit per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < kRsBins) s_base[threadIdx.x] = offsets[(size_t)threadIdx.x * n_tiles + blockIdx.x];
  const size_t base = (size_t)blockIdx.x * kRsTile;
  for (int r = 0; r < kRsRounds; ++r) {
    if (threadIdx.x < W * kRsBins) (&s_warp[0][0])[threadIdx.x] = 0;
    __syncthreads();
    const size_t i = base + (size_t)r * kRsThreads + threadIdx.x;
    const bool on = i < n;
    const uint32_t key = on ? in[i] : 0u;
    const unsigned bin = (key >> shift) & (kRsBins - 1);
    const unsigned act = __ballot_sync(0xffffffffu, on);
    unsigned rank = 0;
    if (on) {
      const unsigned peers = __match_any_sync(act, bin);
      rank = __popc(peers & ((1u << lane) - 1u));
      if (lane == __ffs(peers) - 1) s_warp[warp][bin] = (uint32_t)__popc(peers);
    }
    __syncthreads();
    if (on) {
      uint32_t pos = s_base[bin] + rank;
      for (int w = 0; w < warp; ++w) pos += s_warp[w][bin];
      out[pos] = key;
    }
    __syncthreads();
    if (threadIdx.x < kRsBins) {
      uint32_t t = 0;
#pragma unroll
      for (int w = 0; w < W; ++w) t += s_warp[w][threadIdx.x];
      s_base[threadIdx.x] += t;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Compaction sink: the values of `n_planes` feature planes at the selected voxels, in voxel
// order, as dense rows out[k][j].  Two kernels around a host-side prefix over the tiles'
// counts (a few thousand integers): tile counts, then a gather that ranks the selected voxels
// of a tile with ballots.
// ---------------------------------------------------------------------------------------
constexpr int kCsThreads = 256, kCsTile = kCsThreads * 16;

__global__ void __launch_bounds__(kCsThreads)
cs_count_kernel(const uint8_t* __restrict__ select, size_t n, uint32_t* __restrict__ tile_counts) {
  __shared__ uint32_t s_total;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * kCsTile;
  uint32_t c = 0;
  for (int r = 0; r < kCsTile / kCsThreads; ++r) {
    const size_t i = base + (size_t)r * kCsThreads + threadIdx.x;
    c += (i < n && select[i] != 0) ? 1u : 0u;
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_total, c);
  __syncthreads();
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = s_total;
}

__global__ void __launch_bounds__(kCsThreads)
cs_gather_kernel(const uint8_t* __restrict__ select, size_t n, const uint64_t* __restrict__ tile_offsets,
                 const float* __restrict__ planes, size_t plane_stride, int n_planes, float* __restrict__ out,
                 size_t out_stride) {
  constexpr int W = kCsThreads / 32;
  __shared__ uint32_t s_warp[W];
  __shared__ uint64_t s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = tile_offsets[blockIdx.x];
  const size_t base = (size_t)blockIdx.x * kCsTile;
  for (int r = 0; r < kCsTile / kCsThreads; ++r) {
    const size_t i = base + (size_t)r * kCsThreads + threadIdx.x;
    const bool on = i < n && select[i] != 0;
    const unsigned b = __ballot_sync(0xffffffffu, on);
    if (lane == 0) s_warp[warp] = (uint32_t)__popc(b);
    __syncthreads();
    if (on) {
      uint64_t pos = s_base + __popc(b & ((1u << lane) - 1u));
      for (int w = 0; w < warp; ++w) pos += s_warp[w];
      for (int k = 0; k < n_planes; ++k) out[(size_t)k * out_stride + pos] = planes[(size_t)k * plane_stride + i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
#pragma unroll
      for (int w = 0; w < W; ++w) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
}

// the same rows at an explicit list of voxel indices (random sampling, repeats allowed)
__global__ void cs_gather_index_kernel(const long long* __restrict__ index, size_t n_index, const float* __restrict__ planes,
                                       size_t plane_stride, int n_planes, float* __restrict__ out, size_t out_stride) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_index; j += (size_t)gridDim.x * blockDim.x) {
    const size_t i = (size_t)index[j];
    for (int k = 0; k < n_planes; ++k) out[(size_t)k * out_stride + j] = planes[(size_t)k * plane_stride + i];
  }
}

}  // namespace ife
