// Support box of the output mask.
//
// Every consumer of the smoothed volume on the masked paths -- the fused feature kernel behind
// ImageToEmphysemaFeaturesFilter (its eight MaskImageFilters,
// include/ife/Filters/ImageToEmphysemaFeaturesFilter.hxx:44-54) and the histogram loop of
// tools/MakeBag.cxx:448-457 -- only looks at voxels inside the mask, and the stencils of such a
// voxel reach one voxel further in every direction.  So the smoothed volume is only needed
// inside the mask's bounding box grown by one voxel.  The recursion itself has infinite
// support and must still run along complete lines, but lines that miss the box need not run at
// all and, along a line, nothing below the box needs the anticausal sweep and nothing above it
// the causal one.  The extents are reduced on the device (this kernel) and read back once per
// call; the host turns them into pointer offsets, line counts and output ranges of the
// unchanged pass kernels (smooth_volume in ife_cuda.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ife {

constexpr int kBoxBig = 1 << 30;

// raw[2*d] = max(kBoxBig - min coordinate), raw[2*d+1] = max(coordinate + 1) over the non-zero
// voxels; all zero (as cudaMemsetAsync leaves it) = no voxel yet.
// Requires nx % 16 == 0 and a 16-byte aligned mask.  A warp walks rows (y, z); a lane tests 16
// voxels of the row with one 16-byte load.
__global__ void __launch_bounds__(256)
mask_box_kernel(const uint8_t* __restrict__ mask, int nx, int ny, unsigned n_rows, int* __restrict__ raw) {
  const int ppr = nx >> 4;   // 16-byte pieces per row
  int xlo = kBoxBig, xhi = 0, ylo = kBoxBig, yhi = 0, zlo = kBoxBig, zhi = 0;
  const uint4* m16 = reinterpret_cast<const uint4*>(mask);
  const int lane = threadIdx.x & 31;
  const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
  // four rows per trip: the four 16-byte loads of a lane are in flight together
  for (unsigned row0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4u; row0 < n_rows; row0 += n_warps * 4u) {
    for (int pc = lane; pc < ppr; pc += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = row0 + u < n_rows ? __ldg(m16 + (size_t)(row0 + u) * (size_t)ppr + pc) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if ((v[u].x | v[u].y | v[u].z | v[u].w) == 0u) continue;
        const unsigned w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        int first = 16, last = -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // bit 7 of every non-zero byte
          const unsigned t = (w[k] | ((w[k] & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;
          if (t) {
            first = min(first, 4 * k + ((__ffs(t) - 1) >> 3));
            last = max(last, 4 * k + ((31 - __clz(t)) >> 3));
          }
        }
        const unsigned row = row0 + u;
        const int z = (int)(row / (unsigned)ny), y = (int)(row - (unsigned)z * (unsigned)ny);
        xlo = min(xlo, 16 * pc + first); xhi = max(xhi, 16 * pc + last + 1);
        ylo = min(ylo, y); yhi = max(yhi, y + 1);
        zlo = min(zlo, z); zhi = max(zhi, z + 1);
      }
    }
  }
  const unsigned full = 0xffffffffu;
  xlo = __reduce_min_sync(full, xlo); xhi = __reduce_max_sync(full, xhi);
  ylo = __reduce_min_sync(full, ylo); yhi = __reduce_max_sync(full, yhi);
  zlo = __reduce_min_sync(full, zlo); zhi = __reduce_max_sync(full, zhi);
  if ((threadIdx.x & 31) == 0 && xhi > 0) {
    atomicMax(raw + 0, kBoxBig - xlo); atomicMax(raw + 1, xhi);
    atomicMax(raw + 2, kBoxBig - ylo); atomicMax(raw + 3, yhi);
    atomicMax(raw + 4, kBoxBig - zlo); atomicMax(raw + 5, zhi);
  }
}


// Crop / un-crop between the full volume and the box [o, o + c) (x origin and extent multiples
// of 4, rows 16-byte aligned): four voxels per thread.  See smooth_masked in ife_cuda.cu: the
// masked smoothing runs on the cropped copy, because everything outside the mask's bounding
// box is an exact zero for the recursion.
__global__ void __launch_bounds__(256)
crop_box_kernel(const float* __restrict__ img, const uint8_t* __restrict__ mask, float* __restrict__ cimg,
                uint8_t* __restrict__ cmask, int nx, int ny, int ox, int oy, int oz, int cx, int cy,
                long long n4) {
  const int cx4 = cx >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cx4;
    const int x4 = (int)(i - r * cx4), y = (int)(r % cy), z = (int)(r / cy);
    const size_t src = ((size_t)(z + oz) * ny + (size_t)(y + oy)) * nx + (size_t)(ox + 4 * x4);
    reinterpret_cast<float4*>(cimg)[i] = __ldg(reinterpret_cast<const float4*>(img + src));
    reinterpret_cast<uchar4*>(cmask)[i] = __ldg(reinterpret_cast<const uchar4*>(mask + src));
  }
}

__global__ void __launch_bounds__(256)
uncrop_kernel(const float* __restrict__ cvol, float* __restrict__ vol, int nx, int ny, int ox, int oy, int oz,
              int cx, int cy, long long n4) {
  const int cx4 = cx >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cx4;
    const int x4 = (int)(i - r * cx4), y = (int)(r % cy), z = (int)(r / cy);
    const size_t dst = ((size_t)(z + oz) * ny + (size_t)(y + oy)) * nx + (size_t)(ox + 4 * x4);
    *reinterpret_cast<float4*>(vol + dst) = __ldg(reinterpret_cast<const float4*>(cvol) + i);
  }
}

}  // namespace ife
