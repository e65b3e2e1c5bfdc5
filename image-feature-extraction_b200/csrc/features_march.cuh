// z-marching form of the fused finite-difference Hessian + gradient magnitude + symmetric
// 3x3 eigen solve + eigenvalue features + mask (+ whole-volume DenseHistogram) kernel.
//
// Same arithmetic, operation for operation, as features_kernel in eigen_features.cuh (and
// therefore the same reference lines: Hessian3DImageFilter.hxx:11-60,
// ImageToEmphysemaFeaturesFilter.hxx:27-54, Symmetric3x3EigenvalueSolver.h:33-132,
// EigenvalueFeaturesFunctor.h:20-31, DenseHistogram.h:47-53) -- what changes is how the
// work is laid out, because the brick kernel is bound by instruction issue, not by HBM:
//
//  * a block owns a 32x8 (x,y) column and walks it in z.  Every plane passes through shared
//    memory once (double-buffered, one barrier per plane); the x/y clamps of
//    ZeroFluxNeumann are folded into loop-invariant load offsets, the z clamp into the plane
//    index, so the inner loop has no boundary logic at all.
//  * everything a voxel needs from its own plane (Dx, Dy, Dxx, Dyy, Dxy, gx^2+gy^2) is
//    computed when that plane arrives and carried in registers; the z terms (Dzz, Dxz, Dyz,
//    gz) come from the register copies of the planes below and above.  9 shared-memory
//    loads per voxel instead of 19, and 5 float->double conversions instead of 7.
//  * the solver's range tests (is every division on the branch-free Markstein path?) are one
//    integer min-reduction and one branch per voxel; voxels that fail it -- denormal or
//    huge operands -- take the out-of-line IEEE routine.
#pragma once
#include "eigen_features.cuh"

namespace ife {

constexpr int kMX = 32, kMY = 8;          // the block's (x, y) footprint
constexpr int kMPX = kMX + 2, kMPY = kMY + 2;
constexpr int kMPlane = kMPX * kMPY;      // 340 staged values per plane

// what a voxel carries from its own plane
struct PlaneTerms {
  double cD;        // centre value
  float Dx, Dy;     // first derivatives (DerivativeImageFilter output, float)
  float Dxx, Dyy, Dxy;
  double g2;        // gx^2 + gy^2 of GradientMagnitudeImageFilter (double)
};

// MODE / HIST / UNIT / ALLOUT as in features_kernel.  HIST here is the whole-volume
// histogram only (A.hist.n_roi == 0); ROI lists stay with the brick kernel, whose culling of
// bricks that touch no ROI is worth more than the march.  Masks: uint8 or none.
// Host-checked: zchunk * nx * ny < 2^31 (32-bit offsets relative to per-block base pointers).
template <int MODE, bool HIST, bool UNIT, bool ALLOUT>
__global__ void __launch_bounds__(kMX * kMY)
features_march_kernel(const __grid_constant__ StencilCoef S, const __grid_constant__ FeatArgs A,
                      const int zchunk) {
  constexpr int NFEAT = MODE == 0 ? 8 : (MODE == 1 ? 6 : 1);
  constexpr int NT = kMX * kMY;
  __shared__ float plane[2][kMPlane];
  extern __shared__ unsigned char feat_smem[];
  float* s_edges = reinterpret_cast<float*>(feat_smem);
  uint32_t* s_counts = reinterpret_cast<uint32_t*>(s_edges + NFEAT * A.hist.n_edges);
  const int nb = A.hist.n_edges + 1;
  const int tid = threadIdx.y * kMX + threadIdx.x;
  if (HIST) {
    for (int i = tid; i < NFEAT * A.hist.n_edges; i += NT) s_edges[i] = A.hist.edges[i];
    for (int i = tid; i < NFEAT * nb; i += NT) s_counts[i] = 0u;
  }

  const int nx = A.nx, ny = A.ny;
  const unsigned psz = (unsigned)nx * (unsigned)ny;          // plane stride (elements)
  const int x0 = blockIdx.x * kMX, y0 = blockIdx.y * kMY;
  const int zs = A.zb0 + blockIdx.z * zchunk;
  const int ze = min(zs + zchunk, A.zb1);                    // output planes [zs, ze)
  const int zlo = max(zs - 1, 0);                            // first plane this block touches
  // block-uniform base pointers; everything below is a 32-bit offset from them
  const float* __restrict__ vol = A.vol + (size_t)psz * zlo;
  const uint8_t* __restrict__ mask = A.mask_u8 ? A.mask_u8 + (size_t)psz * zlo : nullptr;
  float* outp[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    outp[k] = (k < NFEAT && A.out[k]) ? A.out[k] + ((size_t)psz * (zs - A.zb0) + (size_t)nx * y0 + x0) : nullptr;

  // staging slots of this thread: element tid and (for the first 84 threads) tid + 256 of
  // the padded 34x10 plane; the x/y clamps live in these loop-invariant offsets
  const int e0r = tid / kMPX, e0c = tid - e0r * kMPX;
  const int e1 = min(tid + NT, kMPlane - 1);
  const bool has1 = tid + NT < kMPlane;
  const int e1r = e1 / kMPX, e1c = e1 - e1r * kMPX;
  const unsigned g0 = (unsigned)min(max(x0 - 1 + e0c, 0), nx - 1) + (unsigned)nx * (unsigned)min(max(y0 - 1 + e0r, 0), ny - 1);
  const unsigned g1 = (unsigned)min(max(x0 - 1 + e1c, 0), nx - 1) + (unsigned)nx * (unsigned)min(max(y0 - 1 + e1r, 0), ny - 1);

  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  const bool in_xy = x < nx && y < ny;
  const unsigned vox_xy = in_xy ? (unsigned)x + (unsigned)nx * (unsigned)y : 0u;   // mask offset in a plane
  const unsigned out_xy = threadIdx.x + (unsigned)nx * threadIdx.y;              // relative to outp[]
  const int lc = (threadIdx.y + 1) * kMPX + threadIdx.x + 1;                       // centre in the padded plane
  const int zlast = A.nzb - 1 - zlo;                                               // last plane, relative to zlo

  // software pipeline, one plane per iteration:  registers (v0, v1, mraw) hold plane `pz`
  // (fetched an iteration ago) -> shared memory -> in-plane terms N; the voxel of plane
  // pz-1 is finished from (P, C, N).  Planes zs-1 and ze only feed their neighbours.
  float v0, v1;
  unsigned mraw = 0;
  v0 = __ldg(vol + g0);                     // plane zs-1 clamped = plane zlo
  v1 = has1 ? __ldg(vol + g1) : 0.0f;
  PlaneTerms P, C;
  bool insideC = false;
  P.cD = 0.0; P.Dx = P.Dy = 0.0f;
  C.cD = 0.0; C.Dx = C.Dy = C.Dxx = C.Dyy = C.Dxy = 0.0f; C.g2 = 0.0;
  float cC = 0.0f;
  unsigned oz = out_xy - 2u * psz;          // output offset of plane pz-1 (first used at pz = zs+1)
  int buf = 0;
#pragma unroll 1
  for (int pz = zs - 1; pz <= ze; ++pz) {
    float* pl = plane[buf];
    pl[tid] = v0;
    if (has1) pl[e1] = v1;
    // mask of plane pz: false for the two feeder planes and outside the image
    const bool m_own = in_xy && pz >= zs && pz < ze && (mask == nullptr || mraw != 0u);
    {   // fetch plane pz+1 (clamped; one wasted fetch after the last plane) and its mask
      const unsigned rel = (unsigned)min(pz + 1 - zlo, zlast);
      const unsigned po = psz * rel;
      v0 = __ldg(vol + (po + g0));
      if (has1) v1 = __ldg(vol + (po + g1));
      if (mask != nullptr) mraw = __ldg(mask + (po + vox_xy));
    }
    __syncthreads();

    // ---- in-plane terms of plane pz ----
    PlaneTerms N;
    const float c000 = pl[lc];
    const float xm = pl[lc - 1], xp = pl[lc + 1];
    const float ym = pl[lc - kMPX], yp = pl[lc + kMPX];
    N.cD = (double)c000;
    N.Dx = deriv1<UNIT>(S.d1[0], xm, xp);
    N.Dy = deriv1<UNIT>(S.d1[1], ym, yp);
    N.Dxx = N.Dyy = N.Dxy = 0.0f;
    N.g2 = 0.0;
    if (m_own) {
      const double dxm = (double)xm, dxp = (double)xp, dym = (double)ym, dyp = (double)yp;
      if (MODE == 0 || MODE == 1) {
        N.Dxx = deriv2<UNIT>(S.d2a[0], S.d2b[0], dxm, N.cD, dxp);
        N.Dyy = deriv2<UNIT>(S.d2a[1], S.d2b[1], dym, N.cD, dyp);
        const float dx_ym = deriv1<UNIT>(S.d1[0], pl[lc - kMPX - 1], pl[lc - kMPX + 1]);
        const float dx_yp = deriv1<UNIT>(S.d1[0], pl[lc + kMPX - 1], pl[lc + kMPX + 1]);
        N.Dxy = deriv1<UNIT>(S.d1[1], dx_ym, dx_yp);
      }
      if (MODE == 0 || MODE == 2) {
        double gx, gy;
        if (UNIT) {
          gx = __dsub_rn(dxp, dxm); gy = __dsub_rn(dyp, dym);
        } else {
          gx = __dadd_rn(__dmul_rn(-S.g1[0], dxm), __dmul_rn(S.g1[0], dxp));
          gy = __dadd_rn(__dmul_rn(-S.g1[1], dym), __dmul_rn(S.g1[1], dyp));
        }
        N.g2 = __dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy));
      }
    }

    // ---- finish the voxel of plane z = pz-1 (planes P, C, N) ----
    if (pz > zs) {
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = 0.0f;
      if (insideC) {
        if (MODE == 0 || MODE == 2) {
          // sqrt of a sum of squares of float differences: zero or far inside the double range
          double a2, scale;
          if (UNIT) {
            const double gz = __dsub_rn(N.cD, P.cD);
            a2 = __dadd_rn(C.g2, __dmul_rn(gz, gz));
            scale = 0.5;
          } else {
            const double gz = __dadd_rn(__dmul_rn(-S.g1[2], P.cD), __dmul_rn(S.g1[2], N.cD));
            a2 = __dadd_rn(C.g2, __dmul_rn(gz, gz));
            scale = 1.0;
          }
          float gm;
          if (UNIT) gm = (float)(scale * dsqrt_rn_inrange(a2));
          else gm = (float)__dsqrt_rn(a2);      // arbitrary spacing: keep the general routine
          gm = a2 > 0.0 ? gm : (float)a2;       // sqrt(+0) = +0 (a2 is never negative; NaN propagates)
          if (MODE == 0) { f[0] = cC; f[1] = gm; } else f[0] = gm;
        }
        if (MODE == 0 || MODE == 1) {
          float H[6], e[6];
          H[0] = C.Dxx;
          H[1] = C.Dxy;
          H[2] = deriv1<UNIT>(S.d1[2], P.Dx, N.Dx);                      // Dxz = Dz(Dx)
          H[3] = C.Dyy;
          H[4] = A.dy_bug ? H[2] : deriv1<UNIT>(S.d1[2], P.Dy, N.Dy);    // Dyz = Dz(Dy)
          H[5] = deriv2<UNIT>(S.d2a[2], S.d2b[2], P.cD, C.cD, N.cD);     // Dzz
          eigen_features6_lean(H, e);
          constexpr int o6 = MODE == 0 ? 2 : 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) f[o6 + k] = e[k];
        }
      }
      if (in_xy) {
#pragma unroll
        for (int k = 0; k < NFEAT; ++k)
          if (ALLOUT || outp[k]) outp[k][oz] = f[k];
      }
      if (HIST) {
#pragma unroll
        for (int k = 0; k < NFEAT; ++k) {
          const int bin = insideC ? dense_bin(s_edges + k * A.hist.n_edges, A.hist.n_edges, f[k]) : 0;
          hist_add(s_counts + k * nb, bin, insideC);
        }
      }
    }

    P.cD = C.cD; P.Dx = C.Dx; P.Dy = C.Dy;
    C = N;
    cC = c000;
    insideC = m_own;
    oz += psz;
    buf ^= 1;
  }

  if (HIST) {
    __syncthreads();
    for (int i = tid; i < NFEAT * nb; i += NT) {
      const uint32_t c = s_counts[i];
      if (c) atomicAdd(A.hist.counts + i, c);
    }
  }
}

}  // namespace ife
