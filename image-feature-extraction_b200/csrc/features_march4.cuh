// Fused finite-difference Hessian + gradient magnitude + symmetric 3x3 eigen solve + eigenvalue
// features + mask (+ whole-volume DenseHistogram), FOUR voxels per thread.
//
// Same arithmetic, operation for operation, as features_march_kernel (features_march.cuh) and
// therefore the same reference lines (Hessian3DImageFilter.hxx:11-60,
// ImageToEmphysemaFeaturesFilter.hxx:27-54, Symmetric3x3EigenvalueSolver.h:33-132,
// EigenvalueFeaturesFunctor.h:20-31, DenseHistogram.h:47-53).  What changes is the layout,
// because that kernel is bound by instruction issue and a third of what it issues is not
// arithmetic but addressing, staging and loop control paid once per voxel:
//
//  * a thread owns four x-adjacent voxels, a block of 32x4 threads a 128x4 (x, y) column that
//    it walks in z.  Planes pass through shared memory once, in a 4-slot ring filled two planes
//    ahead with 16-BYTE cp.async (one or two per thread and plane instead of three 4-byte
//    ones per voxel); the ZeroFluxNeumann x/y clamps live in loop-invariant source offsets;
//  * every output plane gets ONE 16-byte store per thread (8 per four voxels instead of 32
//    scalar stores, each with its own 64-bit address);
//  * the stencil reads its rows with LDS.128 (4.75 shared loads per voxel instead of 13); the
//    x-neighbours of the four voxels are each other, so a row of six values serves all four;
//  * nothing is carried between planes: the three planes a voxel needs are all still in the
//    ring, so there is no register rotation and no 3x unrolled role loop -- the code of one
//    step is the whole loop body.
// Requires nx % 4 == 0 and 16-byte aligned volume / output pointers (host-checked; anything
// else takes features_march_kernel).
#pragma once
#include "features_march.cuh"

namespace ife {

constexpr int kQX = 32, kQY = 4, kQV = 4;
constexpr int kQW = kQX * kQV;                // 128 voxels per row of the block's footprint
constexpr int kQPitch = kQW + 8;              // staged row: x0-1 at column 3, x0 at column 4 (16-byte aligned), x0+128 at column 132
constexpr int kQRows = kQY + 2;
constexpr int kQPlane = kQRows * kQPitch;     // 816 floats
constexpr int kQChunks = kQRows * (kQW / 4);  // 16-byte pieces per plane (192)

#ifndef IFE_MARCH4_GROUP
#define IFE_MARCH4_GROUP 2
#endif
#ifndef IFE_MARCH4_MINB
#define IFE_MARCH4_MINB 4
#endif
#ifndef IFE_MARCH4_MINB_HIST
#define IFE_MARCH4_MINB_HIST 4
#endif

__device__ __forceinline__ void row6(const float* __restrict__ row, float (&v)[6]) {   // columns -1 .. 4 around the thread's first voxel
  v[0] = row[-1];
  const float4 q = *reinterpret_cast<const float4*>(row);
  v[1] = q.x; v[2] = q.y; v[3] = q.z; v[4] = q.w;
  v[5] = row[4];
}
__device__ __forceinline__ void row4(const float* __restrict__ row, float (&v)[4]) {
  const float4 q = *reinterpret_cast<const float4*>(row);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}

// MODE / HIST / UNIT / OUTS as in features_march_kernel.
template <int MODE, bool HIST, bool UNIT, int OUTS>
__global__ void __launch_bounds__(kQX * kQY, HIST ? IFE_MARCH4_MINB_HIST : IFE_MARCH4_MINB)
features_march4_kernel(const __grid_constant__ StencilCoef S, const __grid_constant__ FeatArgs A,
                       const int zchunk) {
  constexpr int NFEAT = MODE == 0 ? 8 : ((MODE == 1 || MODE == 3) ? 6 : 1);
  constexpr int NT = kQX * kQY;
  __shared__ __align__(16) float plane[4][kQPlane];
  extern __shared__ unsigned char feat_smem[];
  float* s_edges = reinterpret_cast<float*>(feat_smem);          // rows padded with +inf to ep = 2^k
  const int ep = hist_edge_pitch(A.hist.n_edges);
  unsigned* s_priv = reinterpret_cast<unsigned*>(s_edges + NFEAT * ep);
  const int nb = A.hist.n_edges + 1;
  const int tid = threadIdx.y * kQX + threadIdx.x;
  const unsigned cnt0 = (unsigned)(NFEAT * ep) * 4u + (unsigned)((tid & 31) * 4 + (tid >> 5));
  if (HIST) {
    for (int i = tid; i < NFEAT * ep; i += NT) {
      const int k = i / ep, j = i - k * ep;
      s_edges[i] = j < A.hist.n_edges ? A.hist.edges[k * A.hist.n_edges + j] : __int_as_float(0x7f800000);
    }
    for (int i = tid; i < NFEAT * nb * 32; i += NT) s_priv[i] = 0u;
  }

  const int nx = A.nx, ny = A.ny;
  const size_t psz = (size_t)nx * (size_t)ny;
  const int x0 = blockIdx.x * kQW, y0 = blockIdx.y * kQY;
  const int zs = A.zb0 + blockIdx.z * zchunk;
  const int ze = min(zs + zchunk, A.zb1);                    // output planes [zs, ze)
  const int zlo = max(zs - 1, 0);

  // ---- staging slots of this thread (loop-invariant) ----
  // 16-byte pieces tid and tid + 128 of the plane's 192; a piece that starts at x == nx holds the
  // clamped right neighbour of the last voxel in its first element (nx % 4 == 0: pieces are
  // either inside the row or beyond it)
  int e_dst[2];
  long long e_src[2];
  int e_kind[2];   // 0 = nothing, 1 = 16 bytes, 2 = one float (clamped right neighbour)
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int id = tid + s * NT;
    const int r = id / (kQW / 4), c = id - r * (kQW / 4);
    const int gx = x0 + 4 * c;
    const int gy = min(max(y0 - 1 + r, 0), ny - 1);
    e_dst[s] = r * kQPitch + 4 + 4 * c;
    e_kind[s] = id >= kQChunks ? 0 : (gx < nx ? 1 : (gx == nx ? 2 : 0));
    e_src[s] = (long long)gy * nx + (gx < nx ? gx : nx - 1);
  }
  // halo columns: threads 0 .. 11 copy one float each (left: x0-1, right: x0+128, clamped)
  const bool halo_copier = tid < 2 * kQRows;
  const int h_r = tid >> 1, h_side = tid & 1;
  const int h_dst = (halo_copier ? h_r : 0) * kQPitch + (h_side ? 4 + kQW : 3);
  const long long h_src = (long long)min(max(y0 - 1 + (halo_copier ? h_r : 0), 0), ny - 1) * nx +
                          (h_side ? min(x0 + kQW, nx - 1) : max(x0 - 1, 0));

  const int x = x0 + kQV * threadIdx.x, y = y0 + threadIdx.y;
  const bool in_xy = x < nx && y < ny;                       // all four voxels or none (nx % 4 == 0)
  const bool has_mask = A.mask_u8 != nullptr;
  const float* pv = A.vol + psz * zlo;                       // running plane pointer of the staged plane
  const uint8_t* pm = has_mask ? A.mask_u8 + (in_xy ? (size_t)y * nx + x : 0) : nullptr;   // this thread's four mask bytes, plane 0
  const size_t vox = (size_t)nx * (size_t)y + (size_t)x;
  int first_out = 0;
#pragma unroll
  for (int k = NFEAT - 1; k >= 0; --k)
    if (A.out[k]) first_out = k;
  float* po = A.out[first_out] ? A.out[first_out] + psz * (size_t)(zs - A.zb0) + vox : nullptr;
  long long dk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    dk[k] = (k < NFEAT && A.out[k]) ? (long long)(A.out[k] - A.out[first_out]) : 0;
  size_t pkz = psz * (size_t)(zs - A.zb0);

  int pnext = zs - 1;                        // plane `pv` refers to (before clamping)
  auto issue = [&](int slot) {
    float* dst = plane[slot];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (e_kind[s] == 1) cp_async16(dst + e_dst[s], pv + e_src[s]);
      else if (e_kind[s] == 2) cp_async4(dst + e_dst[s], pv + e_src[s]);
    }
    if (halo_copier) cp_async4(dst + h_dst, pv + h_src);
    cp_async_commit();
    if (pnext >= 0 && pnext < A.nzb - 1) pv += psz;
    ++pnext;
  };
  issue(0);   // plane zs-1
  issue(1);   // plane zs
  issue(2);   // plane zs+1
  unsigned m_next = 0xffffffffu;             // mask bytes of the plane computed next (0xff.. = no mask)
  if (has_mask && in_xy) m_next = __ldg(reinterpret_cast<const unsigned*>(pm + psz * (size_t)zs));

  const int lc = (threadIdx.y + 1) * kQPitch + 4 + kQV * threadIdx.x;   // the thread's first voxel in a staged plane

#pragma unroll 1
  for (int z = zs; z < ze; ++z) {
    // planes z-1, z, z+1 are needed; z+1 is the copy issued a step ago.  The barrier also says that
    // every thread is done with step z-1, i.e. with plane z-2, whose slot the next copy takes.
    cp_async_wait<0>();
    __syncthreads();
    issue((z - zs + 3) & 3);                 // plane z+2, needed one step from now
    const unsigned m_cur = opaque_u32(m_next);
    if (has_mask && in_xy && z + 1 < ze) m_next = __ldg(reinterpret_cast<const unsigned*>(pm + psz * (size_t)(z + 1)));
    const float* pP = plane[(z - zs) & 3] + lc;       // plane z-1
    const float* pC = plane[(z - zs + 1) & 3] + lc;   // plane z
    const float* pN = plane[(z - zs + 2) & 3] + lc;   // plane z+1

    float f[kQV][8];
    bool inside[kQV];
    bool any_inside = false;
#pragma unroll
    for (int i = 0; i < kQV; ++i) {
      inside[i] = in_xy && ((m_cur >> (8 * i)) & 0xffu) != 0u;
      any_inside = any_inside || inside[i];
    }

    if (!any_inside) {
#pragma unroll
      for (int i = 0; i < kQV; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) f[i][k] = 0.0f;
    } else {
#pragma unroll
      for (int i = 0; i < kQV; ++i)
#pragma unroll
        for (int k = NFEAT; k < 8; ++k) f[i][k] = 0.0f;   // slots this mode never writes
      float a[6], b[6], c[6], bm[6], bn[6];
      row6(pC, b);
      row6(pP, bm);
      row6(pN, bn);
      float am[4], cm[4], an[4], cn[4];   // rows y-1 / y+1 of the planes below and above
      if (MODE != 2) {
        row6(pC - kQPitch, a);
        row6(pC + kQPitch, c);
        row4(pP - kQPitch, am);
        row4(pP + kQPitch, cm);
        row4(pN - kQPitch, an);
        row4(pN + kQPitch, cn);
      } else {
        float t4[4];
        row4(pC - kQPitch, t4);
        a[1] = t4[0]; a[2] = t4[1]; a[3] = t4[2]; a[4] = t4[3];
        row4(pC + kQPitch, t4);
        c[1] = t4[0]; c[2] = t4[1]; c[3] = t4[2]; c[4] = t4[3];
      }
      double bd[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) bd[j] = (double)b[j];
      // all four voxels are computed whenever one of them is wanted (branch-free: the compiler
      // interleaves their dependency chains); unwanted ones are zeroed at the end
      float H[kQV][6];
#pragma unroll
      for (int i = 0; i < kQV; ++i) {
        const double cD = bd[1 + i], dxm = bd[i], dxp = bd[2 + i];
        const double dym = (double)a[1 + i], dyp = (double)c[1 + i];
        const double PcD = (double)bm[1 + i], NcD = (double)bn[1 + i];
        if (MODE == 0 || MODE == 2) {
          // GradientMagnitudeImageFilter (double accumulate), sqrt of a sum of squares of float differences
          float gm;
          if (UNIT) {
            const double gx = __dsub_rn(dxp, dxm), gy = __dsub_rn(dyp, dym);
            const double g2 = __dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy));
            const double gz = __dsub_rn(NcD, PcD);
            const double a2 = __dadd_rn(g2, __dmul_rn(gz, gz));
            gm = (float)(0.5 * dsqrt_rn_inrange(a2));
            gm = a2 > 0.0 ? gm : (float)a2;     // sqrt(+0) = +0 (a2 is never negative; NaN propagates)
          } else {
            const double gx = __dadd_rn(__dmul_rn(-S.g1[0], dxm), __dmul_rn(S.g1[0], dxp));
            const double gy = __dadd_rn(__dmul_rn(-S.g1[1], dym), __dmul_rn(S.g1[1], dyp));
            const double g2 = __dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy));
            const double gz = __dadd_rn(__dmul_rn(-S.g1[2], PcD), __dmul_rn(S.g1[2], NcD));
            gm = (float)__dsqrt_rn(__dadd_rn(g2, __dmul_rn(gz, gz)));
          }
          if (MODE == 0) { f[i][0] = b[1 + i]; f[i][1] = gm; } else f[i][0] = gm;
        }
        if (MODE == 0 || MODE == 1 || MODE == 3) {
          H[i][0] = deriv2<UNIT>(S.d2a[0], S.d2b[0], dxm, cD, dxp);                               // Dxx
          const float dx_ym = deriv1<UNIT>(S.d1[0], a[i], a[2 + i]);
          const float dx_yp = deriv1<UNIT>(S.d1[0], c[i], c[2 + i]);
          H[i][1] = deriv1<UNIT>(S.d1[1], dx_ym, dx_yp);                                          // Dxy = Dy(Dx)
          const float PDx = deriv1<UNIT>(S.d1[0], bm[i], bm[2 + i]);
          const float NDx = deriv1<UNIT>(S.d1[0], bn[i], bn[2 + i]);
          H[i][2] = deriv1<UNIT>(S.d1[2], PDx, NDx);                                              // Dxz = Dz(Dx)
          H[i][3] = deriv2<UNIT>(S.d2a[1], S.d2b[1], dym, cD, dyp);                               // Dyy
          const float PDy = deriv1<UNIT>(S.d1[1], am[i], cm[i]);
          const float NDy = deriv1<UNIT>(S.d1[1], an[i], cn[i]);
          H[i][4] = A.dy_bug ? H[i][2] : deriv1<UNIT>(S.d1[2], PDy, NDy);                         // Dyz = Dz(Dy)
          H[i][5] = deriv2<UNIT>(S.d2a[2], S.d2b[2], PcD, cD, NcD);                               // Dzz
        }
      }
      if (MODE == 0 || MODE == 1 || MODE == 3) {
        constexpr int o6 = MODE == 0 ? 2 : 0;
        if (MODE == 3) {
#pragma unroll
          for (int i = 0; i < kQV; ++i)
#pragma unroll
            for (int k = 0; k < 6; ++k) f[i][o6 + k] = H[i][k];
        } else {
          constexpr int G = IFE_MARCH4_GROUP;   // voxels whose eigen solves run as one straight-line block
#pragma unroll
          for (int g = 0; g < kQV; g += G) {
            float Hg[G][6], eg[G][6];
#pragma unroll
            for (int i = 0; i < G; ++i)
#pragma unroll
              for (int k = 0; k < 6; ++k) Hg[i][k] = H[g + i][k];
            eigen_features6_lean_n<G>(Hg, eg);
#pragma unroll
            for (int i = 0; i < G; ++i)
#pragma unroll
              for (int k = 0; k < 6; ++k) f[g + i][o6 + k] = eg[i][k];
          }
        }
      }
      if (!(inside[0] && inside[1] && inside[2] && inside[3])) {   // rare inside a mask, never without one
#pragma unroll
        for (int i = 0; i < kQV; ++i)
#pragma unroll
          for (int k = 0; k < 8; ++k) f[i][k] = inside[i] ? f[i][k] : 0.0f;
      }
    }

    if (in_xy) {
#pragma unroll
      for (int k = 0; k < NFEAT; ++k)
        if (OUTS == 1 || (OUTS == 0 && A.out[k]))
          *reinterpret_cast<float4*>(reinterpret_cast<char*>(po) + dk[k] * 4) = make_float4(f[0][k], f[1][k], f[2][k], f[3][k]);
    }

    if (HIST) {
      unsigned long long packed[kQV];
      // 4 x NFEAT independent searches advance together (step-major): every load of a step is in
      // flight at once; voxels outside the mask search too (their f is 0) but never count
      unsigned off[kQV][NFEAT];
      const char* eb = reinterpret_cast<const char*>(s_edges);
#pragma unroll
      for (int i = 0; i < kQV; ++i) packed[i] = ~0ull;
      if (!any_inside) {
        // nothing to count in this thread's four voxels
      } else {
      if (ep == 64) {
#pragma unroll
        for (int i = 0; i < kQV; ++i)
#pragma unroll
          for (int k = 0; k < NFEAT; ++k) off[i][k] = 0u;
#pragma unroll
        for (int step = 32; step >= 1; step >>= 1) {
#pragma unroll
          for (int i = 0; i < kQV; ++i) {
            float ev[NFEAT];
#pragma unroll
            for (int k = 0; k < NFEAT; ++k)
              ev[k] = *reinterpret_cast<const float*>(eb + off[i][k] + (unsigned)(k * 256 + (step - 1) * 4));
#pragma unroll
            for (int k = 0; k < NFEAT; ++k)
              if (ev[k] < f[i][k]) off[i][k] += (unsigned)(step * 4);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < kQV; ++i)
#pragma unroll
          for (int k = 0; k < NFEAT; ++k) off[i][k] = 4u * (unsigned)dense_bin_padded_rt(s_edges + k * ep, ep, f[i][k]);
      }
#pragma unroll
      for (int i = 0; i < kQV; ++i) {
        if (A.hist.packed) {
          unsigned long long w = 0;
#pragma unroll
          for (int k = 0; k < NFEAT; ++k) w |= (unsigned long long)(off[i][k] >> 2) << (8 * k);
          if (inside[i]) packed[i] = w;
        } else if (inside[i]) {
          unsigned char* cb = reinterpret_cast<unsigned char*>(s_edges) + cnt0;
          unsigned cv[NFEAT];
#pragma unroll
          for (int k = 0; k < NFEAT; ++k) cv[k] = cb[(unsigned)(k * nb) * 128u + (off[i][k] << 5)];
#pragma unroll
          for (int k = 0; k < NFEAT; ++k) cb[(unsigned)(k * nb) * 128u + (off[i][k] << 5)] = (unsigned char)(cv[k] + 1u);
        }
      }
      }
      if (A.hist.packed && in_xy) {
        ulonglong2* dst = reinterpret_cast<ulonglong2*>(A.hist.packed + pkz + vox);
        dst[0] = make_ulonglong2(packed[0], packed[1]);
        dst[1] = make_ulonglong2(packed[2], packed[3]);
      }
      pkz += psz;
    }
    po += psz;
  }
  cp_async_wait<0>();

  if (HIST) {
    __syncthreads();
    for (int i = tid; i < NFEAT * nb; i += NT) {
      const unsigned* w = s_priv + i * 32;
      unsigned total = 0;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) total = __dp4a(w[(j + tid) & 31], 0x01010101u, total);
      if (total) atomicAdd(A.hist.counts + i, total);
    }
  }
}

}  // namespace ife
