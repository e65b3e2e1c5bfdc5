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
#include "cp_async.cuh"
#include "eigen_features.cuh"

namespace ife {

#ifndef IFE_MARCH_MINB
#define IFE_MARCH_MINB 8   // resident 128-thread blocks per SM the register budget is cut for
#endif
#ifndef IFE_MARCH_MINB_HIST
#define IFE_MARCH_MINB_HIST 3   // histogram variant: 42 KB of private counters per block allow four
                                // blocks per SM; the looser register bound (114 instead of 90) saves
                                // rematerialisation and still fits four (2.41 -> 2.33 ms)
#endif
constexpr int kMX = 32, kMY = 4;

// smallest power of two above n_edges (edge rows are padded with +inf up to it)
__host__ __device__ inline int hist_edge_pitch(int n_edges) {
  int ep = 2;
  while (ep <= n_edges) ep <<= 1;
  return ep;
}          // the block's (x, y) footprint
constexpr int kMPX = kMX + 2, kMPY = kMY + 2;
constexpr int kMPlane = kMPX * kMPY;      // 340 staged values per plane

// what a voxel carries from its own plane
struct PlaneTerms {
  double cD;        // centre value
  double g2;        // gx^2 + gy^2 of GradientMagnitudeImageFilter (double)
  float c;
  float Dx, Dy;     // first derivatives (DerivativeImageFilter output, float)
  float Dxx, Dyy, Dxy;
  bool inside;      // in the image, in the output range and in the mask
  unsigned mraw;    // this plane's mask byte, fetched two steps ahead of its use
};

// keeps the compiler from testing a prefetched value the moment it arrives (which would stall
// on the load an iteration early): the value stays opaque until the statement executes
__device__ __forceinline__ unsigned opaque_u32(unsigned v) {
  asm volatile("" : "+r"(v));
  return v;
}

// MODE / HIST / UNIT as in features_kernel; MODE 3 (this kernel only) = the six Hessian entries
// themselves, no eigen solve.  HIST here is the whole-volume
// histogram only (A.hist.n_roi == 0); ROI lists stay with the brick kernel, whose culling of
// bricks that touch no ROI is worth more than the march.  Masks: uint8 or none.
// OUTS: 1 = every output plane pointer is set, 2 = none is (histograms only), 0 = test each
template <int MODE, bool HIST, bool UNIT, int OUTS>
__global__ void __launch_bounds__(kMX * kMY, HIST ? IFE_MARCH_MINB_HIST : IFE_MARCH_MINB)
features_march_kernel(const __grid_constant__ StencilCoef S, const __grid_constant__ FeatArgs A,
                      const int zchunk) {
  constexpr int NFEAT = MODE == 0 ? 8 : ((MODE == 1 || MODE == 3) ? 6 : 1);
  constexpr int NT = kMX * kMY;
  __shared__ float plane[4][kMPlane];   // ring: plane pz+2 lands while plane pz is consumed
  __shared__ __align__(16) unsigned char s_mask[4][kMX * kMY];   // the mask bytes of the same planes
  extern __shared__ unsigned char feat_smem[];
  // Histogram sink.  Shared-memory atomics run at about one lane per clock per SM, far too
  // slow for eight inserts per voxel, so there are none: every thread owns a private 8-bit
  // counter per (feature, bin), byte (tid >> 5) of word [(k*nb + bin)*32 + (tid & 31)] -- a
  // warp's 32 lanes always hit 32 different banks whatever their bins, the four warps of the
  // block own the four bytes of a word -- an insert is a byte load/add/store at
  // (bin << 7) + a per-thread constant, and a thread sees at most zchunk <= 255 voxels, so a
  // counter cannot overflow.  Rows are summed once per block.
  float* s_edges = reinterpret_cast<float*>(feat_smem);          // rows padded with +inf to ep = 2^k
  const int ep = hist_edge_pitch(A.hist.n_edges);
  unsigned* s_priv = reinterpret_cast<unsigned*>(s_edges + NFEAT * ep);
  const int nb = A.hist.n_edges + 1;
  const int tid = threadIdx.y * kMX + threadIdx.x;
  // this thread's counter of (feature 0, bin 0), as a byte offset from s_edges (everything the
  // sink touches is addressed by 32-bit offsets from that one shared-memory base)
  const unsigned cnt0 = (unsigned)(NFEAT * ep) * 4u + (unsigned)((tid & 31) * 4 + (tid >> 5));
  if (HIST) {
    for (int i = tid; i < NFEAT * ep; i += NT) {
      const int k = i / ep, j = i - k * ep;
      s_edges[i] = j < A.hist.n_edges ? A.hist.edges[k * A.hist.n_edges + j] : __int_as_float(0x7f800000);
    }
    for (int i = tid; i < NFEAT * nb * 32; i += NT) s_priv[i] = 0u;
  }

  const int nx = A.nx, ny = A.ny;
  const size_t psz = (size_t)nx * (size_t)ny;                // plane stride (elements)
  const int x0 = blockIdx.x * kMX, y0 = blockIdx.y * kMY;
  const int zs = A.zb0 + blockIdx.z * zchunk;
  const int ze = min(zs + zchunk, A.zb1);                    // output planes [zs, ze)
  const int zlo = max(zs - 1, 0);                            // first plane this block touches

  // staging slots of this thread: elements tid and (while inside the padded (kMX+2) x (kMY+2)
  // plane) tid + NT; the x/y clamps live in these loop-invariant offsets
  const int e0r = tid / kMPX, e0c = tid - e0r * kMPX;
  const int e1 = min(tid + NT, kMPlane - 1);
  const bool has1 = tid + NT < kMPlane;
  const int e1r = e1 / kMPX, e1c = e1 - e1r * kMPX;
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  const bool in_xy = x < nx && y < ny;
  const bool has_mask = A.mask_u8 != nullptr;
  // 4-byte aligned mask rows travel with cp.async like the image (threads 0 .. 8*kMY-1 copy four
  // bytes each); anything else is fetched per voxel into a register two steps ahead
  const bool mask_async = has_mask && (nx & 3) == 0 && (reinterpret_cast<size_t>(A.mask_u8) & 3) == 0;
  const bool mask_copier = mask_async && tid < 8 * kMY;
  // running pointers: plane `pz+1` of the two staged elements and of this voxel's mask byte
  const float* p0 = A.vol + psz * zlo + (size_t)min(max(x0 - 1 + e0c, 0), nx - 1) +
                    (size_t)nx * (size_t)min(max(y0 - 1 + e0r, 0), ny - 1);
  const float* p1 = A.vol + psz * zlo + (size_t)min(max(x0 - 1 + e1c, 0), nx - 1) +
                    (size_t)nx * (size_t)min(max(y0 - 1 + e1r, 0), ny - 1);
  const uint8_t* pm = (has_mask ? A.mask_u8 : reinterpret_cast<const uint8_t*>(A.vol)) + psz * zlo;
  if (mask_async) {
    if (mask_copier) pm += (size_t)nx * (size_t)min(y0 + (tid >> 3), ny - 1) + (size_t)min(x0 + 4 * (tid & 7), nx - 4);
  } else {
    pm += in_xy ? (size_t)x + (size_t)nx * (size_t)y : 0;
  }
  // output pointer of plane 0 of this voxel; the other planes are uniform byte offsets away
  const size_t o_first = psz * (size_t)(zs - A.zb0) + (size_t)nx * (size_t)y + (size_t)x;
  int first_out = 0;
#pragma unroll
  for (int k = NFEAT - 1; k >= 0; --k)
    if (A.out[k]) first_out = k;
  float* po = A.out[first_out] ? A.out[first_out] + o_first : nullptr;
  long long dk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    dk[k] = (k < NFEAT && A.out[k]) ? (long long)(A.out[k] - A.out[first_out]) : 0;
  const int lc = (threadIdx.y + 1) * kMPX + threadIdx.x + 1;  // centre in the padded plane
  const size_t vox_pk = (size_t)nx * (size_t)y + (size_t)x;    // packed-bin sink: voxel offset in a plane
  size_t pkz = psz * (size_t)(zs - A.zb0);                    //   and the running plane offset

  // software pipeline, one plane per step.  Plane pz+2 is copied global -> shared with
  // cp.async (two steps of work hide the HBM latency, no registers held), its mask byte
  // goes into the PlaneTerms that will play the role of N two steps later; the in-plane
  // terms N of plane pz are computed from shared memory and the voxel of plane pz-1 is
  // finished from (P, C, N).  Planes zs-1 and ze only feed their neighbours.  A ring slot is
  // rewritten two barriers after its last reader, so one barrier per plane is enough.
  int pnext = zs - 1;                        // plane the pointers refer to (before clamping)
  auto advance = [&]() {
    if (pnext >= 0 && pnext < A.nzb - 1) { p0 += psz; p1 += psz; pm += psz; }
    ++pnext;
  };
  auto issue = [&](int slot) {
    cp_async4(&plane[slot][tid], p0);
    if (has1) cp_async4(&plane[slot][e1], p1);
    if (mask_copier) cp_async4(&s_mask[slot][4 * tid], pm);
    cp_async_commit();
  };
  PlaneTerms T0, T1, T2;
  T0.cD = T1.cD = T2.cD = 0.0; T0.g2 = T1.g2 = T2.g2 = 0.0;
  T0.c = T1.c = T2.c = 0.0f; T0.Dx = T1.Dx = T2.Dx = 0.0f; T0.Dy = T1.Dy = T2.Dy = 0.0f;
  T0.Dxx = T1.Dxx = T2.Dxx = 0.0f; T0.Dyy = T1.Dyy = T2.Dyy = 0.0f; T0.Dxy = T1.Dxy = T2.Dxy = 0.0f;
  T0.inside = T1.inside = T2.inside = false;
  T0.mraw = T1.mraw = T2.mraw = 0u;
  issue(0);                                  // plane zs-1
  advance();
  issue(1);                                  // plane zs; T0 is N at step zs
  if (has_mask && !mask_async) T0.mraw = __ldg(pm);
  advance();

  auto step = [&](const int pz, const PlaneTerms& P, PlaneTerms& C, PlaneTerms& N) {
    issue((pz - zs + 3) & 3);                // plane pz+2; C is N at step pz+2
    if (has_mask && !mask_async) C.mraw = __ldg(pm);
    advance();
    cp_async_wait<2>();
    __syncthreads();
    const float* pl = plane[(pz - zs + 1) & 3];
    // mask of plane pz: false for the two feeder planes and outside the image
    const unsigned mbyte = mask_async ? (unsigned)s_mask[(pz - zs + 1) & 3][tid] : opaque_u32(N.mraw);
    N.inside = in_xy && pz >= zs && pz < ze && (!has_mask || mbyte != 0u);

    // ---- in-plane terms of plane pz ----
    const float c000 = pl[lc];
    const float xm = pl[lc - 1], xp = pl[lc + 1];
    const float ym = pl[lc - kMPX], yp = pl[lc + kMPX];
    N.c = c000;
    N.cD = (double)c000;
    N.Dx = deriv1<UNIT>(S.d1[0], xm, xp);
    N.Dy = deriv1<UNIT>(S.d1[1], ym, yp);
    if (N.inside) {
      const double dxm = (double)xm, dxp = (double)xp, dym = (double)ym, dyp = (double)yp;
      if (MODE == 0 || MODE == 1 || MODE == 3) {
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
      if (C.inside) {
        if (MODE == 0 || MODE == 2) {
          // sqrt of a sum of squares of float differences: zero or far inside the double range
          float gm;
          if (UNIT) {
            const double gz = __dsub_rn(N.cD, P.cD);
            const double a2 = __dadd_rn(C.g2, __dmul_rn(gz, gz));
#if IFE_EXP == 2
            gm = (float)(0.5 * a2);
#else
            gm = (float)(0.5 * dsqrt_rn_inrange(a2));
#endif
            gm = a2 > 0.0 ? gm : (float)a2;     // sqrt(+0) = +0 (a2 is never negative; NaN propagates)
          } else {                              // arbitrary spacing: keep the general routine
            const double gz = __dadd_rn(__dmul_rn(-S.g1[2], P.cD), __dmul_rn(S.g1[2], N.cD));
            gm = (float)__dsqrt_rn(__dadd_rn(C.g2, __dmul_rn(gz, gz)));
          }
          if (MODE == 0) { f[0] = C.c; f[1] = gm; } else f[0] = gm;
        }
        if (MODE == 0 || MODE == 1 || MODE == 3) {
          float H[6], e[6];
          H[0] = C.Dxx;
          H[1] = C.Dxy;
          H[2] = deriv1<UNIT>(S.d1[2], P.Dx, N.Dx);                      // Dxz = Dz(Dx)
          H[3] = C.Dyy;
          H[4] = A.dy_bug ? H[2] : deriv1<UNIT>(S.d1[2], P.Dy, N.Dy);    // Dyz = Dz(Dy)
          H[5] = deriv2<UNIT>(S.d2a[2], S.d2b[2], P.cD, C.cD, N.cD);     // Dzz
#if IFE_EXP == 5
#pragma unroll
          for (int k = 0; k < 6; ++k) e[k] = H[k];
#else
          if (MODE == 3) {   // itk::Hessian3DImageFilter's own output: [Dxx, Dxy, Dxz, Dyy, Dyz, Dzz] (.hxx:53-59)
#pragma unroll
            for (int k = 0; k < 6; ++k) e[k] = H[k];
          } else {
            eigen_features6_lean(H, e);
          }
#endif
          constexpr int o6 = MODE == 0 ? 2 : 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) f[o6 + k] = e[k];
        }
#if IFE_EXP == 4
        { float ssum = 0.0f;
#pragma unroll
          for (int k = 0; k < NFEAT; ++k) ssum += f[k];
          *po = ssum; }
#else
#pragma unroll
        for (int k = 0; k < NFEAT; ++k)
          if (OUTS == 1 || (OUTS == 0 && A.out[k])) *reinterpret_cast<float*>(reinterpret_cast<char*>(po) + dk[k] * 4) = f[k];
#endif
      } else {
#pragma unroll
        for (int k = 0; k < NFEAT; ++k) f[k] = 0.0f;
        if (in_xy) {
#pragma unroll
          for (int k = 0; k < NFEAT; ++k)
            if (OUTS == 1 || (OUTS == 0 && A.out[k])) *reinterpret_cast<float*>(reinterpret_cast<char*>(po) + dk[k] * 4) = 0.0f;
        }
      }
      if (HIST && C.inside) {
        // eight independent searches, then eight counter updates: nothing here waits on
        // anything but its own feature.  off[k] = 4 * (number of edges < f[k]).
        unsigned off[NFEAT];
        const char* eb = reinterpret_cast<const char*>(s_edges);
        if (ep == 64) {
          // step-major order: the eight searches advance together, so the eight loads of a
          // step are in flight at once instead of 48 dependent loads back to back; a step is
          // a load at an immediate offset, a compare and a predicated add
#pragma unroll
          for (int k = 0; k < NFEAT; ++k) off[k] = 0u;
#pragma unroll
          for (int step = 32; step >= 1; step >>= 1) {
            float ev[NFEAT];
#pragma unroll
            for (int k = 0; k < NFEAT; ++k)
              ev[k] = *reinterpret_cast<const float*>(eb + off[k] + (unsigned)(k * 256 + (step - 1) * 4));
#pragma unroll
            for (int k = 0; k < NFEAT; ++k)
              if (ev[k] < f[k]) off[k] += (unsigned)(step * 4);
          }
        } else {
#pragma unroll
          for (int k = 0; k < NFEAT; ++k) off[k] = 4u * (unsigned)dense_bin_padded_rt(s_edges + k * ep, ep, f[k]);
        }
        if (A.hist.packed) {
          unsigned long long w = 0;
#pragma unroll
          for (int k = 0; k < NFEAT; ++k) w |= (unsigned long long)(off[k] >> 2) << (8 * k);
          A.hist.packed[pkz + vox_pk] = w;
        } else {
          unsigned char* cb = reinterpret_cast<unsigned char*>(s_edges) + cnt0;
          unsigned cv[NFEAT];
#pragma unroll
          for (int k = 0; k < NFEAT; ++k) cv[k] = cb[(unsigned)(k * nb) * 128u + (off[k] << 5)];
#pragma unroll
          for (int k = 0; k < NFEAT; ++k) cb[(unsigned)(k * nb) * 128u + (off[k] << 5)] = (unsigned char)(cv[k] + 1u);
        }
      } else if (HIST && A.hist.packed && in_xy) {
        A.hist.packed[pkz + vox_pk] = ~0ull;              // outside the mask
      }
      if (HIST) pkz += psz;
      po += psz;
    }
  };

  // the three roles rotate through T0/T1/T2, so nothing is ever shifted between registers
#pragma unroll 1
  for (int pz = zs - 1; pz <= ze; pz += 3) {
    step(pz, T0, T1, T2);
    if (pz + 1 > ze) break;
    step(pz + 1, T1, T2, T0);
    if (pz + 2 > ze) break;
    step(pz + 2, T2, T0, T1);
  }
  cp_async_wait<0>();

  if (HIST) {
    __syncthreads();
    // thread i sums counter row i = (feature, bin): 32 words of four 8-bit counters each
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
