// TEST INFRASTRUCTURE ONLY (oracle).  CPU restatement of the ITK primitives the reference
// delegates its smoothing / stencil arithmetic to.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may call this; the product
// (libife_cuda.so) never links or loads it.
//
// PARITY UNPINNED for this file: ITK (Insight Toolkit 4.x, version not pinned by the
// reference: CMakeLists.txt:14 `find_package( ITK REQUIRED )`) is NOT under /root/reference
// and is not installed in this image, and the reference ships no test or fixture for any
// of these stages.  What follows restates ITK 4.x's published
// algorithms; every recalled constant is named in the block below so that it can be
// diffed against a real ITK checkout later.  Call sites in the reference that fix WHICH
// primitives are used and how they are wired:
//   itk::SmoothingRecursiveGaussianImageFilter  include/ife/Filters/NormalizedGaussianConvolutionImageFilter.h:50,72 ; .hxx:51-55
//   itk::MultiplyImageFilter / DivideImageFilter include/ife/Filters/NormalizedGaussianConvolutionImageFilter.hxx:48-49,57-58
//   itk::DerivativeImageFilter                   include/ife/Filters/Hessian3DImageFilter.hxx:19-51 ; tools/FiniteDifference_HessianFeatures.cxx:127-172
//   itk::GradientMagnitudeImageFilter            include/ife/Filters/ImageToEmphysemaFeaturesFilter.h:80-82 ; tools/FiniteDifference_GradientFeatures.cxx:105-107
//   itk::MaskImageFilter / CastImageFilter       include/ife/Filters/ImageToEmphysemaFeaturesFilter.hxx:21,44-54,110-116
//
// Memory layout everywhere: x fastest, idx = x + nx*(y + ny*z); float32 pixels.
//
// Arithmetic modes of the recursive Gaussian ("arith"):
//   0 = PLAIN : every multiply and add rounded separately, left to right, exactly as
//               ITK's source reads when built without FMA contraction (x86-64 default).
//   1 = FMA   : the same expressions contracted the way `g++ -O2 -mfma` contracts a
//               left-to-right sum of products: t = a0*c0; t = fma(a1,c1,t); ...
//               (IEEE fma, so CPU and GPU agree bit for bit).
// Both are "the reference" up to its build flags; the product implements both and is
// compared bit-exactly against the matching mode.
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "oracle.h"
#include "oracle_line.h"

// defined in oracle_itk_fma.cpp (compiled with -mfma)
extern "C" void orc_line_fma(const double* c20, const double* data, double* outs, double* scratch, int ln);

namespace {

// ---------------------------------------------------------------------------------
// [ITK-recalled] constants of itk::RecursiveGaussianImageFilter (zero order, Deriche
// 4th-order approximation; Farneback & Westin parameters).
// ---------------------------------------------------------------------------------
const double kA1 = 1.3530, kB1 = 1.8151, kW1 = 0.6681, kL1 = -1.3932;
const double kA2 = -0.3531, kB2 = 0.0902, kW2 = 2.0787, kL2 = -1.3732;
const double kSpacingTolerance = 1e-8;  // below this ITK uses sigma unscaled

}  // namespace

// [ITK-recalled] RecursiveGaussianImageFilter::SetUp, zero order, NormalizeAcrossScale
// = false (SmoothingRecursiveGaussianImageFilter default).  Output order:
// N0..N3, D1..D4, M1..M4, BN1..BN4, BM1..BM4  (20 doubles).
extern "C" void orc_gaussian_coefficients(double sigma, double spacing, double* c20) {
  const double sigmad = (spacing < kSpacingTolerance) ? sigma : sigma / spacing;

  const double sin1 = std::sin(kW1 / sigmad), sin2 = std::sin(kW2 / sigmad);
  const double cos1 = std::cos(kW1 / sigmad), cos2 = std::cos(kW2 / sigmad);
  const double exp1 = std::exp(kL1 / sigmad), exp2 = std::exp(kL2 / sigmad);

  // D coefficients
  double D4 = exp1 * exp1 * exp2 * exp2;
  double D3 = -2 * cos1 * exp1 * exp2 * exp2;
  D3 += -2 * cos2 * exp2 * exp1 * exp1;
  double D2 = 4 * cos2 * cos1 * exp1 * exp2;
  D2 += exp1 * exp1 + exp2 * exp2;
  double D1 = -2 * (exp2 * cos2 + exp1 * cos1);
  const double SD = 1.0 + D1 + D2 + D3 + D4;

  // N coefficients (zero order)
  double N0 = kA1 + kA2;
  double N1 = exp2 * (kB2 * sin2 - (kA2 + 2 * kA1) * cos2);
  N1 += exp1 * (kB1 * sin1 - (kA1 + 2 * kA2) * cos1);
  double N2 = (kA1 + kA2) * cos2 * cos1;
  N2 -= kB1 * cos2 * sin1 + kB2 * cos1 * sin2;
  N2 *= 2 * exp1 * exp2;
  N2 += kA2 * exp1 * exp1 + kA1 * exp2 * exp2;
  double N3 = exp2 * exp1 * exp1 * (kB2 * sin2 - kA2 * cos2);
  N3 += exp1 * exp2 * exp2 * (kB1 * sin1 - kA1 * cos1);
  const double SN0 = N0 + N1 + N2 + N3;

  // unit DC gain
  const double across_scale_normalization = 1.0;
  const double alpha0 = 2 * SN0 / SD - N0;
  N0 *= across_scale_normalization / alpha0;
  N1 *= across_scale_normalization / alpha0;
  N2 *= across_scale_normalization / alpha0;
  N3 *= across_scale_normalization / alpha0;

  // symmetric anticausal part
  const double M1 = N1 - D1 * N0;
  const double M2 = N2 - D2 * N0;
  const double M3 = N3 - D3 * N0;
  const double M4 = -D4 * N0;

  // boundary coefficients: the edge value is assumed to extend to infinity
  const double SN = N0 + N1 + N2 + N3;
  const double SM = M1 + M2 + M3 + M4;
  const double SDb = 1.0 + D1 + D2 + D3 + D4;
  const double BN1 = D1 * SN / SDb, BN2 = D2 * SN / SDb, BN3 = D3 * SN / SDb, BN4 = D4 * SN / SDb;
  const double BM1 = D1 * SM / SDb, BM2 = D2 * SM / SDb, BM3 = D3 * SM / SDb, BM4 = D4 * SM / SDb;

  const double out[20] = {N0, N1, N2, N3, D1, D2, D3, D4, M1, M2,
                          M3, M4, BN1, BN2, BN3, BN4, BM1, BM2, BM3, BM4};
  std::memcpy(c20, out, sizeof(out));
}

namespace {

// One RecursiveGaussianImageFilter pass along `axis` (0=x,1=y,2=z): every line is copied
// into a double buffer, filtered in double, stored back as float.  [ITK-recalled]
template <bool FMA>
void gaussian_pass(const float* in, float* out, int nx, int ny, int nz, int axis,
                   const double* c20, int n_threads) {
  const size_t sx = 1, sy = (size_t)nx, sz = (size_t)nx * ny;
  const int ln = axis == 0 ? nx : (axis == 1 ? ny : nz);
  const size_t stride = axis == 0 ? sx : (axis == 1 ? sy : sz);
  // enumerate lines by the two other axes (a fastest)
  const int na = axis == 0 ? ny : nx;
  const int nb = axis == 2 ? ny : nz;
  const size_t sa = axis == 0 ? sy : sx;
  const size_t sb = axis == 2 ? sy : sz;
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
  {
    std::vector<double> inps(ln), outs(ln), scratch(ln);
#pragma omp for schedule(static) collapse(2)
    for (int b = 0; b < nb; ++b) {
      for (int a = 0; a < na; ++a) {
        const size_t base = (size_t)a * sa + (size_t)b * sb;
        for (int i = 0; i < ln; ++i) inps[i] = in[base + i * stride];
        if (FMA)
          orc_line_fma(c20, inps.data(), outs.data(), scratch.data(), ln);
        else
          orc_detail::filter_line<false>(c20, inps.data(), outs.data(), scratch.data(), ln);
        for (int i = 0; i < ln; ++i) out[base + i * stride] = static_cast<float>(outs[i]);
      }
    }
  }
}

}  // namespace

extern "C" {

// itk::SmoothingRecursiveGaussianImageFilter on a float image: cascade of zero-order
// recursive Gaussian passes, last image dimension first then x then y (z, x, y), float
// storage between passes, sigma in physical units.  [ITK-recalled]  Returns 1 if any
// axis has fewer than 4 samples (ITK throws in that case), else 0.
int orc_smoothing_recursive_gaussian(const float* in, float* out, int nx, int ny, int nz,
                                     const double* spacing, double sigma, int arith,
                                     int n_threads) {
  if (nx < 4 || ny < 4 || nz < 4) return 1;
  const size_t n = (size_t)nx * ny * nz;
  std::vector<float> tmp(n);
  double cz[20], cx[20], cy[20];
  orc_gaussian_coefficients(sigma, spacing[2], cz);
  orc_gaussian_coefficients(sigma, spacing[0], cx);
  orc_gaussian_coefficients(sigma, spacing[1], cy);
  if (arith) {
    gaussian_pass<true>(in, out, nx, ny, nz, 2, cz, n_threads);
    gaussian_pass<true>(out, tmp.data(), nx, ny, nz, 0, cx, n_threads);
    gaussian_pass<true>(tmp.data(), out, nx, ny, nz, 1, cy, n_threads);
  } else {
    gaussian_pass<false>(in, out, nx, ny, nz, 2, cz, n_threads);
    gaussian_pass<false>(out, tmp.data(), nx, ny, nz, 0, cx, n_threads);
    gaussian_pass<false>(tmp.data(), out, nx, ny, nz, 1, cy, n_threads);
  }
  return 0;
}

// A single line through the recursion (for unit tests of the kernel's streaming form).
void orc_gaussian_line(const double* c20, const double* data, double* outs, int ln, int arith) {
  std::vector<double> scratch(ln);
  if (arith)
    orc_line_fma(c20, data, outs, scratch.data(), ln);
  else
    orc_detail::filter_line<false>(c20, data, outs, scratch.data(), ln);
}

// itk::MultiplyImageFilter<float>: out = float(a*b).  [ITK-recalled]
void orc_multiply(const float* a, const float* b, float* out, size_t n, int n_threads) {
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
  for (long long i = 0; i < (long long)n; ++i) out[i] = a[i] * b[i];
}

// itk::DivideImageFilter<float>: b != 0 ? a/b : NumericTraits<float>::max().  [ITK-recalled]
void orc_divide(const float* a, const float* b, float* out, size_t n, int n_threads) {
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
  for (long long i = 0; i < (long long)n; ++i)
    out[i] = (b[i] != 0.0f) ? a[i] / b[i] : std::numeric_limits<float>::max();
}

// itk::MaskImageFilter with the default outside value 0: mask != 0 ? v : 0.  In place ok.
void orc_mask_u8(const float* v, const uint8_t* mask, float* out, size_t n, int n_threads) {
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
  for (long long i = 0; i < (long long)n; ++i) out[i] = mask[i] != 0 ? v[i] : 0.0f;
}
void orc_mask_f32(const float* v, const float* mask, float* out, size_t n, int n_threads) {
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
  for (long long i = 0; i < (long long)n; ++i) out[i] = mask[i] != 0.0f ? v[i] : 0.0f;
}

// itk::DerivativeImageFilter<float image -> float image>, UseImageSpacing = true:
// DerivativeOperator<float> of the given order along `axis` (radius 1), coefficients
// scaled ONCE by 1/spacing[axis] whatever the order and stored as float, flipped so the
// result is a convolution; ZeroFluxNeumann boundary (index clamped to the nearest valid
// voxel); inner product accumulated in double in neighbourhood order (low index first),
// result stored as float.  [ITK-recalled]
//   order 1: 0.5*(f[i+1] - f[i-1]) / spacing       order 2: (f[i-1] - 2 f[i] + f[i+1]) / spacing
void orc_derivative(const float* in, float* out, int nx, int ny, int nz, int axis, int order,
                    const double* spacing, int n_threads) {
  const double s = 1.0 / spacing[axis];
  double cm, c0, cp;  // coefficients applied to f[i-1], f[i], f[i+1]
  if (order == 1) {
    cm = (double)(float)(-0.5 * s);
    c0 = (double)(float)(0.0 * s);
    cp = (double)(float)(0.5 * s);
  } else {
    cm = (double)(float)(1.0 * s);
    c0 = (double)(float)(-2.0 * s);
    cp = (double)(float)(1.0 * s);
  }
  const int dims[3] = {nx, ny, nz};
  const size_t strides[3] = {1, (size_t)nx, (size_t)nx * ny};
  const int ln = dims[axis];
  const size_t st = strides[axis];
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int x = 0; x < nx; ++x) {
        const int p[3] = {x, y, z};
        const size_t idx = x + strides[1] * y + strides[2] * z;
        const int i = p[axis];
        const float fm = in[i > 0 ? idx - st : idx];
        const float f0 = in[idx];
        const float fp = in[i < ln - 1 ? idx + st : idx];
        double sum = 0.0;
        sum += cm * (double)fm;
        sum += c0 * (double)f0;
        sum += cp * (double)fp;
        out[idx] = static_cast<float>(sum);
      }
}

// itk::GradientMagnitudeImageFilter<float,float>, UseImageSpacing = true: per axis a
// DerivativeOperator<double> (order 1, radius 1) scaled by 1/spacing, ZeroFluxNeumann
// boundary, g_d accumulated in double; out = float(sqrt(sum_d g_d^2)).  [ITK-recalled]
void orc_gradient_magnitude(const float* in, float* out, int nx, int ny, int nz,
                            const double* spacing, int n_threads) {
  const size_t sy = (size_t)nx, sz = (size_t)nx * ny;
  const double cx = 0.5 * (1.0 / spacing[0]);
  const double cy = 0.5 * (1.0 / spacing[1]);
  const double cz = 0.5 * (1.0 / spacing[2]);
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int x = 0; x < nx; ++x) {
        const size_t idx = x + sy * y + sz * z;
        const double xm = in[x > 0 ? idx - 1 : idx], xp = in[x < nx - 1 ? idx + 1 : idx];
        const double ym = in[y > 0 ? idx - sy : idx], yp = in[y < ny - 1 ? idx + sy : idx];
        const double zm = in[z > 0 ? idx - sz : idx], zp = in[z < nz - 1 ? idx + sz : idx];
        const double f0 = in[idx];
        double a = 0.0;
        double g;
        g = 0.0; g += (-cx) * xm; g += 0.0 * f0; g += cx * xp; a += g * g;
        g = 0.0; g += (-cy) * ym; g += 0.0 * f0; g += cy * yp; a += g * g;
        g = 0.0; g += (-cz) * zm; g += 0.0 * f0; g += cz * zp; a += g * g;
        out[idx] = static_cast<float>(std::sqrt(a));
      }
}

}  // extern "C"
