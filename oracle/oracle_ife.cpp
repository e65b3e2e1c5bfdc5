// TEST INFRASTRUCTURE ONLY (oracle).  CPU restatement of the reference's OWN numerics
// (eigen solver, eigen features, DenseHistogram, equalized edges) and of the filter
// compositions that wire the ITK primitives of oracle_itk.cpp together.  Used only as
// the parity checker / CPU baseline; the product never links or loads it.
//
// Pinning: the solver, functor, histogram and edge functions here are checked in
// tests/test_oracle.py against every golden vector of the reference's three test
// programs (test/Symmetric3x3EigenvalueSolverTest.cxx:48-90, test/DenseHistogramTest.cxx:10-55,
// test/DetermineEdgesForEqualizedHistogramTest.cxx:30-120) and, bit for bit on random
// inputs, against the reference headers themselves compiled into oracle/_ref.  The
// compositions follow the reference's .hxx wiring line by line but rest on the
// ITK-recalled primitives, so the image-level path is PARITY UNPINNED (see oracle_itk.cpp).
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

const double kPi = 3.14159265358979323846;  // M_PI

// Reference: include/ife/Numerics/Symmetric3x3EigenvalueSolver.h:33-132.
// Input [A11,A12,A13,A22,A23,A33]; output sorted so that |e0| >= |e1| >= |e2|.
// MATH_DOUBLE: sqrt/acos/cos are the C double functions applied to promoted arguments
// and every expression they appear in is evaluated in double before being narrowed back
// to T on assignment (what the header does for T=float when only <cmath> is visible).
template <typename T, bool MATH_DOUBLE>
inline void solve3(const T* A, T* ev) {
  const T A11 = A[0], A12 = A[1], A13 = A[2], A22 = A[3], A23 = A[4], A33 = A[5];
  T p = A12 * A12 + A13 * A13 + A23 * A23;
  if (p == 0) {
    // diagonal matrix (:44-83): order the diagonal by decreasing magnitude with the
    // reference's strict '>' comparisons (ties fall to the later branch).
    const T a1 = std::abs(A11), a2 = std::abs(A22), a3 = std::abs(A33);
    if (a1 > a2) {
      if (a1 > a3) {
        ev[0] = A11;
        if (a2 > a3) { ev[1] = A22; ev[2] = A33; } else { ev[1] = A33; ev[2] = A22; }
      } else {
        ev[0] = A33; ev[1] = A11; ev[2] = A22;
      }
    } else {
      if (a2 > a3) {
        ev[0] = A22;
        if (a1 > a3) { ev[1] = A11; ev[2] = A33; } else { ev[1] = A33; ev[2] = A11; }
      } else {
        ev[0] = A33; ev[1] = A22; ev[2] = A11;
      }
    }
    return;
  }
  const T q = (A11 + A22 + A33) / 3;                                             // :85
  p = (A11 - q) * (A11 - q) + (A22 - q) * (A22 - q) + (A33 - q) * (A33 - q) + 2 * p;  // :86-87
  if (MATH_DOUBLE) p = static_cast<T>(std::sqrt(static_cast<double>(p / 6)));    // :88
  else p = std::sqrt(p / 6);
  const T B11 = (A11 - q) / p, B12 = A12 / p, B13 = A13 / p;                     // :92-97
  const T B22 = (A22 - q) / p, B23 = A23 / p, B33 = (A33 - q) / p;
  // :98-103 -- the bracket is T arithmetic, the final "/ 2.0" is a double division
  const T r = static_cast<T>(static_cast<double>(B11 * B22 * B33 + 2 * B12 * B13 * B23 -
                                                 B23 * B23 * B11 - B13 * B13 * B22 -
                                                 B12 * B12 * B33) / 2.0);
  T phi;                                                                         // :107-116
  if (r <= -1) phi = static_cast<T>(kPi / 3);
  else if (r >= 1) phi = 0;
  else if (MATH_DOUBLE) phi = static_cast<T>(std::acos(static_cast<double>(r)) / 3);
  else phi = std::acos(r) / 3;
  T e0, e2;                                                                      // :119-120
  if (MATH_DOUBLE)
    e0 = static_cast<T>(static_cast<double>(q) +
                        static_cast<double>(2 * p) * std::cos(static_cast<double>(phi)));
  else
    e0 = q + 2 * p * std::cos(phi);
  // phi + M_PI*(2.0/3.0) is a double expression whatever T is
  e2 = static_cast<T>(static_cast<double>(q) +
                      static_cast<double>(2 * p) *
                          std::cos(static_cast<double>(phi) + kPi * (2.0 / 3.0)));
  T e1 = 3 * q - e0 - e2;                                                        // :121
  if (std::abs(e0) < std::abs(e2)) std::swap(e0, e2);                            // :123-125
  if (std::abs(e1) < std::abs(e2)) std::swap(e1, e2);                            // :127-129
  ev[0] = e0; ev[1] = e1; ev[2] = e2;
}

// Reference: include/ife/Numerics/EigenvalueFeaturesFunctor.h:20-31.
template <typename T, bool MATH_DOUBLE>
inline void features6(const T* A, T* f) {
  T ev[3];
  solve3<T, MATH_DOUBLE>(A, ev);
  f[0] = ev[0];
  f[1] = ev[1];
  f[2] = ev[2];
  f[3] = ev[0] + ev[1] + ev[2];
  f[4] = ev[0] * ev[1] * ev[2];
  f[5] = std::sqrt(ev[0] * ev[0] + ev[1] * ev[1] + ev[2] * ev[2]);  // std::sqrt: T overload
}

// Reference: include/ife/Statistics/DenseHistogram.h:47-53.  bin = number of edges
// strictly less than v  (bins (-inf,e0], (e0,e1], ..., (e_{n-1},inf); NaN -> bin 0).
inline int bin_of(const float* edges, int n_edges, float v) {
  return static_cast<int>(std::lower_bound(edges, edges + n_edges, v) - edges);
}

// Reference: include/ife/Statistics/DetermineEdgesForEqualizedHistogram.h:21-139,
// restated with indices instead of iterators.
template <typename T>
int determine_edges(const T* v, size_t n, T* edges, size_t n_bins) {
  if (n < n_bins) return 1;  // std::out_of_range in the reference (:37-39)
  const size_t per_bin = n / n_bins;
  size_t surplus = n - per_bin * n_bins, deficit = 0, n_edge = 0, pos = 0;
  while (n_edge + 1 < n_bins) {
    size_t step = per_bin;
    if (surplus) {
      size_t s = surplus / (n_bins - n_edge);
      if (s == 0) s = 1;
      step += s;
      surplus -= s;
    } else if (deficit) {
      size_t d = deficit / (n_bins - n_edge);
      if (d == 0) d = 1;
      step -= d;
      deficit -= d;
    }
    pos += step;
    const size_t lb = std::lower_bound(v, v + pos, v[pos]) - v;
    if (lb != pos) {
      const size_t ub = std::upper_bound(v + pos, v + n, v[pos]) - v;
      if (ub == n) {
        pos = lb;
      } else {
        const size_t lbdist = pos - lb, ubdist = ub - pos;
        if (lbdist < ubdist || (lbdist == ubdist && deficit)) {
          pos = lb;
          if (lbdist > deficit) { surplus = lbdist - deficit; deficit = 0; }
          else deficit -= lbdist;
        } else {
          pos = ub;
          if (ubdist > surplus) { deficit = ubdist - surplus; surplus = 0; }
          else surplus -= ubdist;
        }
      }
    }
    edges[n_edge++] = v[pos];
  }
  return 0;
}

}  // namespace

extern "C" {

void orc_eig_f32(const float* A6, float* out3, size_t n, int math_mode) {
  for (size_t i = 0; i < n; ++i) {
    if (math_mode == 0) solve3<float, true>(A6 + 6 * i, out3 + 3 * i);
    else solve3<float, false>(A6 + 6 * i, out3 + 3 * i);
  }
}
void orc_eig_f64(const double* A6, double* out3, size_t n) {
  for (size_t i = 0; i < n; ++i) solve3<double, true>(A6 + 6 * i, out3 + 3 * i);
}
void orc_features_f32(const float* A6, float* out6, size_t n, int math_mode) {
  for (size_t i = 0; i < n; ++i) {
    if (math_mode == 0) features6<float, true>(A6 + 6 * i, out6 + 6 * i);
    else features6<float, false>(A6 + 6 * i, out6 + 6 * i);
  }
}
void orc_features_f64(const double* A6, double* out6, size_t n) {
  for (size_t i = 0; i < n; ++i) features6<double, true>(A6 + 6 * i, out6 + 6 * i);
}

// UnaryFunctorImageFilter<EigenvalueFeaturesFunctor<float>> over an interleaved Hessian
// (ImageToEmphysemaFeaturesFilter.hxx:33-35), or with a mask the masked in-place loop of
// tools/FiniteDifference_HessianFeatures.cxx:209-229 (mask == 0 -> six zeros).
void orc_functor_volume_f32(const float* hess6, const uint8_t* mask, float* out6, size_t n,
                            int n_threads) {
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
  for (long long i = 0; i < (long long)n; ++i) {
    if (mask && mask[i] == 0) {
      for (int k = 0; k < 6; ++k) out6[6 * i + k] = 0.0f;
    } else {
      float h[6], f[6];
      std::memcpy(h, hess6 + 6 * i, sizeof(h));
      features6<float, true>(h, f);
      std::memcpy(out6 + 6 * i, f, sizeof(f));
    }
  }
}

// DenseHistogram<float>: insert n values; counts (unsigned) and frequencies.
// getFrequencies (DenseHistogram.h:55-60) accumulates the counts into an int, converts
// it to float and divides float(count) by it.
void orc_hist_f32(const float* edges, int n_edges, const float* values, size_t n,
                  uint32_t* counts, float* freqs) {
  std::vector<uint32_t> c(n_edges + 1, 0u);
  for (size_t i = 0; i < n; ++i) ++c[bin_of(edges, n_edges, values[i])];
  if (counts) std::copy(c.begin(), c.end(), counts);
  if (freqs) {
    int isum = 0;
    for (uint32_t x : c) isum += (int)x;
    const float sum = (float)isum;
    for (int b = 0; b <= n_edges; ++b) freqs[b] = (float)c[b] / sum;
  }
}

int orc_determine_edges_f64(const double* sorted, size_t n, double* edges, size_t n_bins) {
  return determine_edges(sorted, n, edges, n_bins);
}
int orc_determine_edges_f32(const float* sorted, size_t n, float* edges, size_t n_bins) {
  return determine_edges(sorted, n, edges, n_bins);
}

// Hessian3DImageFilter (include/ife/Filters/Hessian3DImageFilter.hxx:11-60): eight chained
// DerivativeImageFilters; cross terms are first-order filters applied to the float output
// of a first-order filter; components composed as [Dxx,Dxy,Dxz,Dyy,Dyz,Dzz] (:53-59).
// fdhf_tool_bug != 0 reproduces tools/FiniteDifference_HessianFeatures.cxx:153-156, whose
// "dy" filter is set to direction 0, so that its Dyz is really Dz(Dx).
void orc_hessian6(const float* in, float* hess6, int nx, int ny, int nz, const double* spacing,
                  int fdhf_tool_bug, int n_threads) {
  const size_t n = (size_t)nx * ny * nz;
  std::vector<float> comp(n), d1x(n), d1y(n);
  auto put = [&](int k) {
    for (size_t i = 0; i < n; ++i) hess6[6 * i + k] = comp[i];
  };
  orc_derivative(in, comp.data(), nx, ny, nz, 0, 2, spacing, n_threads); put(0);  // Dxx
  orc_derivative(in, comp.data(), nx, ny, nz, 1, 2, spacing, n_threads); put(3);  // Dyy
  orc_derivative(in, comp.data(), nx, ny, nz, 2, 2, spacing, n_threads); put(5);  // Dzz
  orc_derivative(in, d1x.data(), nx, ny, nz, 0, 1, spacing, n_threads);           // Dx
  orc_derivative(in, d1y.data(), nx, ny, nz, fdhf_tool_bug ? 0 : 1, 1, spacing, n_threads);  // Dy
  orc_derivative(d1x.data(), comp.data(), nx, ny, nz, 1, 1, spacing, n_threads); put(1);  // Dxy
  orc_derivative(d1x.data(), comp.data(), nx, ny, nz, 2, 1, spacing, n_threads); put(2);  // Dxz
  orc_derivative(d1y.data(), comp.data(), nx, ny, nz, 2, 1, spacing, n_threads); put(4);  // Dyz
}

// NormalizedGaussianConvolutionImageFilter::GenerateData
// (include/ife/Filters/NormalizedGaussianConvolutionImageFilter.hxx:37-63):
// out = G(c*T) / G(c), G = SmoothingRecursiveGaussianImageFilter(sigma).
int orc_normalized_gaussian(const float* img, const float* certainty, float* out, int nx, int ny,
                            int nz, const double* spacing, double sigma, int arith,
                            int n_threads) {
  const size_t n = (size_t)nx * ny * nz;
  std::vector<float> ct(n), g1(n), g2(n);
  orc_multiply(img, certainty, ct.data(), n, n_threads);                              // :48-49
  int rc = orc_smoothing_recursive_gaussian(ct.data(), g1.data(), nx, ny, nz, spacing, sigma,
                                            arith, n_threads);                        // :51,54
  if (rc) return rc;
  rc = orc_smoothing_recursive_gaussian(certainty, g2.data(), nx, ny, nz, spacing, sigma, arith,
                                        n_threads);                                   // :52,55
  if (rc) return rc;
  orc_divide(g1.data(), g2.data(), out, n, n_threads);                                // :57-58
  return 0;
}

// ImageToEmphysemaFeaturesFilter (include/ife/Filters/ImageToEmphysemaFeaturesFilter.hxx:
// 11-55 wiring, :94-121 GenerateData).  out8 = SoA planes
// [Blur, GradMag, e1, e2, e3, LoG, Curv, Frob] (tools/ExtractFeatures.cxx:126-130).
int orc_emphysema_features(const float* img, const uint8_t* mask, float* out8, int nx, int ny,
                           int nz, const double* spacing, double sigma, int arith,
                           int n_threads) {
  const size_t n = (size_t)nx * ny * nz;
  std::vector<float> maskf(n), blur(n), gm(n), hess(6 * n), feat(6 * n);
  for (size_t i = 0; i < n; ++i) maskf[i] = static_cast<float>(mask[i]);   // CastImageFilter :21,110
  int rc = orc_normalized_gaussian(img, maskf.data(), blur.data(), nx, ny, nz, spacing, sigma,
                                   arith, n_threads);                       // :24-25,111-112
  if (rc) return rc;
  orc_gradient_magnitude(blur.data(), gm.data(), nx, ny, nz, spacing, n_threads);  // :27-28
  orc_hessian6(blur.data(), hess.data(), nx, ny, nz, spacing, 0, n_threads);       // :30-31
  orc_functor_volume_f32(hess.data(), nullptr, feat.data(), n, n_threads);         // :33-35
  orc_mask_u8(blur.data(), mask, out8, n, n_threads);                              // :44-54
  orc_mask_u8(gm.data(), mask, out8 + n, n, n_threads);
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
  for (long long i = 0; i < (long long)n; ++i)
    for (int k = 0; k < 6; ++k) out8[(size_t)(2 + k) * n + i] = mask[i] != 0 ? feat[6 * i + k] : 0.0f;
  return 0;
}

// tools/FiniteDifference_HessianFeatures.cxx:127-229: un-smoothed Hessian of the image,
// per-voxel eigen features where mask != 0, six zeros elsewhere; out6 = SoA planes
// [eig1, eig2, eig3, LoG, Curvature, Frobenius] (:255-257).  sigma > 0 first smooths with
// the library Gaussian (BASELINE.json configs[0] "single sigma=1.0"); sigma <= 0 is the
// tool as shipped.  mask may be null (all inside).
int orc_fd_hessian_features(const float* img, const uint8_t* mask, float* out6, int nx, int ny,
                            int nz, const double* spacing, double sigma, int arith,
                            int fdhf_tool_bug, int n_threads) {
  const size_t n = (size_t)nx * ny * nz;
  std::vector<float> smooth, hess(6 * n), feat(6 * n);
  const float* src = img;
  if (sigma > 0) {
    smooth.resize(n);
    int rc = orc_smoothing_recursive_gaussian(img, smooth.data(), nx, ny, nz, spacing, sigma,
                                              arith, n_threads);
    if (rc) return rc;
    src = smooth.data();
  }
  orc_hessian6(src, hess.data(), nx, ny, nz, spacing, fdhf_tool_bug, n_threads);
  orc_functor_volume_f32(hess.data(), mask, feat.data(), n, n_threads);
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
  for (long long i = 0; i < (long long)n; ++i)
    for (int k = 0; k < 6; ++k) out6[(size_t)k * n + i] = feat[6 * i + k];
  return 0;
}

// tools/FiniteDifference_GradientFeatures.cxx:105-113: gradient magnitude of the raw image,
// masked by a mask that the tool reads as float.
void orc_fd_gradient_features(const float* img, const float* mask, float* out, int nx, int ny,
                              int nz, const double* spacing, int n_threads) {
  const size_t n = (size_t)nx * ny * nz;
  orc_gradient_magnitude(img, out, nx, ny, nz, spacing, n_threads);
  if (mask) orc_mask_f32(out, mask, out, n, n_threads);
}

// The insert loop of tools/MakeBag.cxx:425-470: for every ROI box {x0,y0,z0,sx,sy,sz} and
// every voxel of it with mask != 0, insert feature k into histogram k.  feats = SoA planes
// [n_feat][n]; edges [n_feat][n_edges]; counts [n_roi][n_feat][n_edges+1].  n_roi == 0
// means one ROI covering the whole volume (counts [1][n_feat][n_edges+1]).  Single-
// threaded like the reference loop.
void orc_features_histograms(const float* feats, int n_feat, const uint8_t* mask, int nx, int ny,
                             int nz, const int* roi_boxes, int n_roi, const float* edges,
                             int n_edges, uint32_t* counts) {
  const size_t n = (size_t)nx * ny * nz;
  const int whole[6] = {0, 0, 0, nx, ny, nz};
  const int rois = n_roi > 0 ? n_roi : 1;
  std::fill(counts, counts + (size_t)rois * n_feat * (n_edges + 1), 0u);
  for (int r = 0; r < rois; ++r) {
    const int* b = n_roi > 0 ? roi_boxes + 6 * r : whole;
    uint32_t* c = counts + (size_t)r * n_feat * (n_edges + 1);
    for (int z = b[2]; z < b[2] + b[5]; ++z)
      for (int y = b[1]; y < b[1] + b[4]; ++y)
        for (int x = b[0]; x < b[0] + b[3]; ++x) {
          const size_t idx = x + (size_t)nx * (y + (size_t)ny * z);
          if (mask && mask[idx] == 0) continue;
          for (int k = 0; k < n_feat; ++k)
            ++c[(size_t)k * (n_edges + 1) + bin_of(edges + (size_t)k * n_edges, n_edges,
                                                   feats[(size_t)k * n + idx])];
        }
  }
}

}  // extern "C"
