// TEST INFRASTRUCTURE ONLY (oracle/_ref).  Not part of the product; never linked into
// libife_cuda.so.  This translation unit compiles the reference's own ITK-light headers
// UNMODIFIED, from where they lie under /root/reference/include, against the tiny
// itk::VariableLengthVector stand-in in oracle/itk_shim/, and exports them through a flat
// C interface so that tests can pin oracle/oracle_ife.cpp (the restatement) against the
// real reference code.  Built only in the authoring container (where /root/reference is
// mounted) into oracle/_ref/libife_ref.so; the .so travels to the GPU box.
//
// Reference code reached from here:
//   include/ife/Numerics/Symmetric3x3EigenvalueSolver.h:33-132   (ref_eig_*)
//   include/ife/Numerics/EigenvalueFeaturesFunctor.h:20-31       (ref_features_*)
//   include/ife/Statistics/DenseHistogram.h:47-64                (ref_hist_*)
//   include/ife/Statistics/DetermineEdgesForEqualizedHistogram.h:21-139 (ref_determine_edges)
//
// Only <cmath> is visible when the solver header is parsed (as in the header itself,
// Symmetric3x3EigenvalueSolver.h:5), so its unqualified sqrt/acos/cos bind to the C
// double functions; ref_math_overload_is_double() reports what this build did.
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "ife/Numerics/Symmetric3x3EigenvalueSolver.h"
#include "ife/Numerics/EigenvalueFeaturesFunctor.h"
#include "ife/Statistics/DenseHistogram.h"
#include "ife/Statistics/DetermineEdgesForEqualizedHistogram.h"

namespace {
// What does an unqualified acos(float) resolve to in this TU?  (Same lookup context as
// the reference header: global namespace, after <cmath>.)
constexpr bool kAcosIsDouble = std::is_same<decltype(acos(1.0f)), double>::value;

template <typename T>
void eig_batch(const T* A6, T* out3, size_t n) {
  Symmetric3x3EigenvalueSolver<T> solver;
  for (size_t i = 0; i < n; ++i) {
    typename Symmetric3x3EigenvalueSolver<T>::InputType A(A6 + 6 * i, 6);
    auto ev = solver(A);
    out3[3 * i + 0] = ev[0];
    out3[3 * i + 1] = ev[1];
    out3[3 * i + 2] = ev[2];
  }
}

template <typename T>
void feat_batch(const T* A6, T* out6, size_t n) {
  EigenvalueFeaturesFunctor<T> functor;
  for (size_t i = 0; i < n; ++i) {
    typename EigenvalueFeaturesFunctor<T>::InputType A(A6 + 6 * i, 6);
    auto f = functor(A);
    for (int k = 0; k < 6; ++k) out6[6 * i + k] = f[k];
  }
}
}  // namespace

extern "C" {

int ref_math_overload_is_double() { return kAcosIsDouble ? 1 : 0; }

void ref_eig_f32(const float* A6, float* out3, size_t n) { eig_batch<float>(A6, out3, n); }
void ref_eig_f64(const double* A6, double* out3, size_t n) { eig_batch<double>(A6, out3, n); }
void ref_features_f32(const float* A6, float* out6, size_t n) { feat_batch<float>(A6, out6, n); }
void ref_features_f64(const double* A6, double* out6, size_t n) { feat_batch<double>(A6, out6, n); }

// One DenseHistogram<float>: insert n values, return counts and frequencies.
void ref_hist_f32(const float* edges, int n_edges, const float* values, size_t n,
                  uint32_t* counts, float* freqs) {
  DenseHistogram<float> hist(edges, edges + n_edges);
  for (size_t i = 0; i < n; ++i) hist.insert(values[i]);
  auto c = hist.getCounts();
  auto f = hist.getFrequencies();
  for (size_t b = 0; b < c.size(); ++b) {
    if (counts) counts[b] = c[b];
    if (freqs) freqs[b] = f[b];
  }
}

// returns 0 ok, 1 std::out_of_range, 2 std::logic_error
int ref_determine_edges_f64(const double* sorted, size_t n, double* edges, size_t n_bins) {
  try {
    determineEdgesForEqualizedHistogram(sorted, sorted + n, edges, n_bins);
  } catch (const std::out_of_range&) {
    return 1;
  } catch (const std::logic_error&) {
    return 2;
  }
  return 0;
}
int ref_determine_edges_f32(const float* sorted, size_t n, float* edges, size_t n_bins) {
  try {
    determineEdgesForEqualizedHistogram(sorted, sorted + n, edges, n_bins);
  } catch (const std::out_of_range&) {
    return 1;
  } catch (const std::logic_error&) {
    return 2;
  }
  return 0;
}

// The per-voxel masked eigen loop of tools/FiniteDifference_HessianFeatures.cxx:209-229 /
// the UnaryFunctorImageFilter of ImageToEmphysemaFeaturesFilter.hxx:33-35, run with the
// reference functor (float) over an interleaved 6-component Hessian buffer.  mask may be
// null (all voxels inside).  Used by the CPU baseline so that the per-voxel cost (two
// VariableLengthVector heap allocations per call) is the reference's own.
void ref_functor_volume_f32(const float* hess6, const uint8_t* mask, float* out6, size_t n,
                            int n_threads) {
  EigenvalueFeaturesFunctor<float> functor;
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
  for (long long i = 0; i < (long long)n; ++i) {
    if (mask && mask[i] == 0) {
      for (int k = 0; k < 6; ++k) out6[6 * i + k] = 0.0f;
    } else {
      EigenvalueFeaturesFunctor<float>::InputType A(hess6 + 6 * i, 6);
      auto f = functor(A);
      for (int k = 0; k < 6; ++k) out6[6 * i + k] = f[k];
    }
  }
}

}  // extern "C"
