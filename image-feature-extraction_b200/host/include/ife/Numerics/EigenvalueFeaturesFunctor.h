// Host-side mirrors of Symmetric3x3EigenvalueSolver<float> and EigenvalueFeaturesFunctor<float>
// (reference include/ife/Numerics/Symmetric3x3EigenvalueSolver.h:11-133,
// EigenvalueFeaturesFunctor.h:9-32).  Input layout [A11,A12,A13,A22,A23,A33]; eigenvalues
// come back ordered |e1| >= |e2| >= |e3|.  The per-matrix operator() of the reference is
// kept for API compatibility, but the GPU is meant to be fed batches (operator()(A6, out, n)).
#ifndef IFE_B200_EIGENVALUE_FEATURES_FUNCTOR_H
#define IFE_B200_EIGENVALUE_FEATURES_FUNCTOR_H
#include <cstddef>
#include <vector>

#include "ife/Context.h"

namespace ife {

template <typename TRealType = float>
struct EigenvalueFeaturesFunctor {
  typedef TRealType RealType;
  typedef std::vector<RealType> InputType;
  typedef std::vector<RealType> OutputType;
  bool operator!=(const EigenvalueFeaturesFunctor&) const { return false; }
  bool operator==(const EigenvalueFeaturesFunctor& o) const { return !(*this != o); }

  // n interleaved matrices -> n x 6 features
  void operator()(const float* A6, float* out6, size_t n) const {
    CudaContext& c = CudaContext::Instance();
    c.Check(ife_cuda_eigen_features_batch(c.Handle(), A6, out6, n, IFE_MEM_HOST));
  }
  OutputType operator()(const InputType& A) const {
    if (A.size() != 6) throw ExceptionObject(IFE_E_INVALID, "EigenvalueFeaturesFunctor: need 6 matrix entries");
    OutputType f(6);
    (*this)(A.data(), f.data(), 1);
    return f;
  }
};

template <typename TRealType = float>
struct Symmetric3x3EigenvalueSolver {
  typedef TRealType RealType;
  typedef std::vector<RealType> InputType;
  typedef std::vector<RealType> OutputType;
  bool operator!=(const Symmetric3x3EigenvalueSolver&) const { return false; }
  bool operator==(const Symmetric3x3EigenvalueSolver& o) const { return !(*this != o); }
  OutputType operator()(const InputType& A) const {
    OutputType f = EigenvalueFeaturesFunctor<TRealType>()(A);
    f.resize(3);
    return f;
  }
};

}  // namespace ife
#endif
