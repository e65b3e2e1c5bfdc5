// Shared by the fused feature kernel and the Gaussian passes.
#pragma once
#include <cuda_runtime.h>

namespace ife {

// ---------------------------------------------------------------------------------------
// IEEE-correct float divisions with the slow paths hoisted out.  q = RN(a/b) from a
// correctly rounded reciprocal y = RN(1/b):  q0 = RN(a*y); then twice r = a - b*q (exact in
// an FMA), q = RN(q + r*y)  (Markstein).  Valid while nothing underflows or overflows, which the
// range tests guarantee; anything else takes the compiler's full division.  The solver
// divides six numbers by the same p, so one reciprocal serves all six.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool div_safe_num(float a) {
  const float m = fabsf(a);
  return m == 0.0f || (m > 0x1p-60f && m < 0x1p60f);
}
__device__ __forceinline__ float div_with_rcp(float a, float b, float y) {
  // q0 can be 1.5 ulp off; the first correction makes it faithful (< 1 ulp), which is the
  // precondition under which the second correction is provably the correctly rounded a/b
  const float q0 = __fmul_rn(a, y);
  const float q1 = __fmaf_rn(__fmaf_rn(-b, q0, a), y, q0);
  return __fmaf_rn(__fmaf_rn(-b, q1, a), y, q1);
}
}  // namespace ife
