// TEST INFRASTRUCTURE ONLY (oracle).  The per-line recursion of ITK's recursive Gaussian,
// shared by oracle_itk.cpp (PLAIN arithmetic) and oracle_itk_fma.cpp (FMA arithmetic, the
// only translation unit compiled with -mfma).  See oracle_itk.cpp for provenance
// ([ITK-recalled], parity unpinned).
#ifndef ORACLE_LINE_H
#define ORACLE_LINE_H
namespace orc_detail {

inline double fma_(double a, double b, double c) { return __builtin_fma(a, b, c); }

// sum of four products, left to right.  PLAIN: ((a0*c0 + a1*c1) + a2*c2) + a3*c3.
template <bool FMA>
inline double dot4(double a0, double c0, double a1, double c1, double a2, double c2, double a3,
                   double c3);
template <>
inline double dot4<false>(double a0, double c0, double a1, double c1, double a2, double c2,
                          double a3, double c3) {
  return a0 * c0 + a1 * c1 + a2 * c2 + a3 * c3;
}
template <>
inline double dot4<true>(double a0, double c0, double a1,
                                                        double c1, double a2, double c2,
                                                        double a3, double c3) {
  double t = a0 * c0;
  t = fma_(a1, c1, t);
  t = fma_(a2, c2, t);
  t = fma_(a3, c3, t);
  return t;
}

// [ITK-recalled] RecursiveSeparableImageFilter::FilterDataArray: one line of `ln` (>= 4)
// samples, all in double.  outs[i] = causal[i] + anticausal[i].
template <bool FMA>
void filter_line(const double* c, const double* data, double* outs, double* scratch, int ln) {
  const double N0 = c[0], N1 = c[1], N2 = c[2], N3 = c[3];
  const double D1 = c[4], D2 = c[5], D3 = c[6], D4 = c[7];
  const double M1 = c[8], M2 = c[9], M3 = c[10], M4 = c[11];
  const double BN1 = c[12], BN2 = c[13], BN3 = c[14], BN4 = c[15];
  const double BM1 = c[16], BM2 = c[17], BM3 = c[18], BM4 = c[19];

  // causal pass; the first sample is assumed to extend to -infinity
  const double v1 = data[0];
  scratch[0] = dot4<FMA>(v1, N0, v1, N1, v1, N2, v1, N3);
  scratch[1] = dot4<FMA>(data[1], N0, v1, N1, v1, N2, v1, N3);
  scratch[2] = dot4<FMA>(data[2], N0, data[1], N1, v1, N2, v1, N3);
  scratch[3] = dot4<FMA>(data[3], N0, data[2], N1, data[1], N2, v1, N3);
  scratch[0] -= dot4<FMA>(v1, BN1, v1, BN2, v1, BN3, v1, BN4);
  scratch[1] -= dot4<FMA>(scratch[0], D1, v1, BN2, v1, BN3, v1, BN4);
  scratch[2] -= dot4<FMA>(scratch[1], D1, scratch[0], D2, v1, BN3, v1, BN4);
  scratch[3] -= dot4<FMA>(scratch[2], D1, scratch[1], D2, scratch[0], D3, v1, BN4);
  for (int i = 4; i < ln; ++i) {
    scratch[i] = dot4<FMA>(data[i], N0, data[i - 1], N1, data[i - 2], N2, data[i - 3], N3);
    scratch[i] -=
        dot4<FMA>(scratch[i - 1], D1, scratch[i - 2], D2, scratch[i - 3], D3, scratch[i - 4], D4);
  }
  for (int i = 0; i < ln; ++i) outs[i] = scratch[i];

  // anticausal pass; the last sample is assumed to extend to +infinity
  const double v2 = data[ln - 1];
  scratch[ln - 1] = dot4<FMA>(v2, M1, v2, M2, v2, M3, v2, M4);
  scratch[ln - 2] = dot4<FMA>(data[ln - 1], M1, v2, M2, v2, M3, v2, M4);
  scratch[ln - 3] = dot4<FMA>(data[ln - 2], M1, data[ln - 1], M2, v2, M3, v2, M4);
  scratch[ln - 4] = dot4<FMA>(data[ln - 3], M1, data[ln - 2], M2, data[ln - 1], M3, v2, M4);
  scratch[ln - 1] -= dot4<FMA>(v2, BM1, v2, BM2, v2, BM3, v2, BM4);
  scratch[ln - 2] -= dot4<FMA>(scratch[ln - 1], D1, v2, BM2, v2, BM3, v2, BM4);
  scratch[ln - 3] -= dot4<FMA>(scratch[ln - 2], D1, scratch[ln - 1], D2, v2, BM3, v2, BM4);
  scratch[ln - 4] -=
      dot4<FMA>(scratch[ln - 3], D1, scratch[ln - 2], D2, scratch[ln - 1], D3, v2, BM4);
  for (int i = ln - 4; i > 0; --i) {
    scratch[i - 1] = dot4<FMA>(data[i], M1, data[i + 1], M2, data[i + 2], M3, data[i + 3], M4);
    scratch[i - 1] -=
        dot4<FMA>(scratch[i], D1, scratch[i + 1], D2, scratch[i + 2], D3, scratch[i + 3], D4);
  }
  for (int i = 0; i < ln; ++i) outs[i] += scratch[i];
}

}  // namespace orc_detail
#endif
