// TEST INFRASTRUCTURE ONLY (oracle).  FMA-arithmetic instantiation of the recursive
// Gaussian line filter; the only oracle file compiled with -mfma (and, like the rest,
// -ffp-contract=off, so the ONLY fused operations are the explicit __builtin_fma calls in
// oracle_line.h).  See oracle_itk.cpp for provenance.
#include "oracle_line.h"

extern "C" void orc_line_fma(const double* c20, const double* data, double* outs, double* scratch,
                             int ln) {
  orc_detail::filter_line<true>(c20, data, outs, scratch, ln);
}
