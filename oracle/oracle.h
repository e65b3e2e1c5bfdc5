// TEST INFRASTRUCTURE ONLY.  C interface of the CPU oracle (oracle/liboracle.so): a
// restatement of the reference's dense per-voxel hot path used as the parity checker by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// The product library (libife_cuda.so) never includes, links or loads anything here.
//
// All volumes: float32, x fastest (idx = x + nx*(y + ny*z)); masks uint8; multi-component
// results are returned as SoA planes out[k*n + idx] unless a name says "interleaved".
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// ---- ITK primitives restated (oracle_itk.cpp) -- PARITY UNPINNED, see that file ----
void orc_gaussian_coefficients(double sigma, double spacing, double* c20);
void orc_gaussian_line(const double* c20, const double* data, double* outs, int ln, int arith);
int orc_smoothing_recursive_gaussian(const float* in, float* out, int nx, int ny, int nz,
                                     const double* spacing, double sigma, int arith,
                                     int n_threads);
void orc_multiply(const float* a, const float* b, float* out, size_t n, int n_threads);
void orc_divide(const float* a, const float* b, float* out, size_t n, int n_threads);
void orc_mask_u8(const float* v, const uint8_t* mask, float* out, size_t n, int n_threads);
void orc_mask_f32(const float* v, const float* mask, float* out, size_t n, int n_threads);
void orc_derivative(const float* in, float* out, int nx, int ny, int nz, int axis, int order,
                    const double* spacing, int n_threads);
void orc_gradient_magnitude(const float* in, float* out, int nx, int ny, int nz,
                            const double* spacing, int n_threads);

// ---- reference's own numerics restated (oracle_ife.cpp) -- pinned by the reference's
//      golden vectors and, bit for bit, by oracle/_ref ----
// math_mode: 0 = unqualified sqrt/acos/cos bind to the C double functions (what the
// reference header does when only <cmath> is visible; default everywhere), 1 = float
// overloads.
void orc_eig_f32(const float* A6, float* out3, size_t n, int math_mode);
void orc_eig_f64(const double* A6, double* out3, size_t n);
void orc_features_f32(const float* A6, float* out6, size_t n, int math_mode);
void orc_features_f64(const double* A6, double* out6, size_t n);
void orc_functor_volume_f32(const float* hess6_interleaved, const uint8_t* mask,
                            float* out6_interleaved, size_t n, int n_threads);
void orc_hist_f32(const float* edges, int n_edges, const float* values, size_t n,
                  uint32_t* counts, float* freqs);
int orc_determine_edges_f64(const double* sorted, size_t n, double* edges, size_t n_bins);
int orc_determine_edges_f32(const float* sorted, size_t n, float* edges, size_t n_bins);

// ---- compositions (oracle_ife.cpp), each citing the reference wiring it follows ----
void orc_hessian6(const float* in, float* hess6_interleaved, int nx, int ny, int nz,
                  const double* spacing, int fdhf_tool_bug, int n_threads);
int orc_normalized_gaussian(const float* img, const float* certainty, float* out, int nx, int ny,
                            int nz, const double* spacing, double sigma, int arith,
                            int n_threads);
int orc_emphysema_features(const float* img, const uint8_t* mask, float* out8, int nx, int ny,
                           int nz, const double* spacing, double sigma, int arith,
                           int n_threads);
int orc_fd_hessian_features(const float* img, const uint8_t* mask, float* out6, int nx, int ny,
                            int nz, const double* spacing, double sigma, int arith,
                            int fdhf_tool_bug, int n_threads);
void orc_fd_gradient_features(const float* img, const float* mask, float* out, int nx, int ny,
                              int nz, const double* spacing, int n_threads);
void orc_features_histograms(const float* feats, int n_feat, const uint8_t* mask, int nx, int ny,
                             int nz, const int* roi_boxes, int n_roi, const float* edges,
                             int n_edges, uint32_t* counts);

#ifdef __cplusplus
}
#endif
#endif
