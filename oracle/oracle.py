"""TEST INFRASTRUCTURE ONLY: ctypes view of the CPU oracle (oracle/liboracle.so) and of the
reference's own headers compiled into oracle/_ref/libife_ref.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (image-feature-extraction_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")

FEATURE_NAMES8 = ["GaussianBlur", "GradientMagnitude", "Eigenvalue1", "Eigenvalue2",
                  "Eigenvalue3", "LaplacianOfGaussian", "GaussianCurvature", "FrobeniusNorm"]
ARITH_PLAIN, ARITH_FMA = 0, 1


def build(force=False):
    """Compile liboracle.so (and _ref when /root/reference is mounted)."""
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so) or os.path.isdir("/root/reference/include/ife"):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)
    return so


def _opt(arr, ptr_type):
    return None if arr is None else arr.ctypes.data_as(ptr_type)


class _Lib:
    def __init__(self):
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = self.L = C.CDLL(so)
        L.orc_gaussian_coefficients.argtypes = [C.c_double, C.c_double, _f64p]
        L.orc_gaussian_line.argtypes = [_f64p, _f64p, _f64p, C.c_int, C.c_int]
        L.orc_smoothing_recursive_gaussian.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int,
                                                       _f64p, C.c_double, C.c_int, C.c_int]
        L.orc_derivative.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     _f64p, C.c_int]
        L.orc_gradient_magnitude.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _f64p, C.c_int]
        L.orc_eig_f32.argtypes = [_f32p, _f32p, C.c_size_t, C.c_int]
        L.orc_eig_f64.argtypes = [_f64p, _f64p, C.c_size_t]
        L.orc_features_f32.argtypes = [_f32p, _f32p, C.c_size_t, C.c_int]
        L.orc_features_f64.argtypes = [_f64p, _f64p, C.c_size_t]
        L.orc_functor_volume_f32.argtypes = [_f32p, C.c_void_p, _f32p, C.c_size_t, C.c_int]
        L.orc_hist_f32.argtypes = [_f32p, C.c_int, _f32p, C.c_size_t, _u32p, _f32p]
        L.orc_determine_edges_f64.argtypes = [_f64p, C.c_size_t, _f64p, C.c_size_t]
        L.orc_determine_edges_f32.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t]
        L.orc_hessian6.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _f64p, C.c_int, C.c_int]
        L.orc_normalized_gaussian.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _f64p,
                                              C.c_double, C.c_int, C.c_int]
        L.orc_emphysema_features.argtypes = [_f32p, _u8p, _f32p, C.c_int, C.c_int, C.c_int, _f64p,
                                             C.c_double, C.c_int, C.c_int]
        L.orc_fd_hessian_features.argtypes = [_f32p, C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int,
                                              _f64p, C.c_double, C.c_int, C.c_int, C.c_int]
        L.orc_fd_gradient_features.argtypes = [_f32p, C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int,
                                               _f64p, C.c_int]
        L.orc_features_histograms.argtypes = [_f32p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_int, _f32p, C.c_int, _u32p]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib.L


def n_threads_default():
    return os.cpu_count() or 1


def _sp(spacing):
    return np.ascontiguousarray(spacing if spacing is not None else (1.0, 1.0, 1.0), np.float64)


def _dims(vol):
    nz, ny, nx = vol.shape
    return nx, ny, nz


# ---------------------------------------------------------------- ITK primitives
def gaussian_coefficients(sigma, spacing=1.0):
    c = np.zeros(20, np.float64)
    lib().orc_gaussian_coefficients(sigma, spacing, c)
    return c


def gaussian_line(c20, data, arith=ARITH_PLAIN):
    data = np.ascontiguousarray(data, np.float64)
    out = np.zeros_like(data)
    lib().orc_gaussian_line(np.ascontiguousarray(c20, np.float64), data, out, data.size, arith)
    return out


def smoothing_recursive_gaussian(vol, sigma, spacing=None, arith=ARITH_PLAIN, threads=None):
    vol = np.ascontiguousarray(vol, np.float32)
    out = np.empty_like(vol)
    rc = lib().orc_smoothing_recursive_gaussian(vol, out, *_dims(vol), _sp(spacing), sigma, arith,
                                                threads or n_threads_default())
    if rc:
        raise ValueError("recursive Gaussian needs at least 4 samples per axis")
    return out


def derivative(vol, axis, order, spacing=None, threads=None):
    vol = np.ascontiguousarray(vol, np.float32)
    out = np.empty_like(vol)
    lib().orc_derivative(vol, out, *_dims(vol), axis, order, _sp(spacing), threads or n_threads_default())
    return out


def gradient_magnitude(vol, spacing=None, threads=None):
    vol = np.ascontiguousarray(vol, np.float32)
    out = np.empty_like(vol)
    lib().orc_gradient_magnitude(vol, out, *_dims(vol), _sp(spacing), threads or n_threads_default())
    return out


# ---------------------------------------------------------------- reference numerics restated
def eig_f32(A6, math_mode=0):
    A6 = np.ascontiguousarray(A6, np.float32).reshape(-1, 6)
    out = np.empty((A6.shape[0], 3), np.float32)
    lib().orc_eig_f32(A6, out, A6.shape[0], math_mode)
    return out


def eig_f64(A6):
    A6 = np.ascontiguousarray(A6, np.float64).reshape(-1, 6)
    out = np.empty((A6.shape[0], 3), np.float64)
    lib().orc_eig_f64(A6, out, A6.shape[0])
    return out


def features_f32(A6, math_mode=0):
    A6 = np.ascontiguousarray(A6, np.float32).reshape(-1, 6)
    out = np.empty((A6.shape[0], 6), np.float32)
    lib().orc_features_f32(A6, out, A6.shape[0], math_mode)
    return out


def features_f64(A6):
    A6 = np.ascontiguousarray(A6, np.float64).reshape(-1, 6)
    out = np.empty((A6.shape[0], 6), np.float64)
    lib().orc_features_f64(A6, out, A6.shape[0])
    return out


def hist_f32(edges, values):
    edges = np.ascontiguousarray(edges, np.float32)
    values = np.ascontiguousarray(values, np.float32).ravel()
    counts = np.zeros(edges.size + 1, np.uint32)
    freqs = np.zeros(edges.size + 1, np.float32)
    lib().orc_hist_f32(edges, edges.size, values, values.size, counts, freqs)
    return counts, freqs


def determine_edges(sorted_samples, n_bins, dtype=np.float64):
    s = np.ascontiguousarray(sorted_samples, dtype)
    edges = np.zeros(max(n_bins - 1, 1), dtype)
    fn = lib().orc_determine_edges_f64 if dtype == np.float64 else lib().orc_determine_edges_f32
    rc = fn(s, s.size, edges, n_bins)
    if rc == 1:
        raise IndexError("Too many bins. Number of bins must be less or equal to number of samples")
    return edges[: n_bins - 1]


# ---------------------------------------------------------------- compositions
def hessian6(vol, spacing=None, fdhf_tool_bug=False, threads=None):
    """-> (nz, ny, nx, 6) interleaved [Dxx,Dxy,Dxz,Dyy,Dyz,Dzz]"""
    vol = np.ascontiguousarray(vol, np.float32)
    out = np.empty(vol.shape + (6,), np.float32)
    lib().orc_hessian6(vol, out.reshape(-1), *_dims(vol), _sp(spacing), int(fdhf_tool_bug),
                       threads or n_threads_default())
    return out


def functor_volume(hess6, mask=None, threads=None):
    hess6 = np.ascontiguousarray(hess6, np.float32)
    out = np.empty_like(hess6)
    n = hess6.size // 6
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    lib().orc_functor_volume_f32(hess6.reshape(-1), _opt(m, C.c_void_p), out.reshape(-1), n,
                                 threads or n_threads_default())
    return out


def normalized_gaussian(img, certainty, sigma, spacing=None, arith=ARITH_PLAIN, threads=None):
    img = np.ascontiguousarray(img, np.float32)
    certainty = np.ascontiguousarray(certainty, np.float32)
    out = np.empty_like(img)
    rc = lib().orc_normalized_gaussian(img, certainty, out, *_dims(img), _sp(spacing), sigma, arith,
                                       threads or n_threads_default())
    if rc:
        raise ValueError("recursive Gaussian needs at least 4 samples per axis")
    return out


def emphysema_features(img, mask, sigma, spacing=None, arith=ARITH_PLAIN, threads=None):
    """-> (8, nz, ny, nx) SoA planes in FEATURE_NAMES8 order"""
    img = np.ascontiguousarray(img, np.float32)
    mask = np.ascontiguousarray(mask, np.uint8)
    out = np.empty((8,) + img.shape, np.float32)
    rc = lib().orc_emphysema_features(img, mask, out.reshape(-1), *_dims(img), _sp(spacing), sigma,
                                      arith, threads or n_threads_default())
    if rc:
        raise ValueError("recursive Gaussian needs at least 4 samples per axis")
    return out


def fd_hessian_features(img, mask=None, sigma=0.0, spacing=None, arith=ARITH_PLAIN,
                        fdhf_tool_bug=False, threads=None):
    """-> (6, nz, ny, nx) SoA planes [eig1, eig2, eig3, LoG, Curvature, Frobenius]"""
    img = np.ascontiguousarray(img, np.float32)
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    out = np.empty((6,) + img.shape, np.float32)
    rc = lib().orc_fd_hessian_features(img, _opt(m, C.c_void_p), out.reshape(-1), *_dims(img),
                                       _sp(spacing), sigma, arith, int(fdhf_tool_bug),
                                       threads or n_threads_default())
    if rc:
        raise ValueError("recursive Gaussian needs at least 4 samples per axis")
    return out


def fd_gradient_features(img, mask=None, spacing=None, threads=None):
    img = np.ascontiguousarray(img, np.float32)
    m = None if mask is None else np.ascontiguousarray(mask, np.float32)
    out = np.empty_like(img)
    lib().orc_fd_gradient_features(img, _opt(m, C.c_void_p), out, *_dims(img), _sp(spacing),
                                   threads or n_threads_default())
    return out


def features_histograms(feats, mask, edges, rois=None):
    """feats (F, nz, ny, nx); edges (F, E); rois (R, 6) int32 {x0,y0,z0,sx,sy,sz} or None
    -> counts (R or 1, F, E+1) uint32"""
    feats = np.ascontiguousarray(feats, np.float32)
    edges = np.ascontiguousarray(edges, np.float32)
    F = feats.shape[0]
    nz, ny, nx = feats.shape[1:]
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    r = None if rois is None else np.ascontiguousarray(rois, np.int32).reshape(-1, 6)
    R = 1 if r is None else r.shape[0]
    counts = np.zeros((R, F, edges.shape[1] + 1), np.uint32)
    lib().orc_features_histograms(feats.reshape(-1), F, _opt(m, C.c_void_p), nx, ny, nz,
                                  _opt(r, C.c_void_p), 0 if r is None else R, edges.reshape(-1),
                                  edges.shape[1], counts.reshape(-1))
    return counts


# ---------------------------------------------------------------- the reference's own headers
class Ref:
    """oracle/_ref/libife_ref.so: the reference's ITK-light headers compiled unmodified."""

    def __init__(self):
        so = os.path.join(_HERE, "_ref", "libife_ref.so")
        if not os.path.exists(so):
            raise FileNotFoundError(so)
        L = self.L = C.CDLL(so)
        L.ref_math_overload_is_double.restype = C.c_int
        L.ref_eig_f32.argtypes = [_f32p, _f32p, C.c_size_t]
        L.ref_eig_f64.argtypes = [_f64p, _f64p, C.c_size_t]
        L.ref_features_f32.argtypes = [_f32p, _f32p, C.c_size_t]
        L.ref_features_f64.argtypes = [_f64p, _f64p, C.c_size_t]
        L.ref_hist_f32.argtypes = [_f32p, C.c_int, _f32p, C.c_size_t, _u32p, _f32p]
        L.ref_determine_edges_f64.argtypes = [_f64p, C.c_size_t, _f64p, C.c_size_t]
        L.ref_determine_edges_f32.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t]
        L.ref_functor_volume_f32.argtypes = [_f32p, C.c_void_p, _f32p, C.c_size_t, C.c_int]

    def math_overload_is_double(self):
        return bool(self.L.ref_math_overload_is_double())

    def eig_f32(self, A6):
        A6 = np.ascontiguousarray(A6, np.float32).reshape(-1, 6)
        out = np.empty((A6.shape[0], 3), np.float32)
        self.L.ref_eig_f32(A6, out, A6.shape[0])
        return out

    def eig_f64(self, A6):
        A6 = np.ascontiguousarray(A6, np.float64).reshape(-1, 6)
        out = np.empty((A6.shape[0], 3), np.float64)
        self.L.ref_eig_f64(A6, out, A6.shape[0])
        return out

    def features_f32(self, A6):
        A6 = np.ascontiguousarray(A6, np.float32).reshape(-1, 6)
        out = np.empty((A6.shape[0], 6), np.float32)
        self.L.ref_features_f32(A6, out, A6.shape[0])
        return out

    def hist_f32(self, edges, values):
        edges = np.ascontiguousarray(edges, np.float32)
        values = np.ascontiguousarray(values, np.float32).ravel()
        counts = np.zeros(edges.size + 1, np.uint32)
        freqs = np.zeros(edges.size + 1, np.float32)
        self.L.ref_hist_f32(edges, edges.size, values, values.size, counts, freqs)
        return counts, freqs

    def determine_edges(self, sorted_samples, n_bins, dtype=np.float64):
        s = np.ascontiguousarray(sorted_samples, dtype)
        edges = np.zeros(max(n_bins - 1, 1), dtype)
        fn = self.L.ref_determine_edges_f64 if dtype == np.float64 else self.L.ref_determine_edges_f32
        rc = fn(s, s.size, edges, n_bins)
        if rc == 1:
            raise IndexError("out_of_range")
        return edges[: n_bins - 1]

    def functor_volume(self, hess6, mask=None, threads=None):
        hess6 = np.ascontiguousarray(hess6, np.float32)
        out = np.empty_like(hess6)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self.L.ref_functor_volume_f32(hess6.reshape(-1), _opt(m, C.c_void_p), out.reshape(-1),
                                      hess6.size // 6, threads or n_threads_default())
        return out


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libife_ref.so"))


def emphysema_features_reference_arm(img, mask, sigma, spacing=None, arith=ARITH_PLAIN, threads=None):
    """The CPU baseline composition: restated ITK stages wired as
    ImageToEmphysemaFeaturesFilter.hxx:11-55,94-121, with the per-voxel functor taken from
    the reference's own header (oracle/_ref, incl. its per-voxel heap allocations) when
    that library is present, else from the restatement.  -> (8, nz, ny, nx)"""
    img = np.ascontiguousarray(img, np.float32)
    mask = np.ascontiguousarray(mask, np.uint8)
    threads = threads or n_threads_default()
    blur = normalized_gaussian(img, mask.astype(np.float32), sigma, spacing, arith, threads)
    gm = gradient_magnitude(blur, spacing, threads)
    hess = hessian6(blur, spacing, False, threads)
    feat = (Ref().functor_volume(hess, None, threads) if ref_available()
            else functor_volume(hess, None, threads))
    out = np.empty((8,) + img.shape, np.float32)
    inside = mask != 0
    out[0] = np.where(inside, blur, np.float32(0))
    out[1] = np.where(inside, gm, np.float32(0))
    for k in range(6):
        out[2 + k] = np.where(inside, feat[..., k], np.float32(0))
    return out
