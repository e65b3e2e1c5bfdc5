"""CPU tests of the drop-in boundary: the C-ABI shared library loads, exports every symbol
include/ife_cuda.h declares, and refuses to work (loudly) without a CUDA device."""
import ctypes

import pytest

import ife_b200


def test_library_loads_and_exports_every_declared_symbol():
    L = ife_b200.load_library()
    names = ife_b200.declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.ife_cuda_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = ife_b200.load_library()
    h = ctypes.c_void_p()
    assert L.ife_cuda_create(0, ctypes.byref(h)) == -3   # IFE_E_CUDA
    assert not h.value
    with pytest.raises(ife_b200.IfeError):
        ife_b200.Context(0)


def test_slab_partition_host_arithmetic():
    L = ife_b200.load_library()
    for nz, p in ((400, 8), (1024, 8), (17, 4), (5, 5)):
        covered = 0
        for r in range(p):
            z0, z1 = ctypes.c_int(), ctypes.c_int()
            L.ife_cuda_slab_range(nz, p, r, ctypes.byref(z0), ctypes.byref(z1))
            assert (z0.value, z1.value) == ife_b200.slab_range(nz, p, r)
            assert z0.value == covered
            covered = z1.value
        assert covered == nz
    assert L.ife_cuda_slab_halo(4.8, 1.0, 12.0) == ife_b200.slab_halo(4.8) == 63
    assert L.ife_cuda_slab_halo(4.8, 1.0, 0.0) == 63      # default factor
    assert L.ife_cuda_slab_halo(1.0, 2.0, 8.0) == 9
