"""Time of ife_cuda_host_alloc (page-locked host memory) by size: first call (CUDA context creation
included), then 0.42 GB (one volume) and 13.4 GB (the 32 feature volumes of an ExtractFeatures run).
IFE_NO_MAPPED_HOST_ALLOC=1 in the environment takes the cudaHostAlloc path for comparison."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = C.CDLL(os.path.join(ROOT, "image-feature-extraction_b200", "lib", "libife_cuda.so"))
L.ife_cuda_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
L.ife_cuda_host_free.argtypes = [C.c_void_p]
def alloc(n):
    p = C.c_void_p()
    t0 = time.time()
    rc = L.ife_cuda_host_alloc(n, C.byref(p))
    t1 = time.time()
    L.ife_cuda_host_free(p)
    return rc, t1 - t0, time.time() - t1
print("mode:", "cudaHostAlloc" if os.environ.get("IFE_NO_MAPPED_HOST_ALLOC") else "mmap + huge pages + parallel touch + cudaHostRegister")
for label, n in (("first call, 64 MB", 64 << 20), ("0.42 GB", 419430400), ("13.4 GB", 13421772800), ("13.4 GB again", 13421772800)):
    rc, ta, tf = alloc(n)
    print("%-18s rc=%d alloc %.3f s free %.3f s" % (label, rc, ta, tf), flush=True)
