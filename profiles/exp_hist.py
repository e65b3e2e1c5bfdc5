"""Timing experiments: per-kernel CUDA-event times of the histogram workload (whole-mask
DenseHistograms, 40 equalized edges) and of the extract workload at full size, one scale, for
whichever library IFE_CUDA_LIB points to.
Usage: IFE_CUDA_LIB=... python profiles/exp_hist.py [--lung]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-feature-extraction_b200"))
sys.path.insert(0, ROOT)
import torch
import bench
import ife_b200

sig = [1.2]
dev = torch.device("cuda", 0)
ctx = ife_b200.Context(0)
try:
    ctx.set_option("support_box", 0)      # kernels only: identical work for every library
except Exception:
    pass
res = {}
for mask_kind in ("ones", "lung"):
    img, mask = bench.synth_scan_torch(torch, dev, 100, mask_kind)
    edges = bench.equalized_edges_from_scan(torch, ctx, img, mask, 40)[8:16].copy()
    counts = torch.zeros((1, 8, 41), dtype=torch.int32, device=dev)
    nx, ny, nz = bench.DIMS
    out = torch.empty((1, 8, nz, ny, nx), dtype=torch.float32, device=dev)
    for name, fn in (("hist", lambda: ctx.emphysema_histograms_dev(img.data_ptr(), mask.data_ptr(), counts.data_ptr(), bench.DIMS, sig, edges, None)),
                     ("extract", lambda: ctx.emphysema_features_dev(img.data_ptr(), mask.data_ptr(), out.data_ptr(), bench.DIMS, sig))):
        for it in range(3):
            fn()
        ctx.synchronize()
        ctx.profile_enable(True)
        ctx.profile_read()
        for it in range(8):
            fn()
        r = ctx.profile_read()
        ctx.profile_enable(False)
        res[mask_kind + "/" + name] = {k[-1] + k[0]: round(v[0] / max(v[1], 1), 3) for k, v in r.items() if v[1] and k != "other"}
    del out
print(os.path.basename(os.environ.get("IFE_CUDA_LIB", "product")), res)
ctx.close()
