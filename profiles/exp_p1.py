"""Per-kernel CUDA-event times of the un-normalized pipelines at 512x512x400:
P1 = plain Gaussian (one field) + Hessian-eigen features (6 volumes out), P0 = the same without
smoothing (FiniteDifference_HessianFeatures as shipped).  Usage: python profiles/exp_p1.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-feature-extraction_b200"))
sys.path.insert(0, ROOT)
import torch
import bench
import ife_b200

dev = torch.device("cuda", 0)
ctx = ife_b200.Context(0)
img, mask = bench.synth_scan_torch(torch, dev, 100, "ones")
nx, ny, nz = bench.DIMS
out = torch.empty((6, nz, ny, nx), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
for name, sigma in (("P1 sigma=1.2", 1.2), ("P0 no smoothing", 0.0)):
    for it in range(3):
        ctx.hessian_eigen_features_dev(img.data_ptr(), mask.data_ptr(), out.data_ptr(), bench.DIMS, sigma)
    ctx.synchronize()
    ctx.profile_enable(True)
    ctx.profile_read()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    for it in range(8):
        ctx.hessian_eigen_features_dev(img.data_ptr(), mask.data_ptr(), out.data_ptr(), bench.DIMS, sigma)
    r = ctx.profile_read()
    ctx.profile_enable(False)
    per = {k: round(v[0] / max(v[1], 1), 3) for k, v in r.items() if v[1]}
    tot = sum(v[0] for v in r.values()) / 8
    print(name, per, "sum %.3f ms -> %.1f Gvoxel/s" % (tot, nx * ny * nz / tot / 1e6))
ctx.close()
