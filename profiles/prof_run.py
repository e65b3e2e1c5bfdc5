"""Short single-GPU run for ncu: the flagship path at full size (512x512x400), one scale per
call, device-resident inputs.  Usage: python profiles/prof_run.py [sigma ...] [--hist [--eq]] [--lung] [--plain]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-feature-extraction_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import ife_b200

sigmas = [float(a) for a in sys.argv[1:] if not a.startswith("--")] or [1.2]
arith = ife_b200.ARITH_PLAIN if "--plain" in sys.argv else ife_b200.ARITH_FMA
dev = torch.device("cuda", 0)
ctx = ife_b200.Context(0, arith=arith)
img, mask = bench.synth_scan_torch(torch, dev, 100, "lung" if "--lung" in sys.argv else "ones")
nx, ny, nz = bench.DIMS
out = torch.empty((len(sigmas), 8, nz, ny, nx), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
for it in range(2):
    ctx.emphysema_features_dev(img.data_ptr(), mask.data_ptr(), out.data_ptr(), bench.DIMS, sigmas)
    ctx.synchronize()
if "--hist" in sys.argv:
    if "--eq" in sys.argv:    # equal-frequency edges of this scan, as MakeBag gets them
        i0 = bench.SIGMAS.index(sigmas[0]) * 8
        edges = bench.equalized_edges_from_scan(torch, ctx, img, mask, 40)[i0:i0 + 8 * len(sigmas)].copy()
    else:
        edges = np.tile(np.linspace(-1.0, 1.0, 40, dtype=np.float32), (len(sigmas) * 8, 1))
    counts = torch.zeros((1, len(sigmas) * 8, 41), dtype=torch.int32, device=dev)
    ctx.emphysema_histograms_dev(img.data_ptr(), mask.data_ptr(), counts.data_ptr(), bench.DIMS, sigmas, edges)
    ctx.synchronize()
print("prof_run ok, launches:", ctx.launch_count())
ctx.close()
