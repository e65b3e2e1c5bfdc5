"""Timing experiments: per-kernel CUDA-event times of one scale at full size for whichever
library IFE_CUDA_LIB points to (variants built with -DIFE_EXP=n drop one piece of the fused
kernel's arithmetic each; they are NOT products and never pass parity).
Usage: IFE_CUDA_LIB=... python profiles/exp_features.py [sigma] [--lung]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-feature-extraction_b200"))
sys.path.insert(0, ROOT)
import torch
import bench
import ife_b200

sig = [float(a) for a in sys.argv[1:] if not a.startswith("--")] or [1.2]
dev = torch.device("cuda", 0)
ctx = ife_b200.Context(0)
img, mask = bench.synth_scan_torch(torch, dev, 100, "lung" if "--lung" in sys.argv else "ones")
nx, ny, nz = bench.DIMS
out = torch.empty((len(sig), 8, nz, ny, nx), dtype=torch.float32, device=dev)
for it in range(3):
    ctx.emphysema_features_dev(img.data_ptr(), mask.data_ptr(), out.data_ptr(), bench.DIMS, sig)
ctx.synchronize()
ctx.profile_enable(True)
ctx.profile_read()
for it in range(5):
    ctx.emphysema_features_dev(img.data_ptr(), mask.data_ptr(), out.data_ptr(), bench.DIMS, sig)
r = ctx.profile_read()
print(os.path.basename(os.environ.get("IFE_CUDA_LIB", "product")),
      {k: round(v[0] / max(v[1], 1), 3) for k, v in r.items() if v[1]})
ctx.close()
