"""Timing of the many-ROI histogram path (MakeBagDense semantics) at full size: N ROIs of 41^3
centred on in-mask voxels of the 512x512x400 lung-mask scan, all four scales, device-resident.
Usage: python profiles/dense_rois_run.py [N]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-feature-extraction_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import ife_b200

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda", 0)
ctx = ife_b200.Context(0)
img, mask = bench.synth_scan_torch(torch, dev, 100, "lung")
nx, ny, nz = bench.DIMS
edges = bench.equalized_edges_from_scan(torch, ctx, img, mask, 40)
m = mask.clone()
m[:20] = 0; m[-21:] = 0; m[:, :20] = 0; m[:, -21:] = 0; m[:, :, :20] = 0; m[:, :, -21:] = 0
idx = torch.nonzero(m.reshape(-1)).reshape(-1)
idx = idx[:: max(1, idx.numel() // N)][:N].cpu().numpy()
z, y, x = idx // (nx * ny), (idx // nx) % ny, idx % nx
rois = np.stack([x - 20, y - 20, z - 20, np.full_like(x, 41), np.full_like(x, 41), np.full_like(x, 41)], 1).astype(np.int32)
counts = torch.zeros((len(rois), 32, 41), dtype=torch.int32, device=dev)
for it in range(2):
    t0 = time.perf_counter()
    ctx.emphysema_histograms_dev(img.data_ptr(), mask.data_ptr(), counts.data_ptr(), bench.DIMS, bench.SIGMAS, edges, rois)
    ctx.synchronize()
    dt = time.perf_counter() - t0
inserted = int(counts[:, 0].sum().item())
print("many-ROI path: %d ROIs of 41^3, 4 scales: %.1f ms per call (%.2f us per ROI-scale, %.1f G voxel-inserts/s per feature row)"
      % (len(rois), dt * 1e3, dt * 1e6 / len(rois) / 4, inserted * 4 / dt / 1e9))
ctx.close()
