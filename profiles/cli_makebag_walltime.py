"""Wall time of the MakeBag command line at BASELINE's full size: 512x512x400 int16 scan + uint16 lung
mask as .nii.gz (what a user has on disk), four scales, 40 edges per (scale, feature), 50 random
41^3 ROIs.  Usage: python profiles/cli_makebag_walltime.py"""
import os, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, nifti_util
d = tempfile.mkdtemp(prefix="ife_bag_")
dev = torch.device("cuda", 0)
img, mask = bench.synth_scan_torch(torch, dev, 100, "lung")
nifti_util.write(d + "/img.nii.gz", img.cpu().numpy().astype(np.int16))
nifti_util.write(d + "/mask.nii.gz", mask.cpu().numpy().astype(np.uint16))
del img, mask
torch.cuda.empty_cache()
with open(d + "/hist.txt", "w") as f:
    for row in range(32):
        f.write(",".join(repr(float(v)) for v in np.linspace(-900.0 + row, 600.0 + row, 40)) + "\n")
exe = os.path.join(ROOT, "image-feature-extraction_b200", "bin", "MakeBag")
t0 = time.time()
p = subprocess.run([exe, "-i", d + "/img.nii.gz", "-m", d + "/mask.nii.gz", "-H", d + "/hist.txt", "-o", d, "-p", "case",
                    "-s", "0.6", "-s", "1.2", "-s", "2.4", "-s", "4.8", "-n", "50", "-x", "41", "-y", "41", "-z", "41", "-S", "7"],
                   capture_output=True, text=True, env=dict(os.environ, IFE_TIMING="1", IFE_ALLOC_TRACE="1"))
dt = time.time() - t0
rows = open(d + "/case.bag").read().strip().splitlines() if os.path.exists(d + "/case.bag") else []
print("MakeBag 512x512x400, 4 scales, 50 ROIs: rc=%d, %d bag rows, total wall %.2f s" % (p.returncode, len(rows), dt))
print("\n".join(l for l in p.stderr.splitlines() if l.startswith("[")))
if p.returncode: print(p.stderr[-800:])
shutil.rmtree(d)
