"""Wall time of the shipped command-line path at BASELINE's full size: ExtractFeatures on a
512x512x400 int16 scan + uint8 mask (uncompressed .nii inputs in /tmp), four scales.  Prints the
tool's total wall time and the time of its one ife_cuda_emphysema_features call (page-locked
buffers through the C++ facade).  The .nii.gz outputs (the reference's OUT_FILE_TYPE) are
written with zlib and dominate the total; they go to /tmp and are deleted.
Usage: python profiles/cli_walltime.py [nz]"""
import os, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, nifti_util
nz = int(sys.argv[1]) if len(sys.argv) > 1 else 400
d = tempfile.mkdtemp(prefix="ife_cli_")
dev = torch.device("cuda", 0)
img, mask = bench.synth_scan_torch(torch, dev, 100, "lung")
nifti_util.write(d + "/img.nii", img[:nz].cpu().numpy().astype(np.int16))
nifti_util.write(d + "/mask.nii", mask[:nz].cpu().numpy())
del img, mask
torch.cuda.empty_cache()
exe = os.path.join(ROOT, "image-feature-extraction_b200", "bin", "ExtractFeatures")
t0 = time.time()
p = subprocess.run([exe, "-i", d + "/img.nii", "-m", d + "/mask.nii", "-o", d + "/f", "-s", "0.6", "-s", "1.2", "-s", "2.4", "-s", "4.8"],
                   capture_output=True, text=True, env=dict(os.environ, IFE_TIMING="1"))
dt = time.time() - t0
n_out = len([f for f in os.listdir(d) if f.startswith("f_scale")])
print("ExtractFeatures 512x512x%d, 4 scales: rc=%d, %d files, total wall %.1f s; %s" % (nz, p.returncode, n_out, dt, p.stderr.strip().splitlines()[-1] if p.stderr.strip() else ""))
if os.environ.get("IFE_ALLOC_TRACE"):
    print("\n".join(l for l in p.stderr.splitlines() if l.startswith("[ife]")))
shutil.rmtree(d)
