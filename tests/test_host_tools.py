"""The C++ host side: CLI surface of the four tools (flags, exit codes, output names) and,
on a GPU, their outputs against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import nifti_util
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "image-feature-extraction_b200", "bin")


def run(tool, *args):
    return subprocess.run([os.path.join(BIN, tool)] + list(args), capture_output=True, text=True, timeout=600)


def test_host_selftest(tmp_path):
    p = run("ife_host_selftest", str(tmp_path))
    assert p.returncode == 0, p.stdout + p.stderr


@pytest.mark.parametrize("tool,required", [("ExtractFeatures", "-i -m -o -s"),
                                           ("MaskedNormalizedConvolution", "-i -c -s -o"),
                                           ("FiniteDifference_HessianFeatures", "-i -m -o"),
                                           ("FiniteDifference_GradientFeatures", "-i -m -o")])
def test_cli_surface(tool, required):
    p = run(tool, "--help")
    assert p.returncode == 0
    for flag in required.split():
        assert flag + " <" in p.stdout
    assert run(tool, "--version").stdout.strip().endswith("version: 0.1")
    p = run(tool)                      # required arguments missing -> TCLAP-style failure
    assert p.returncode == 1 and "PARSE ERROR" in p.stderr
    p = run(tool, "--nonsense", "1")
    assert p.returncode == 1 and "PARSE ERROR" in p.stderr


def test_cli_reports_unreadable_input(tmp_path):
    p = run("ExtractFeatures", "-i", str(tmp_path / "missing.nii.gz"), "-m", "x", "-o", str(tmp_path / "o"), "-s", "1")
    assert p.returncode == 1 and "Failed to process." in p.stderr and "Image:" in p.stderr


def _eig_close(got, ref):
    same = (got == ref) | (np.isnan(got) & np.isnan(ref))
    return same.mean() > 1 - 1e-4


@pytest.mark.gpu
def test_tools_end_to_end_against_oracle(tmp_path, oracle):
    shape = (20, 24, 28)
    sp = tuple(float(np.float32(v)) for v in (0.8, 0.8, 1.6))   # pixdim is float32 on disk
    img = synth.ct_like(shape, seed=51, n_blobs=6)
    lab = synth.lung_mask(shape)                     # labels {0,1,2}
    d = str(tmp_path)
    nifti_util.write(d + "/img.nii.gz", img.astype(np.int16), sp)      # CT is int16 on disk
    nifti_util.write(d + "/mask.nii.gz", lab, sp)
    imgf = img.astype(np.int16).astype(np.float32)
    m01 = synth.clamp01(lab)

    p = run("ExtractFeatures", "-i", d + "/img.nii.gz", "-m", d + "/mask.nii.gz", "-o", d + "/feat", "-s", "0.6", "--scale", "1.2")
    assert p.returncode == 0, p.stderr
    names = ["GaussianBlur", "GradientMagnitude", "Eigenvalue1", "Eigenvalue2", "Eigenvalue3",
             "LaplacianOfGaussian", "GaussianCurvature", "FrobeniusNorm"]
    for sigma, tag in ((np.float32(0.6), "0.600000"), (np.float32(1.2), "1.200000")):
        ref = oracle.emphysema_features(imgf, m01, float(sigma), spacing=sp, arith=1)
        for k, nm in enumerate(names):
            got, pix = nifti_util.read("%s/feat_scale_%s%s.nii.gz" % (d, tag, nm))
            assert got.dtype == np.float32 and np.allclose(pix, sp)
            assert (np.array_equal(got, ref[k]) if k < 2 else _eig_close(got, ref[k])), nm

    p = run("MaskedNormalizedConvolution", "-i", d + "/img.nii.gz", "-c", d + "/mask.nii.gz", "-s", "1.5", "-o", d, "-m", "true")
    assert p.returncode == 0 and "Processing scale 1.5" in p.stdout, p.stderr
    got, _ = nifti_util.read(d + "/normconv_scale_1.500000.nii.gz")
    cert = lab.astype(np.float32)
    ref = oracle.normalized_gaussian(imgf, cert, 1.5, spacing=sp, arith=1)
    assert np.array_equal(got, np.where(cert != 0, ref, np.float32(0)))

    p = run("FiniteDifference_HessianFeatures", "-i", d + "/img.nii.gz", "-m", d + "/mask.nii.gz", "-o", d, "-p", "h_")
    assert p.returncode == 0, p.stderr
    ref = oracle.fd_hessian_features(imgf, lab, spacing=sp)
    for k, nm in enumerate(["eig1", "eig2", "eig3", "LoG", "Curvature", "Frobenius"]):
        got, _ = nifti_util.read("%s/h_%s.nii.gz" % (d, nm))
        assert _eig_close(got, ref[k]), nm

    p = run("FiniteDifference_GradientFeatures", "-i", d + "/img.nii.gz", "-m", d + "/mask.nii.gz", "-o", d)
    assert p.returncode == 0, p.stderr
    got, _ = nifti_util.read(d + "/gradient_GradientMagnitude.nii.gz")
    assert np.array_equal(got, oracle.fd_gradient_features(imgf, lab.astype(np.float32), spacing=sp))
