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


def run(tool, *args, env=None):
    return subprocess.run([os.path.join(BIN, tool)] + list(args), capture_output=True, text=True, timeout=600,
                          env=None if env is None else dict(os.environ, **env))


def test_host_selftest(tmp_path):
    p = run("ife_host_selftest", str(tmp_path))
    assert p.returncode == 0, p.stdout + p.stderr


def test_multithreaded_gzip_writer_makes_an_ordinary_gz_file(tmp_path):
    """The .nii.gz writer deflates 8 MB pieces on several threads and joins them into ONE gzip member
    (ife/IO/NiftiIO.h, gz_write_parallel): Python's gzip and `gzip -t` must read it like any other."""
    import gzip
    import zlib
    p = run("ife_host_selftest", str(tmp_path), "keep")       # leaves the 21 MB, three-piece volume behind
    assert p.returncode == 0, p.stdout + p.stderr
    path = os.path.join(str(tmp_path), "ife_selftest_big.nii.gz")
    blob = open(path, "rb").read()
    assert blob[:4] == b"\x1f\x8b\x08\x00"
    raw = gzip.decompress(blob)
    assert len(raw) == 352 + 256 * 160 * 131 * 4
    assert int.from_bytes(blob[-8:-4], "little") == zlib.crc32(raw)           # the combined CRC-32 of the pieces
    assert int.from_bytes(blob[-4:], "little") == len(raw) % (1 << 32)
    vol, _ = nifti_util.read(path)
    assert vol.shape == (131, 160, 256) and vol.dtype == np.float32
    flat = vol.ravel()
    assert np.all(flat[:1024] == 0) and np.all(flat[4096:4096 + 1024] == 0) and flat[2000] != 0


@pytest.mark.parametrize("tool,required", [("ExtractFeatures", "-i -m -o -s"),
                                           ("MaskedNormalizedConvolution", "-i -c -s -o"),
                                           ("FiniteDifference_HessianFeatures", "-i -m -o"),
                                           ("FiniteDifference_GradientFeatures", "-i -m -o"),
                                           ("MakeBag", "-i -m -H -o -s"),
                                           ("MakeBagDense", "-i -m -H -o -s"),
                                           ("MakeBagOnlyIntensity", "-i -m -H -o"),
                                           ("DetermineHistogramBinEdges_MultiScaleEigenvalueFeatures", "-i -o -b -S -s -f")])
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


REFERENCE_TOOL = "/root/reference/tools/ExtractFeatures.cxx"


@pytest.mark.skipif(not os.path.exists(REFERENCE_TOOL), reason="the reference checkout is only mounted in the authoring container")
def test_reference_tool_body_compiles_and_links_against_the_facades(tmp_path):
    """The drop-in claim at the C++ level: the BODY of the reference's own ExtractFeatures tool
    (tools/ExtractFeatures.cxx from `typedef float PixelType;` to the end of main: readers, clamp,
    itk::ImageToEmphysemaFeaturesFilter<ImageType, MaskType, VectorImageType>, index selection,
    writer, the SetSigma / UpdateLargestPossibleRegion / writer->Update loop, the
    itk::ExceptionObject handler) is read from the reference checkout at test time -- no
    reference source lives in this repository -- and compiled UNMODIFIED against
    host/include (facades + itk_compat), then linked with libife_cuda.so.  Only the TCLAP
    command-line block in front of it is replaced."""
    lines = open(REFERENCE_TOOL).read().splitlines()
    start = next(i for i, l in enumerate(lines) if "typedef float PixelType;" in l)
    body = "\n".join(lines[start:])
    src = tmp_path / "ref_body.cxx"
    src.write_text("""#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>
#include "itkImageFileReader.h"
#include "itkImageFileWriter.h"
#include "itkVectorIndexSelectionCastImageFilter.h"
#include "itkClampImageFilter.h"
#include "ife/Filters/ImageToEmphysemaFeaturesFilter.h"
#include "ife/Util/Path.h"
const std::string VERSION("0.1");
const std::string OUT_FILE_TYPE(".nii.gz");
int main( int argc, char* argv[] ) {
  if (argc < 5) return EXIT_FAILURE;
  const std::string imagePath( argv[1] );
  const std::string maskPath( argv[2] );
  const std::string outBasePath( argv[3] );
  std::vector< float > scales;
  for (int i = 4; i < argc; ++i) scales.push_back((float)std::atof(argv[i]));
""" + body + "\n")
    pkg = os.path.join(ROOT, "image-feature-extraction_b200")
    exe = tmp_path / "ref_body"
    cmd = ["g++", "-std=c++14", "-Wall", "-I", os.path.join(pkg, "host", "include"),
           "-I", os.path.join(pkg, "host", "include", "itk_compat"), "-I", os.path.join(ROOT, "include"),
           str(src), "-o", str(exe), "-L", os.path.join(pkg, "lib"), "-life_cuda", "-lz",
           "-Wl,-rpath," + os.path.join(pkg, "lib")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    # without a GPU the program must fail the way the reference tool fails: EXIT_FAILURE and the
    # "Failed to process." report (here: unreadable input comes first)
    r = subprocess.run([str(exe), str(tmp_path / "missing.nii"), str(tmp_path / "missing2.nii"), str(tmp_path / "o"), "1.0"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Failed to process." in r.stderr


def test_cli_reports_unreadable_input(tmp_path):
    p = run("ExtractFeatures", "-i", str(tmp_path / "missing.nii.gz"), "-m", "x", "-o", str(tmp_path / "o"), "-s", "1")
    assert p.returncode == 1 and "Failed to process." in p.stderr and "Image:" in p.stderr


@pytest.mark.gpu
def test_tools_honour_ife_cuda_options(tmp_path):
    """IFE_CUDA_OPTIONS selects kernels (never results) for the command-line tools too; a bad entry is an error."""
    shape = (20, 24, 32)
    d = str(tmp_path)
    nifti_util.write(d + "/img.nii.gz", synth.ct_like(shape, seed=52, n_blobs=6))
    nifti_util.write(d + "/mask.nii.gz", synth.lung_mask(shape))
    outs = []
    for tag, env in (("a", None), ("b", {"IFE_CUDA_OPTIONS": "tma_passes=0,march4=0"})):
        p = run("ExtractFeatures", "-i", d + "/img.nii.gz", "-m", d + "/mask.nii.gz", "-o", d + "/" + tag, "-s", "1.2", env=env)
        assert p.returncode == 0, p.stderr
        outs.append([nifti_util.read("%s/%s_scale_1.200000%s.nii.gz" % (d, tag, nm))[0] for nm in ("GaussianBlur", "Eigenvalue1")])
    for x, y in zip(*outs):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    p = run("ExtractFeatures", "-i", d + "/img.nii.gz", "-m", d + "/mask.nii.gz", "-o", d + "/c", "-s", "1.2",
            env={"IFE_CUDA_OPTIONS": "no_such_option=1"})
    assert p.returncode == 1 and "IFE_CUDA_OPTIONS" in p.stderr


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


@pytest.mark.gpu
def test_makebag_against_oracle(tmp_path, oracle):
    """tools/MakeBag.cxx semantics: ROI file in, one CSV row of 8*|scales| histograms'
    frequencies per ROI out; and the random-ROI mode writes a .ROIInfo that reads back."""
    shape = (30, 34, 64)      # nx % 32 == 0: the tool goes through the cropped masked smoothing
    img = synth.ct_like(shape, seed=61, n_blobs=8)
    lab = synth.lung_mask(shape).astype(np.uint16)
    m01 = synth.clamp01(lab.astype(np.uint8))
    d = str(tmp_path)
    nifti_util.write(d + "/img.nii.gz", img)
    nifti_util.write(d + "/mask.nii.gz", lab)
    sigmas = [0.6, 1.2]
    feats = np.concatenate([oracle.emphysema_features(img, m01, float(np.float32(s)), arith=1) for s in sigmas])
    edges = np.stack([synth.equalized_edges(feats[k][m01 != 0], 10) for k in range(16)])
    with open(d + "/hist.txt", "w") as f:
        f.write("# one row per (scale, feature)\n")
        for row in edges:
            f.write(",".join(repr(float(v)) for v in row) + "\n")
    rois = synth.random_rois(m01, 5, (9, 7, 5), seed=4)
    with open(d + "/rois.txt", "w") as f:
        f.write("header line\n")
        for r in rois:
            f.write("[%d, %d, %d][%d, %d, %d]\n" % tuple(r))
    p = run("MakeBag", "-i", d + "/img.nii.gz", "-m", d + "/mask.nii.gz", "-H", d + "/hist.txt", "-o", d,
            "-s", "0.6", "-s", "1.2", "-r", d + "/rois.txt", "-p", "case")
    assert p.returncode == 0, p.stderr
    assert "Got 5 rois." in p.stdout and "Skipping a line" in p.stdout
    bag = np.loadtxt(d + "/case.bag", delimiter=",", ndmin=2)
    assert bag.shape == (5, 16 * 11)
    counts = oracle.features_histograms(feats, m01, edges, rois).astype(np.float32)    # (5, 16, 11)
    freq = counts / counts.sum(2, keepdims=True)
    assert np.allclose(bag.reshape(5, 16, 11), freq, rtol=2e-5, atol=1e-7, equal_nan=True)
    assert np.allclose(np.nansum(bag.reshape(5, 16, 11), axis=2), 1.0, atol=1e-4)
    # random ROI mode
    p = run("MakeBag", "-i", d + "/img.nii.gz", "-m", d + "/mask.nii.gz", "-H", d + "/hist.txt", "-o", d,
            "-s", "0.6", "-s", "1.2", "-n", "7", "-x", "9", "-y", "7", "-z", "5", "-p", "rnd", "-S", "3")
    assert p.returncode == 0, p.stderr
    lines = open(d + "/rnd.ROIInfo").read().strip().splitlines()
    assert len(lines) == 7 and all(l.endswith("[9, 7, 5]") for l in lines)
    assert np.loadtxt(d + "/rnd.bag", delimiter=",", ndmin=2).shape == (7, 16 * 11)
    # wrong number of histogram rows -> the reference's error and exit code
    p = run("MakeBag", "-i", d + "/img.nii.gz", "-m", d + "/mask.nii.gz", "-H", d + "/hist.txt", "-o", d, "-s", "0.6",
            "-r", d + "/rois.txt")
    assert p.returncode == 1 and "Number of histograms must match" in p.stderr


@pytest.mark.gpu
def test_makebagonlyintensity(tmp_path, ctx):
    """tools/MakeBagOnlyIntensity.cxx semantics: one intensity histogram per ROI over its in-mask
    voxels; exactly one edge row accepted."""
    shape = (22, 26, 30)
    img = synth.ct_like(shape, seed=63, n_blobs=6)
    lab = synth.lung_mask(shape).astype(np.uint16)
    m01 = synth.clamp01(lab.astype(np.uint8))
    d = str(tmp_path)
    nifti_util.write(d + "/img.nii.gz", img)
    nifti_util.write(d + "/mask.nii.gz", lab)
    edges = synth.equalized_edges(img[m01 != 0], 14)
    open(d + "/hist.txt", "w").write(",".join(repr(float(v)) for v in edges) + "\n")
    rois = synth.random_rois(m01, 6, (9, 7, 5), seed=5)
    with open(d + "/rois.txt", "w") as f:
        f.write("header line\n")
        for r in rois:
            f.write("[%d, %d, %d][%d, %d, %d]\n" % tuple(r))
    p = run("MakeBagOnlyIntensity", "-i", d + "/img.nii.gz", "-m", d + "/mask.nii.gz", "-H", d + "/hist.txt", "-o", d,
            "-r", d + "/rois.txt", "-p", "int")
    assert p.returncode == 0, p.stderr
    bag = np.loadtxt(d + "/int.bag", delimiter=",", ndmin=2)
    assert bag.shape == (6, 15)
    for j, (x0, y0, z0, sx, sy, sz) in enumerate(rois):
        box, mb = img[z0:z0 + sz, y0:y0 + sy, x0:x0 + sx], m01[z0:z0 + sz, y0:y0 + sy, x0:x0 + sx] != 0
        cnt = np.bincount(np.searchsorted(edges, box[mb], side="left"), minlength=15).astype(np.float32)
        assert np.allclose(bag[j], cnt / np.float32(cnt.sum()), rtol=2e-5, atol=1e-7, equal_nan=True)   # 6 printed digits
    assert np.array_equal(ctx.intensity_roi_histograms(img, m01, edges, rois).sum(1),
                          [(m01[z:z + c, y:y + b, x:x + a] != 0).sum() for x, y, z, a, b, c in rois])
    open(d + "/two.txt", "w").write("0,1\n2,3\n")
    p = run("MakeBagOnlyIntensity", "-i", d + "/img.nii.gz", "-m", d + "/mask.nii.gz", "-H", d + "/two.txt", "-o", d,
            "-r", d + "/rois.txt")
    assert p.returncode == 1 and "Expected exactly one histogram" in p.stderr


@pytest.mark.gpu
def test_makebagdense_against_oracle(tmp_path, oracle):
    """tools/MakeBagDense.cxx semantics: one ROI per non-zero voxel of the ROI mask (here label 2
    of a ROI-mask file; then the image mask itself), written to .ROIInfo in raster order, and one
    CSV row of frequencies per ROI."""
    shape = (20, 24, 30)
    img = synth.ct_like(shape, seed=62, n_blobs=6)
    lab = synth.lung_mask(shape).astype(np.uint16)
    m01 = synth.clamp01(lab.astype(np.uint8))
    d = str(tmp_path)
    nifti_util.write(d + "/img.nii.gz", img)
    nifti_util.write(d + "/mask.nii.gz", lab)
    sigmas = [0.6, 1.2]
    feats = np.concatenate([oracle.emphysema_features(img, m01, float(np.float32(s)), arith=1) for s in sigmas])
    edges = np.stack([synth.equalized_edges(feats[k][m01 != 0], 10) for k in range(16)])
    with open(d + "/hist.txt", "w") as f:
        for row in edges:
            f.write(",".join(repr(float(v)) for v in row) + "\n")
    for tag, roimask, extra in (("lab2", (lab == 2), ["-M", d + "/mask.nii.gz", "-v", "2"]), ("all", m01 != 0, [])):
        p = run("MakeBagDense", "-i", d + "/img.nii.gz", "-m", d + "/mask.nii.gz", "-H", d + "/hist.txt", "-o", d,
                "-s", "0.6", "-s", "1.2", "-x", "7", "-y", "5", "-z", "3", "-p", tag, *extra)
        assert p.returncode == 0, p.stderr
        rois = synth.dense_rois(roimask, (7, 5, 3))
        lines = open("%s/%s.ROIInfo" % (d, tag)).read().strip().splitlines()
        assert lines == ["[%d, %d, %d][%d, %d, %d]" % tuple(r) for r in rois]
        bag = np.loadtxt("%s/%s.bag" % (d, tag), delimiter=",", ndmin=2)
        assert bag.shape == (len(rois), 16 * 11)
        pick = np.arange(0, len(rois), 11)
        counts = oracle.features_histograms(feats, m01, edges, rois[pick]).astype(np.float32)
        with np.errstate(invalid="ignore"):
            freq = counts / counts.sum(2, keepdims=True)
        assert np.allclose(bag[pick].reshape(-1, 16, 11), freq, rtol=2e-5, atol=1e-7, equal_nan=True)


@pytest.mark.gpu
def test_determine_bin_edges_tool_against_oracle(tmp_path, oracle, ctx):
    """tools/DetermineHistogramBinEdges_MultiScaleEigenvalueFeatures.cxx with -S 0 (all
    foreground voxels of two scans): edges == the reference's edge walk on the sorted oracle
    features; the output feeds MakeBag unchanged."""
    d = str(tmp_path)
    shapes = [(20, 24, 28), (18, 26, 22)]
    per_row = [[] for _ in range(16)]
    with open(d + "/pairs.txt", "w") as f:
        for i, shape in enumerate(shapes):
            img = synth.ct_like(shape, seed=70 + i, n_blobs=6)
            lab = synth.lung_mask(shape).astype(np.uint16)
            nifti_util.write("%s/img%d.nii.gz" % (d, i), img)
            nifti_util.write("%s/mask%d.nii.gz" % (d, i), lab)
            f.write("%s/img%d.nii.gz, %s/mask%d.nii.gz\n\n" % (d, i, d, i))
            m01 = synth.clamp01(lab.astype(np.uint8))
            for s, sigma in enumerate((0.6, 1.2)):
                feats = oracle.emphysema_features(img, m01, float(np.float32(sigma)), arith=1)
                for k in range(8):
                    per_row[s * 8 + k].append(feats[k][lab == 2])        # foreground value 2 only
    p = run("DetermineHistogramBinEdges_MultiScaleEigenvalueFeatures", "-i", d + "/pairs.txt", "-o", d + "/edges.txt",
            "-b", "11", "-S", "0", "-s", "0.6", "-s", "1.2", "-f", "2")
    assert p.returncode == 0, p.stderr
    lines = open(d + "/edges.txt").read().splitlines()
    assert lines[0].startswith("# Features: GaussianBlur") and lines[1] == "# Scales: 0.6 1.2"
    rows = [np.array([float(v) for v in l.split(",")]) for l in lines[2:]]
    assert len(rows) == 16 and all(len(r) == 10 for r in rows)
    for r, parts in zip(rows, per_row):
        s = np.sort(np.concatenate(parts).astype(np.float32))
        ref = oracle.determine_edges(s, 11, dtype=np.float32)
        assert np.allclose(r, ref, rtol=2e-5, atol=1e-12)
    # the device sort on its own, incl. duplicates, negative zero and infinities
    v = np.concatenate([np.random.default_rng(0).standard_normal(100001).astype(np.float32),
                        np.array([0.0, -0.0, np.inf, -np.inf, 1.0, 1.0], np.float32)])
    assert np.array_equal(ctx.sort(v), np.sort(v))
    assert ctx.sort(np.zeros(0, np.float32)).size == 0
