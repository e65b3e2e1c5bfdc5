"""Seeded synthetic volumes shared by tests/, bench.py and __graft_entry__.smoke().

CT-like image: background -800 HU, a few hundred Gaussian blobs (amplitude -600..400,
radius 3..12 voxels), a couple of bright tubes and plates (vessel / fissure like, so all
eigenvalue signatures occur) and N(0, 30^2) noise.  Lung-like mask: two ellipsoids minus
a central cylinder, labels {1, 2} (the tools clamp to {0, 1}).  All float32 / uint8,
layout (nz, ny, nx).
"""
import numpy as np


def ct_like(shape, seed=1, n_blobs=None, noise=30.0):
    nz, ny, nx = shape
    rng = np.random.default_rng(seed)
    vol = np.full(shape, -800.0, np.float32)
    if n_blobs is None:
        n_blobs = max(8, int(64 * (nz * ny * nx) / 128 ** 3))
    zz, yy, xx = np.meshgrid(np.arange(nz, dtype=np.float32), np.arange(ny, dtype=np.float32),
                             np.arange(nx, dtype=np.float32), indexing="ij", sparse=True)
    for _ in range(n_blobs):
        c = rng.uniform(0, 1, 3) * np.array([nz, ny, nx])
        r = rng.uniform(3, 12)
        a = rng.uniform(-600, 400)
        z0, z1 = int(max(0, c[0] - 3 * r)), int(min(nz, c[0] + 3 * r + 1))
        y0, y1 = int(max(0, c[1] - 3 * r)), int(min(ny, c[1] + 3 * r + 1))
        x0, x1 = int(max(0, c[2] - 3 * r)), int(min(nx, c[2] + 3 * r + 1))
        d2 = ((zz[z0:z1] - c[0]) ** 2 + (yy[:, y0:y1] - c[1]) ** 2 + (xx[:, :, x0:x1] - c[2]) ** 2)
        vol[z0:z1, y0:y1, x0:x1] += (a * np.exp(-d2 / (2 * r * r))).astype(np.float32)
    # a tube along z and a plate normal to y
    ty, tx = ny * 0.37, nx * 0.61
    vol += (500.0 * np.exp(-((yy - ty) ** 2 + (xx - tx) ** 2) / (2 * 2.0 ** 2))).astype(np.float32)
    vol += (300.0 * np.exp(-((yy - ny * 0.7) ** 2) / (2 * 1.5 ** 2))).astype(np.float32)
    if noise > 0:
        vol += rng.normal(0, noise, shape).astype(np.float32)
    return np.ascontiguousarray(vol, np.float32)


def lung_mask(shape, labels=True):
    nz, ny, nx = shape
    zz, yy, xx = np.meshgrid(np.arange(nz, dtype=np.float32), np.arange(ny, dtype=np.float32),
                             np.arange(nx, dtype=np.float32), indexing="ij", sparse=True)
    cz, cy = nz / 2.0, ny / 2.0
    az, ay, ax = 0.375 * nz, 0.273 * ny, 0.195 * nx
    left = ((zz - cz) / az) ** 2 + ((yy - cy) / ay) ** 2 + ((xx - (nx / 2.0 - 0.234 * nx)) / ax) ** 2 <= 1
    right = ((zz - cz) / az) ** 2 + ((yy - cy) / ay) ** 2 + ((xx - (nx / 2.0 + 0.234 * nx)) / ax) ** 2 <= 1
    cyl = ((yy - cy) ** 2 + (xx - nx / 2.0) ** 2 <= (0.06 * nx) ** 2) & (zz >= 0)
    m = np.zeros(shape, np.uint8)
    m[left & ~cyl] = 1
    m[right & ~cyl] = 2 if labels else 1
    return m


def clamp01(mask):
    """itk::ClampImageFilter(0, 1) as the tools apply to the mask (ExtractFeatures.cxx:99-104)."""
    return np.minimum(mask, 1).astype(np.uint8)


def random_rois(mask, n, size, seed=7):
    """n boxes {x0,y0,z0,sx,sy,sz} of `size` centred on in-mask voxels, fully inside."""
    nz, ny, nx = mask.shape
    sx, sy, sz = size
    rng = np.random.default_rng(seed)
    ok = np.zeros_like(mask, bool)
    ok[sz // 2: nz - (sz - sz // 2) + 1, sy // 2: ny - (sy - sy // 2) + 1,
       sx // 2: nx - (sx - sx // 2) + 1] = True
    cand = np.argwhere((mask != 0) & ok)
    pick = cand[rng.choice(len(cand), n, replace=len(cand) < n)]
    return np.array([[x - sx // 2, y - sy // 2, z - sz // 2, sx, sy, sz] for z, y, x in pick], np.int32)


def dense_rois(mask, size):
    """DenseROIGenerator (include/ife/ROI/DenseROIGenerator.hxx:22-45): one box of `size` per
    non-zero mask voxel, start = index - size/2, kept when it lies inside the image; raster
    order (x fastest)."""
    nz, ny, nx = mask.shape
    sx, sy, sz = size
    out = []
    for z, y, x in np.argwhere(mask != 0):          # argwhere is z-major, x fastest: ITK order
        x0, y0, z0 = x - sx // 2, y - sy // 2, z - sz // 2
        if x0 >= 0 and y0 >= 0 and z0 >= 0 and x0 + sx <= nx and y0 + sy <= ny and z0 + sz <= nz:
            out.append([x0, y0, z0, sx, sy, sz])
    return np.array(out, np.int32).reshape(-1, 6)


def equalized_edges(samples, n_edges):
    """n_edges equal-frequency edges from samples (strictly increasing where possible)."""
    q = np.quantile(samples.astype(np.float64), (np.arange(n_edges) + 1) / (n_edges + 1.0))
    e = q.astype(np.float32)
    for i in range(1, n_edges):  # keep them sorted and distinct
        if not e[i] > e[i - 1]:
            e[i] = np.nextafter(e[i - 1], np.float32(np.inf))
    return e


def special_matrices(n, seed=3):
    """Random / near-degenerate / degenerate / diagonal symmetric 3x3 matrices (n, 6) float32."""
    rng = np.random.default_rng(seed)
    k = n // 5
    parts = [rng.standard_normal((k, 6))]
    parts.append(rng.standard_normal((k, 6)) * np.float32(1000.0))
    # a*I + small noise
    a = rng.standard_normal((k, 1))
    m = np.zeros((k, 6)); m[:, [0, 3, 5]] = a
    parts.append(m + 1e-3 * rng.standard_normal((k, 6)))
    # exactly two equal eigenvalues: u u^T + b I
    u = rng.standard_normal((k, 3)); b = rng.standard_normal((k, 1))
    m = np.stack([u[:, 0] * u[:, 0], u[:, 0] * u[:, 1], u[:, 0] * u[:, 2], u[:, 1] * u[:, 1],
                  u[:, 1] * u[:, 2], u[:, 2] * u[:, 2]], 1)
    m[:, [0, 3, 5]] += b
    parts.append(m)
    # diagonal incl. ties and zeros
    d = np.zeros((n - 4 * k, 6)); d[:, [0, 3, 5]] = rng.integers(-3, 4, (n - 4 * k, 3))
    parts.append(d)
    return np.ascontiguousarray(np.concatenate(parts, 0), np.float32)
