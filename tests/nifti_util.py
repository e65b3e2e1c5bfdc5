"""Tiny independent NIfTI-1 reader/writer (numpy + gzip) used to cross-check the C++ IO of
the host tools."""
import gzip
import struct

import numpy as np

_DT = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 512: np.uint16}
_CODE = {np.dtype(v): k for k, v in _DT.items()}


def write(path, vol, spacing=(1.0, 1.0, 1.0)):
    vol = np.ascontiguousarray(vol)
    nz, ny, nx = vol.shape
    h = bytearray(352)
    struct.pack_into("<i", h, 0, 348)
    struct.pack_into("<8h", h, 40, 3, nx, ny, nz, 1, 1, 1, 1)
    struct.pack_into("<h", h, 70, _CODE[vol.dtype])
    struct.pack_into("<h", h, 72, vol.dtype.itemsize * 8)
    struct.pack_into("<8f", h, 76, 1.0, spacing[0], spacing[1], spacing[2], 0, 0, 0, 0)
    struct.pack_into("<f", h, 108, 352.0)
    struct.pack_into("<f", h, 112, 1.0)
    struct.pack_into("<h", h, 254, 1)
    struct.pack_into("<4f", h, 280, spacing[0], 0, 0, 0)
    struct.pack_into("<4f", h, 296, 0, spacing[1], 0, 0)
    struct.pack_into("<4f", h, 312, 0, 0, spacing[2], 0)
    h[344:348] = b"n+1\0"
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "wb") as f:
        f.write(bytes(h))
        f.write(vol.tobytes())


def read(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        raw = f.read()
    assert struct.unpack_from("<i", raw, 0)[0] == 348 and raw[344:347] == b"n+1"
    dim = struct.unpack_from("<8h", raw, 40)
    dt = _DT[struct.unpack_from("<h", raw, 70)[0]]
    pixdim = struct.unpack_from("<8f", raw, 76)
    off = int(struct.unpack_from("<f", raw, 108)[0])
    nx, ny, nz = dim[1:4]
    vol = np.frombuffer(raw, dt, nx * ny * nz, off).reshape(nz, ny, nx)
    return vol, pixdim[1:4]
