"""Minimal image-stack I/O for the drivers: .npy and MRC2014 mode-2 (.mrc/.mrcs) stacks in,
per-iteration reference stacks and text parameter rows out.  EMAN2 HDF/BDB containers are the
reference's formats (test_mref.py:155, :285, :304-313) and are out of scope here."""
import os
import struct

import numpy as np


def _open_stack(path):
    """Memory-mapped view [n][ny][nx] of a stack file (nothing is read until it is sliced)."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        a = np.load(path, mmap_mode="r")
    elif ext in (".mrc", ".mrcs", ".st"):
        with open(path, "rb") as f:
            hdr = f.read(1024)
        nx, ny, nz, mode = struct.unpack("<4i", hdr[:16])
        nsymbt = struct.unpack("<i", hdr[92:96])[0]
        if mode != 2:
            raise ValueError("only MRC mode 2 (float32) stacks are supported, got mode %d" % mode)
        a = np.memmap(path, dtype="<f4", mode="r", offset=1024 + nsymbt, shape=(nz, ny, nx))
    else:
        raise ValueError("unsupported stack format '%s' (use .npy or .mrcs)" % ext)
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3 or a.shape[1] != a.shape[2]:
        raise ValueError("stack must be [n][nx][nx] with square images")
    return a


def stack_shape(path):
    """(number of images, nx) without reading the pixel data."""
    a = _open_stack(path)
    return int(a.shape[0]), int(a.shape[2])


def read_stack(path, start=None, stop=None):
    """Images [start, stop) of the stack as a contiguous float32 array.  Each rank of a multi-GPU run reads only its
    own share (the reference ships images between ranks instead, test_mref_gpu_align.py:1380-1415)."""
    a = _open_stack(path)
    return np.ascontiguousarray(a[slice(start, stop)], np.float32)


def write_stack(path, a):
    a = np.ascontiguousarray(a, np.float32)
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        np.save(path, a)
        return
    nz, ny, nx = a.shape
    hdr = bytearray(1024)
    struct.pack_into("<4i", hdr, 0, nx, ny, nz, 2)
    struct.pack_into("<3i", hdr, 28, nx, ny, nz)
    struct.pack_into("<3f", hdr, 40, float(nx), float(ny), float(nz))
    struct.pack_into("<3f", hdr, 52, 90.0, 90.0, 90.0)
    struct.pack_into("<3i", hdr, 64, 1, 2, 3)
    struct.pack_into("<3f", hdr, 76, float(a.min()), float(a.max()), float(a.mean()))
    hdr[208:212] = b"MAP "
    hdr[212:216] = bytes([0x44, 0x44, 0, 0])
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(a.tobytes())


def write_params(path, params, assign=None, first_index=0):
    """Rows 'idx angle sx sy mirror class' (the layout src/utils_ralib.py:31-32 parses)."""
    with open(path, "w") as f:
        for i, p in enumerate(params):
            cls = int(assign[i]) if assign is not None else 0
            f.write("%d %.6f %.6f %.6f %d %d\n" % (first_index + i, p[0], p[1], p[2], int(p[3]), cls))


def read_params(path):
    """Rows 'idx angle sx sy mirror class' (write_params; src/utils_ralib.py:31-32) or the 4-column
    'angle sx sy mirror' rows of initial2Dparams.txt -> (params [n][4] float64, classes [n] int or None)."""
    rows = np.loadtxt(path, ndmin=2)
    if rows.shape[1] >= 6:
        return rows[:, 1:5].astype(np.float64), rows[:, 5].astype(np.int64)
    if rows.shape[1] == 4:
        return rows.astype(np.float64), None
    raise ValueError("parameter file must have 4 or 6 columns, got %d" % rows.shape[1])


def write_fsc(path, frsc):
    """The text file sp_statistics.fsc writes when given a name (write_text_file of [freq, fsc, n]; the reference's
    drm%03d%04d.txt, test_mref.py:254): one row per shell."""
    with open(path, "w") as f:
        for fr, v, n in zip(*frsc):
            f.write("%12.5g  %12.5g  %12.5g\n" % (fr, v, n))
