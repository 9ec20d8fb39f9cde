"""Host-side alignment bookkeeping, vectorised over particles.

Mirrors the Python half of the reference's per-particle loop (test_mref.py:184-207;
sp_alignment.Numrinit / ringwe / search_range, sp_utilities.combine_params2 /
inverse_transform2 on EMAN2's float32 2-D Transform).  The device does the
arithmetic-heavy part; this module only prepares the per-particle search request
and composes the returned parameters, exactly where the reference does it in Python.
"""
import math

import numpy as np

from .lib import SEARCH_DTYPE, RESULT_DTYPE

f32 = np.float32


def numrinit(first_ring, last_ring, skip=1):
    """sp_alignment.Numrinit(..., mode='F') -> flat [radius, 1-based offset, length]*nring."""
    MAXFFT = 32768
    numr = []
    lcirc = 1
    for k in range(first_ring, last_ring + 1, skip):
        jp = int(2 * math.pi * k + 0.5)
        ip = 2 ** (jp.bit_length())          # 2**(floor(log2 jp)+1)
        if k + skip <= last_ring and jp > ip + ip // 2:
            ip = min(MAXFFT, 2 * ip)
        if k + skip > last_ring and jp > ip + ip // 5:
            ip = min(MAXFFT, 2 * ip)
        numr += [k, lcirc, ip]
        lcirc += ip
    return np.array(numr, np.int32)


def ringwe(numr):
    nring = len(numr) // 3
    maxrin = float(numr[-1])
    return np.array([numr[3 * i] * 2 * math.pi / float(numr[3 * i + 2]) * maxrin / float(numr[3 * i + 2])
                     for i in range(nring)], np.float32)


def search_range(n, radius, shift, rng):
    """sp_alignment.search_range followed by the driver's swap: returns (left, right) arrays."""
    shift = np.asarray(shift, np.float64)
    cn = n // 2 + 1
    ql = np.maximum(cn + shift - radius - 2, 0.0)
    qe = np.maximum(n - cn - shift - radius, 0.0)
    return np.minimum(ql, rng), np.minimum(qe, rng)


# ---- EMAN2 2-D Transform (float32 3x4 matrix; x' = M T R x) -------------------
def _make(alpha, tx, ty, mirror):
    a = np.asarray(alpha, np.float64) * math.pi / 180.0
    c = np.cos(a).astype(f32)
    s = np.sin(a).astype(f32)
    m = np.empty(a.shape + (2, 3), f32)
    m[..., 0, 0] = c; m[..., 0, 1] = s; m[..., 0, 2] = np.asarray(tx, np.float64).astype(f32)
    m[..., 1, 0] = -s; m[..., 1, 1] = c; m[..., 1, 2] = np.asarray(ty, np.float64).astype(f32)
    mir = np.asarray(mirror).astype(bool)
    m[..., 0, :] = np.where(mir[..., None], -m[..., 0, :], m[..., 0, :])
    return m


def _mul(a, b):
    r = np.empty(np.broadcast_shapes(a.shape, b.shape), f32)
    for i in range(2):
        r[..., i, 0] = a[..., i, 0] * b[..., 0, 0] + a[..., i, 1] * b[..., 1, 0]
        r[..., i, 1] = a[..., i, 0] * b[..., 0, 1] + a[..., i, 1] * b[..., 1, 1]
        r[..., i, 2] = a[..., i, 0] * b[..., 0, 2] + a[..., i, 1] * b[..., 1, 2] + a[..., i, 2]
    return r


def _params(m):
    m = m.copy()
    det = m[..., 0, 0] * m[..., 1, 1] - m[..., 0, 1] * m[..., 1, 0]
    mir = det < 0
    m[..., 0, :] = np.where(mir[..., None], -m[..., 0, :], m[..., 0, :])
    a = np.degrees(np.arctan2(m[..., 0, 1].astype(np.float64), m[..., 0, 0].astype(np.float64)))
    a = np.where(a < 0, a + 360.0, a)
    a = np.where(a >= 360.0, a - 360.0, a)
    return a, m[..., 0, 2].astype(np.float64), m[..., 1, 2].astype(np.float64), mir.astype(np.int32)


def combine_params2(a1, sx1, sy1, m1, a2, sx2, sy2, m2):
    """sp_utilities.combine_params2: parameters of T2*T1 (T1 applied first)."""
    return _params(_mul(_make(a2, sx2, sy2, m2), _make(a1, sx1, sy1, m1)))


def inverse_transform2(alpha, tx=0.0, ty=0.0, mirror=0):
    """sp_utilities.inverse_transform2; Transform::invert works in double on the float entries."""
    m = _make(alpha, tx, ty, mirror).astype(np.float64)
    det = m[..., 0, 0] * m[..., 1, 1] - m[..., 0, 1] * m[..., 1, 0]
    r = np.empty_like(m)
    r[..., 0, 0] = m[..., 1, 1] / det; r[..., 0, 1] = -m[..., 0, 1] / det
    r[..., 1, 0] = -m[..., 1, 0] / det; r[..., 1, 1] = m[..., 0, 0] / det
    r[..., 0, 2] = -(r[..., 0, 0] * m[..., 0, 2] + r[..., 0, 1] * m[..., 1, 2])
    r[..., 1, 2] = -(r[..., 1, 0] * m[..., 0, 2] + r[..., 1, 1] * m[..., 1, 2])
    return _params(r.astype(f32))


def _native():
    """The C-ABI library's batched versions of the per-particle bookkeeping steps (csrc/cra_host.cu).  A missing
    library raises LibraryMissing: the numpy restatements below are reached only with native=False (the tests that
    pin the two against each other), never as a silent fallback."""
    from .lib import load_library
    return load_library()


def mref_search_request(params, nx, ou, xr, yr, native=True):
    """test_mref.py:184-198 for every particle.  params [P][4] (alpha, sx, sy, mirror), float64.
    Returns (search array, sxi, syi, params possibly reset to zero).  Runs in the native library
    (cra_mref_search_request) when it is loadable; the numpy body below is the same arithmetic and
    the one the golden Transform tests pin."""
    L = _native() if native else None
    if L is not None:
        params = np.array(params, np.float64, order="C").reshape(-1, 4)
        n = params.shape[0]
        s = np.zeros(n, SEARCH_DTYPE)
        sxi = np.zeros(n, np.float64)
        syi = np.zeros(n, np.float64)
        if L.cra_mref_search_request(n, params.ctypes.data, int(nx), int(ou), float(xr), float(yr),
                                     s.ctypes.data, sxi.ctypes.data, syi.ctypes.data) != 0:
            raise RuntimeError(L.cra_last_error().decode())
        return s, sxi, syi, params
    params = np.array(params, np.float64)
    cnx = nx // 2 + 1
    mashi = cnx - ou - 2
    _, sxi, syi, _ = inverse_transform2(params[:, 0], params[:, 1], params[:, 2])
    reset = (np.abs(sxi) > mashi) | (np.abs(syi) > mashi)
    sxi = np.where(reset, 0.0, sxi)
    syi = np.where(reset, 0.0, syi)
    params[reset] = 0.0
    s = np.zeros(params.shape[0], SEARCH_DTYPE)
    s["xl"], s["xr"] = search_range(nx, ou, sxi, xr)
    s["yl"], s["yr"] = search_range(nx, ou, syi, yr)
    s["cx"] = (cnx + sxi).astype(f32)
    s["cy"] = (cnx + syi).astype(f32)
    return s, sxi, syi, params


def reffree_search_request(params, cs, nx, ou, xr, yr, native=True):
    """ali2d_single_iter's per-particle prologue (test_reffree.py:780-783 -> Sphire):
    fold the average's centre shift cs into the parameters, invert, clamp to +-mashi."""
    params = np.array(params, np.float64, order="C").reshape(-1, 4)
    L = _native() if native else None
    if L is not None:
        n = params.shape[0]
        s = np.zeros(n, SEARCH_DTYPE)
        sxi = np.zeros(n, np.float64)
        syi = np.zeros(n, np.float64)
        if L.cra_reffree_search_request(n, params.ctypes.data, float(cs[0]), float(cs[1]), int(nx), int(ou), float(xr),
                                        float(yr), s.ctypes.data, sxi.ctypes.data, syi.ctypes.data) != 0:
            raise RuntimeError(L.cra_last_error().decode())
        return s, sxi, syi
    cnx = nx // 2 + 1
    mashi = cnx - ou - 2
    a, sx, sy, _ = combine_params2(params[:, 0], params[:, 1], params[:, 2], params[:, 3].astype(int),
                                   0.0, -cs[0], -cs[1], 0)
    _, sxi, syi, _ = inverse_transform2(a, sx, sy)
    sxi = np.clip(sxi, -mashi, mashi)
    syi = np.clip(syi, -mashi, mashi)
    s = np.zeros(params.shape[0], SEARCH_DTYPE)
    s["xl"], s["xr"] = search_range(nx, ou, sxi, xr)
    s["yl"], s["yr"] = search_range(nx, ou, syi, yr)
    s["cx"] = (cnx + sxi).astype(f32)
    s["cy"] = (cnx + syi).astype(f32)
    return s, sxi, syi


def compose_result(sxi, syi, res, native=True):
    """test_mref.py:206: combine_params2(0,-sxi,-syi,0, ang,sxs,sys,mirror) -> [P][4] float64."""
    L = _native() if native else None
    if L is not None and res.dtype == RESULT_DTYPE:
        sxi = np.ascontiguousarray(sxi, np.float64)
        syi = np.ascontiguousarray(syi, np.float64)
        res = np.ascontiguousarray(res)
        out = np.zeros((res.shape[0], 4), np.float64)
        if L.cra_compose_result(res.shape[0], sxi.ctypes.data, syi.ctypes.data, res.ctypes.data, out.ctypes.data) != 0:
            raise RuntimeError(L.cra_last_error().decode())
        return out
    z = np.zeros_like(sxi)
    a, sx, sy, m = combine_params2(z, -sxi, -syi, z.astype(int), res["ang"].astype(np.float64),
                                   res["sxs"].astype(np.float64), res["sys"].astype(np.float64), res["mirror"])
    return np.stack([a, sx, sy, m.astype(np.float64)], axis=1)


def mpi_start_end(n, p, i):
    """sp_applications.MPI_start_end (test_mref.py:90)."""
    return int(round(float(n) / p * i)), int(round(float(n) / p * (i + 1)))
