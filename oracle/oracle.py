"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of ``cra_oracle.c`` plus a numpy restatement of the Python /
EMAN2 pieces of the reference's per-iteration reference update
(test_mref.py:238-296; test_reffree.py:695-755).  The arithmetic follows EMAN2
2.31 / Sphire (un-vendored dependency, README.md:43) as specified in SURVEY.md
Appendix A.  Parity status: Transform algebra pinned by the reference's golden
tuples (cuda/EMAN2_test.ipynb cells 23-25); the conventions and discrete answers
of the alignment (class, mirror, integer shift, angle) pinned to AlignParam[] as
the reference's own CUDA library left it for a deterministic stack
(tests/golden/refcuda_mref_outputs.npz: 512 / 512 identical, angles within
0.12 degrees); the digits at the multiref_polar_ali_2d boundary remain "parity
unpinned" (no EMAN2 vectors exist, EMAN2 cannot run here, and the reference's
CUDA is gpu_isac's arithmetic).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.
"""
import ctypes as C
import math
import os
import random
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_d = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libcra_oracle.so")
    src = os.path.join(_HERE, "cra_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcra_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libcra_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.cra_o_numrinit.argtypes = [C.c_int, C.c_int, C.c_int, _i, C.c_int]
        L.cra_o_numrinit.restype = C.c_int
        L.cra_o_ringwe.argtypes = [_i, C.c_int, _f]
        L.cra_o_polar2dm.argtypes = [_f, C.c_int, C.c_int, C.c_float, C.c_float, _i, C.c_int, _f]
        L.cra_o_normalize_ring.argtypes = [_f, _i, C.c_int]
        L.cra_o_frngs.argtypes = [_f, _i, C.c_int]
        L.cra_o_applyws.argtypes = [_f, _i, C.c_int, _f]
        L.cra_o_rfft_packed.argtypes = [_f, C.c_int]
        L.cra_o_irfft_packed.argtypes = [_d, C.c_int]
        L.cra_o_crosrng_ms.argtypes = [_f, _f, _i, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float),
                                       C.POINTER(C.c_double), C.POINTER(C.c_float), _d, _d]
        L.cra_o_multiref_polar_ali_2d.argtypes = [_f, C.c_int, C.c_int, _f, C.c_int,
                                                  C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                                  _i, C.c_int, C.c_float, C.c_float, C.c_int, _f]
        L.cra_o_combine_params2.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int,
                                            C.c_double, C.c_double, C.c_double, C.c_int, _d]
        L.cra_o_inverse_transform2.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, _d]
        L.cra_o_rot_shift2d.argtypes = [_f, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, _f]
        L.cra_o_model_circle.argtypes = [C.c_float, C.c_int, C.c_int, _f]
        L.cra_o_normalize_mask.argtypes = [_f, _f, C.c_int, C.c_int]
        L.cra_o_search_range.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _d]
        L.cra_o_prepare_refs.argtypes = [_f, C.c_int, C.c_int, _f, _i, C.c_int, _f]
        L.cra_o_mref_iteration.argtypes = [_f, C.c_int, C.c_int, _f, _f, C.c_int, _i, C.c_int,
                                           C.c_double, C.c_double, C.c_double, C.c_int,
                                           _d, _i, _f, _f, _d, C.c_long, C.c_int, C.c_int]
        L.cra_o_align_batch.argtypes = [_f, C.c_int, C.c_int, _f, C.c_int, _i, C.c_int,
                                        _f, _f, C.c_float, C.c_int, _f, C.c_int]
        L.cra_o_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def max_threads():
    return int(lib().cra_o_max_threads())


# ----------------------------------------------------------------- rings
def numrinit(ir, ou, rs=1):
    numr = np.zeros(3 * 4096, np.int32)
    n = lib().cra_o_numrinit(int(ir), int(ou), int(rs), numr, 4096)
    assert n > 0
    return numr[:3 * n].copy()


def ringwe(numr):
    nring = len(numr) // 3
    wr = np.zeros(nring, np.float32)
    lib().cra_o_ringwe(numr, nring, wr)
    return wr


def lcirc_of(numr):
    return int(numr[-2] + numr[-1] - 1)


def polar2dm(img, cx, cy, numr):
    img = np.ascontiguousarray(img, np.float32)
    out = np.zeros(lcirc_of(numr), np.float32)
    lib().cra_o_polar2dm(img, img.shape[1], img.shape[0], cx, cy, numr, len(numr) // 3, out)
    return out


def normalize_ring(circ, numr):
    c = np.array(circ, np.float32)
    lib().cra_o_normalize_ring(c, numr, len(numr) // 3)
    return c


def frngs(circ, numr):
    c = np.array(circ, np.float32)
    lib().cra_o_frngs(c, numr, len(numr) // 3)
    return c


def applyws(circ, numr, wr):
    c = np.array(circ, np.float32)
    lib().cra_o_applyws(c, numr, len(numr) // 3, np.ascontiguousarray(wr, np.float32))
    return c


def rfft_packed(x):
    c = np.array(x, np.float32)
    lib().cra_o_rfft_packed(c, len(c))
    return c


def irfft_packed(x):
    c = np.array(x, np.float64)
    lib().cra_o_irfft_packed(c, len(c))
    return c


def crosrng_ms(ref, img, numr):
    maxrin = int(numr[-1])
    qn, qm = C.c_double(), C.c_double()
    tot, tmt = C.c_float(), C.c_float()
    q = np.zeros(maxrin, np.float64)
    t = np.zeros(maxrin, np.float64)
    lib().cra_o_crosrng_ms(np.ascontiguousarray(ref, np.float32), np.ascontiguousarray(img, np.float32),
                           numr, len(numr) // 3, C.byref(qn), C.byref(tot), C.byref(qm), C.byref(tmt), q, t)
    return dict(qn=qn.value, tot=tot.value, qm=qm.value, tmt=tmt.value, q=q, t=t)


def multiref_polar_ali_2d(img, crefim, xrng, yrng, step, numr, cnx, cny, normalize=True):
    """xrng/yrng are [left, right] lists as the driver passes them
    (test_mref.py:195-201).  Returns [ang, sxs, sys, mirror, iref, peak, -ix, -iy]."""
    img = np.ascontiguousarray(img, np.float32)
    crefim = np.ascontiguousarray(crefim, np.float32)
    R = crefim.shape[0] if crefim.ndim == 2 else 1
    out = np.zeros(8, np.float32)
    lib().cra_o_multiref_polar_ali_2d(img, img.shape[1], img.shape[0], crefim, R,
                                      xrng[0], xrng[1], yrng[0], yrng[1], step,
                                      numr, len(numr) // 3, cnx, cny, int(normalize), out)
    return out


# ------------------------------------------------------------- transforms
def combine_params2(a1, sx1, sy1, m1, a2, sx2, sy2, m2):
    out = np.zeros(4)
    lib().cra_o_combine_params2(a1, sx1, sy1, int(m1), a2, sx2, sy2, int(m2), out)
    return out[0], out[1], out[2], int(out[3])


def inverse_transform2(alpha, tx=0.0, ty=0.0, mirror=0):
    out = np.zeros(4)
    lib().cra_o_inverse_transform2(alpha, tx, ty, int(mirror), out)
    return out[0], out[1], out[2], int(out[3])


def rot_shift2d(img, alpha, sx, sy, mirror):
    img = np.ascontiguousarray(img, np.float32)
    out = np.zeros_like(img)
    lib().cra_o_rot_shift2d(img, img.shape[1], img.shape[0], alpha, sx, sy, int(mirror), out)
    return out


def model_circle(r, nx, ny=None):
    ny = ny or nx
    m = np.zeros((ny, nx), np.float32)
    lib().cra_o_model_circle(r, nx, ny, m)
    return m


def normalize_mask(img, mask, no_sigma):
    c = np.array(img, np.float32)
    lib().cra_o_normalize_mask(c, np.ascontiguousarray(mask, np.float32), c.size, int(no_sigma))
    return c


def search_range(n, radius, shift, rng):
    """[left, right] after the driver's swap (test_mref.py:195-198)."""
    lr = np.zeros(2)
    lib().cra_o_search_range(n, radius, shift, rng, lr)
    return [lr[0], lr[1]]


def prepare_refs(refs, mask, numr):
    """test_mref.py:170-175.  Returns (normalised refs, weighted spectra [R][lcirc])."""
    refs = np.array(refs, np.float32)
    R, nx = refs.shape[0], refs.shape[-1]
    cref = np.zeros((R, lcirc_of(numr)), np.float32)
    lib().cra_o_prepare_refs(refs, R, nx, np.ascontiguousarray(mask, np.float32), numr, len(numr) // 3, cref)
    return refs, cref


def mref_iteration(images, mask, cref, numr, xrng, yrng, step, ou, params, gofs=0, normalize=True,
                   nthreads=1):
    """Per-particle section of one mref_ali2d_MPI iteration (test_mref.py:183-215).
    images are modified in place (normalize.mask, as the reference does)."""
    P, nx = images.shape[0], images.shape[-1]
    R = cref.shape[0]
    params = np.ascontiguousarray(params, np.float64).copy()
    assign = np.zeros(P, np.int32)
    peak = np.zeros(P, np.float32)
    sums = np.zeros((R, 2, nx, nx), np.float32)
    counts = np.zeros(R, np.float64)
    lib().cra_o_mref_iteration(images, P, nx, np.ascontiguousarray(mask, np.float32), cref, R, numr,
                               len(numr) // 3, xrng, yrng, step, int(ou), params, assign, peak,
                               sums, counts, int(gofs), int(normalize), int(nthreads))
    return params, assign, peak, sums, counts


def align_batch(images, cref, numr, centres, win, step, normalize=True, nthreads=1):
    P, nx = images.shape[0], images.shape[-1]
    out = np.zeros((P, 8), np.float32)
    lib().cra_o_align_batch(np.ascontiguousarray(images, np.float32), P, nx,
                            np.ascontiguousarray(cref, np.float32), cref.shape[0], numr, len(numr) // 3,
                            np.ascontiguousarray(centres, np.float32), np.ascontiguousarray(win, np.float32),
                            step, int(normalize), out, int(nthreads))
    return out


# ------------------------------------------------- reference update (numpy)
def _round(x):
    return int(x + 0.5) if x >= 0 else int(x - 0.5)


def fsc(img1, img2, w=1.0):
    """EMData::calc_fourier_shell_correlation via sp_statistics.fsc (test_mref.py:254).
    Returns [freq, fsc, n] lists."""
    ny, nx = img1.shape
    F = np.fft.rfft2(img1.astype(np.float64))
    G = np.fft.rfft2(img2.astype(np.float64))
    nx2, ny2 = nx // 2, ny // 2
    dx2 = 1.0 / nx2 / nx2
    dy2 = 1.0 / ny2 / ny2
    inc = _round(max(nx2, ny2) / w)
    ret = np.zeros(inc + 1); n1 = np.zeros(inc + 1); n2 = np.zeros(inc + 1); lr = np.zeros(inc + 1)
    for iy in range(ny):
        ky = iy - ny if iy > ny2 else iy
        for kx in range(nx // 2 + 1):
            if kx > 0 or ky >= 0:
                argx = 0.5 * math.sqrt(np.float32(ky * ky * dy2 + kx * kx * dx2))
                r = _round(inc * 2 * argx)
                if r <= inc:
                    a, b = F[iy, kx], G[iy, kx]
                    ret[r] += a.real * b.real + a.imag * b.imag
                    n1[r] += a.real * a.real + a.imag * a.imag
                    n2[r] += b.real * b.real + b.imag * b.imag
                    lr[r] += 2
    freq, val, cnt = [], [], []
    for i in range(inc + 1):
        if lr[i] > 0:
            freq.append(float(i) / float(2 * inc))
            den = math.sqrt(n1[i] * n2[i])
            val.append(float(np.float32(ret[i] / den)) if den > 0 else 0.0)
            cnt.append(lr[i])
    return [freq, val, cnt]


def amoeba(var, scale, func, ftolerance=1.e-4, xtolerance=1.e-4, itmax=500, data=None):
    """sp_utilities.amoeba: simplex MAXIMISER."""
    nvar = len(var)
    nsimplex = nvar + 1
    simplex = [0] * (nvar + 1)
    simplex[0] = var[:]
    for i in range(nvar):
        simplex[i + 1] = var[:]
        simplex[i + 1][i] += scale[i]
    fvalue = [func(simplex[i], data=data) for i in range(nsimplex)]
    iteration = 0
    while True:
        ssworst = 0
        ssbest = 0
        for i in range(nsimplex):
            if fvalue[i] > fvalue[ssbest]:
                ssbest = i
            if fvalue[i] < fvalue[ssworst]:
                ssworst = i
        pavg = [0.0] * nvar
        for i in range(nsimplex):
            if i != ssworst:
                for j in range(nvar):
                    pavg[j] += simplex[i][j]
        for j in range(nvar):
            pavg[j] = pavg[j] / nvar
        simscale = 0.0
        for i in range(nvar):
            simscale += abs(pavg[i] - simplex[ssworst][i]) / scale[i]
        simscale = simscale / nvar
        fscale = (abs(fvalue[ssbest]) + abs(fvalue[ssworst])) / 2.0
        frange = abs(fvalue[ssbest] - fvalue[ssworst]) / fscale if fscale != 0.0 else 0.0
        if (((ftolerance <= 0.0 or frange < ftolerance) and (xtolerance <= 0.0 or simscale < xtolerance))
                or (itmax and iteration >= itmax)):
            return simplex[ssbest], fvalue[ssbest], iteration
        pnew = [2.0 * pavg[i] - simplex[ssworst][i] for i in range(nvar)]
        fnew = func(pnew, data=data)
        if fnew <= fvalue[ssworst]:
            for i in range(nsimplex):
                if i != ssbest and i != ssworst:
                    for j in range(nvar):
                        simplex[i][j] = 0.5 * simplex[ssbest][j] + 0.5 * simplex[i][j]
                    fvalue[i] = func(simplex[i], data=data)
            for j in range(nvar):
                pnew[j] = 0.5 * simplex[ssbest][j] + 0.5 * simplex[ssworst][j]
            fnew = func(pnew, data=data)
        elif fnew >= fvalue[ssbest]:
            pnew2 = [3.0 * pavg[i] - 2.0 * simplex[ssworst][i] for i in range(nvar)]
            fnew2 = func(pnew2, data=data)
            if fnew2 > fnew:
                pnew = pnew2
                fnew = fnew2
        for i in range(nvar):
            simplex[ssworst][i] = pnew[i]
        fvalue[ssworst] = fnew
        iteration += 1


def fit_tanh(dres, low=0.1):
    """sp_filter.fit_tanh; dres is mutated exactly as upstream does."""
    def fit_tanh_func(args, data):
        v = 0.0
        if data[1][0] < 0.0:
            data[1][0] *= -1.0
        for i in range(len(data[0])):
            f = 2 * data[1][i] / (1.0 + data[1][i])
            if args[0] == 0 or args[1] == 0:
                qt = 0
            else:
                qt = f - 0.5 * (math.tanh(math.pi * (data[0][i] + args[0]) / 2.0 / args[1] / args[0])
                                - math.tanh(math.pi * (data[0][i] - args[0]) / 2.0 / args[1] / args[0]))
            v -= qt * qt
        return v

    setzero = False
    for i in range(1, len(dres[0])):
        if not setzero:
            if 2 * dres[1][i] / (1.0 + dres[1][i]) < low:
                setzero = True
        if setzero:
            dres[1][i] = 0.0
    freq = -1.0
    for i in range(1, len(dres[0]) - 1):
        if (2 * dres[1][i] / (1.0 + dres[1][i])) < 0.5:
            freq = dres[0][i - 1]
            break
    if freq < 0.0:
        return 0.4, 0.2
    result = amoeba([freq, 0.1], [0.05, 0.05], fit_tanh_func, data=dres)
    return result[0][0], result[0][1]


def _freq_grid(ny, nx):
    iy = np.arange(ny)
    ky = (np.where(iy > ny // 2, iy - ny, iy) / float(ny))[:, None]   # EMAN2: ky = iy-ny only for iy > ny/2
    kx = (np.arange(nx // 2 + 1) / float(nx))[None, :]
    return ky, kx


def filt_tanl(img, fl, aa):
    """sp_filter.filt_tanl (TANH_LOW_PASS, no padding); same H as gpu_aln_noref.cu:799-814."""
    ny, nx = img.shape
    ky, kx = _freq_grid(ny, nx)
    d = np.sqrt(kx * kx + ky * ky)
    c = math.pi / (2.0 * aa * fl)
    H = 0.5 * (np.tanh(c * (d + fl)) - np.tanh(c * (d - fl)))
    return np.fft.irfft2(np.fft.rfft2(img.astype(np.float64)) * H, s=(ny, nx)).astype(np.float32)


def fshift(img, sx, sy):
    """sp_fundamentals.fshift: circular Fourier shift by (+sx, +sy)."""
    ny, nx = img.shape
    ky, kx = _freq_grid(ny, nx)
    ph = np.exp(-2j * math.pi * (kx * sx + ky * sy))
    return np.fft.irfft2(np.fft.rfft2(img.astype(np.float64)) * ph, s=(ny, nx)).astype(np.float32)


def phase_cog(img):
    """EMData::phase_cog (2-D), centre-relative (notebook/00 log: 'Center x = 2.354 ...')."""
    ny, nx = img.shape
    out = []
    for marg, n in ((img.sum(axis=0, dtype=np.float64), nx), (img.sum(axis=1, dtype=np.float64), ny)):
        P = 2 * math.pi / n
        i = np.arange(n)
        Cc = float(np.sum(np.cos(P * i) * marg))
        Ss = float(np.sum(np.sin(P * i) * marg))
        F1 = math.atan2(Ss, Cc)
        if F1 < 0.0:
            F1 += 2 * math.pi
        out.append(F1 / P + 1.0 - (n // 2 + 1))
    return out


def center_2d(img, center):
    if center == 1:
        cs = phase_cog(img)
        return fshift(img, -cs[0], -cs[1]), cs
    if center == 0:
        return img, [0.0, 0.0]
    raise NotImplementedError("only center methods 0, 1 (and -1 in the ref-free driver) are on the path")


def ref_ali2d(mask, center, avg, frsc):
    """sp_user_functions.ref_ali2d (test_mref.py:273-276)."""
    fl, aa = fit_tanh(frsc)
    aa = min(aa, 0.2)
    fl = max(min(0.4, fl), 0.12)
    tavg = filt_tanl(avg, fl, aa)
    tavg, cs = center_2d(tavg, center)
    return tavg, cs, (fl, aa)


def update_refs(sums, counts, images, mask, center=1, rng=None):
    """Rank-0 section of mref_ali2d_MPI (test_mref.py:238-286).  sums [R][2][nx][nx],
    counts [R]; images are the (mask-normalised) particles used for the <4-member
    reseed.  Returns (new refs [R][nx][nx], info dict)."""
    R = sums.shape[0]
    rng = rng or random.Random(1000)
    refs = np.zeros((R,) + sums.shape[2:], np.float32)
    ave_fsc = None
    c_fsc = 0
    frsc = None
    reseeded = {}
    for j in range(R):
        if counts[j] < 4:
            k = rng.randint(0, images.shape[0] - 1)
            reseeded[j] = k
            refs[j] = images[k]
        else:
            frsc = fsc(sums[j, 0], sums[j, 1], 1.0)
            refs[j] = (sums[j, 0] + sums[j, 1]) * np.float32(1.0 / float(counts[j]))
            if ave_fsc is None:
                ave_fsc = list(frsc[1]); c_fsc = 1
            else:
                for i in range(len(frsc[1])):
                    ave_fsc[i] += frsc[1][i]
                c_fsc += 1
    if ave_fsc is not None and sum(ave_fsc) != 0:
        for i in range(len(ave_fsc)):
            ave_fsc[i] /= float(c_fsc)
            frsc[1][i] = ave_fsc[i]
    filt = None
    css = []
    for j in range(R):
        refs[j], cs, filt = ref_ali2d(mask, center, refs[j], frsc)
        css.append(cs)
        refs[j] = normalize_mask(refs[j], mask, 1)
    return refs, dict(frsc=frsc, filter=filt, cs=css, reseeded=reseeded)


def mref_ali2d(images, refs, ir=1, ou=-1, rs=1, xr=0, yr=0, ts=1, center=1, maxit=10, rand_seed=1000,
               nthreads=1, params=None):
    """Whole mref_ali2d_MPI loop on one rank (test_mref.py:48-315), in memory."""
    images = np.array(images, np.float32)
    P, nx = images.shape[0], images.shape[-1]
    if ou == -1:
        ou = nx // 2 - 2
    mask = model_circle(ou, nx)
    numr = numrinit(ir, ou, rs)
    refs = np.array(refs, np.float32)
    params = np.zeros((P, 4)) if params is None else np.array(params, np.float64)
    rng = random.Random(rand_seed)
    history = []
    for it in range(maxit):
        refs_n, cref = prepare_refs(refs, mask, numr)
        params, assign, peak, sums, counts = mref_iteration(images, mask, cref, numr, xr, yr, ts, ou,
                                                            params, 0, True, nthreads)
        refs, info = update_refs(sums, counts, images, mask, center, rng)
        history.append(dict(params=params.copy(), assign=assign.copy(), peak=peak.copy(),
                            sums=sums, counts=counts.copy(), refs=refs.copy(), info=info))
    return params, assign, refs, history


def fsc_mask(img1, img2, mask):
    """sp_statistics.fsc_mask (test_reffree.py:708)."""
    m = mask > 0.5
    a = (img1 - np.float32(img1[m].astype(np.float64).mean())) * mask
    b = (img2 - np.float32(img2[m].astype(np.float64).mean())) * mask
    return fsc(a, b)


def _schedule(xr, yr, ts):
    """get_input_from_string semantics: "4 2 1" / [4, 2, 1] / 4 -> per-step (xr, yr, ts); yr = -1 copies xr,
    a shorter ts list repeats its last entry."""
    def lst(v):
        if isinstance(v, str):
            return [float(x) for x in v.split()]
        try:
            return [float(x) for x in v]
        except TypeError:
            return [float(v)]
    xs, ys, tss = lst(xr), lst(yr), lst(ts)
    if len(ys) == 1 and ys[0] == -1:
        ys = list(xs)
    n = len(xs)
    ys = ys + [ys[-1]] * (n - len(ys))
    tss = tss + [tss[-1]] * (n - len(tss))
    return [(xs[i], ys[i], tss[i]) for i in range(n)]


def ali2d_base(images, ir=1, ou=-1, rs=1, xr=0, yr=0, ts=1, center=-1, maxit=10, nthreads=1):
    """Reference-free alignment, CPU twin (ali2d_base, test_reffree.py:515-837; the per-particle
    step is Sphire's ali2d_single_iter -> ormq: no ring normalisation, clamped shifts, the
    average's centre shift cs folded into the parameters)."""
    images = np.array(images, np.float32)
    P, nx = images.shape[0], images.shape[-1]
    if ou == -1:
        ou = nx // 2 - 2
    mask = model_circle(ou, nx)
    numr = numrinit(ir, ou, rs)
    wr = ringwe(numr)
    imgs = np.stack([normalize_mask(im, mask, 0) for im in images])
    params = np.zeros((P, 4))
    cnx = nx // 2 + 1
    mashi = cnx - ou - 2
    sx_sum = sy_sum = 0.0
    history = []
    tavg = None
    # Sphire ali2d_base: for N_step in range(len(xrng)): for Iter in range(max_iter) -- the shipped
    # driver pins N_step = 0 (test_reffree.py:686); sequences run the whole schedule
    sched = _schedule(xr, yr, ts)
    total = len(sched) * int(maxit)
    for it in range(total):
        xr, yr, ts = sched[it // int(maxit)]
        ave = np.zeros((2, nx, nx), np.float32)
        for i in range(P):
            ave[i % 2] += rot_shift2d(imgs[i], params[i, 0], params[i, 1], params[i, 2], int(params[i, 3]))
        tavg = (ave[0] + ave[1]) / np.float32(P)
        frsc = fsc_mask(ave[0], ave[1], mask)
        if center == -1:
            tavg, _, filt = ref_ali2d(mask, 0, tavg, frsc)
            cs = [float(sx_sum) / P, float(sy_sum) / P]
            tavg = fshift(tavg, -cs[0], -cs[1])
        else:
            tavg, cs, filt = ref_ali2d(mask, center, tavg, frsc)
        history.append(dict(cs=cs, filter=filt, tavg=tavg.copy()))
        if it == total - 1:
            break
        cimage = applyws(frngs(polar2dm(tavg, float(cnx), float(cnx), numr), numr), numr, wr)[None]
        centres = np.zeros((P, 2), np.float32); win = np.zeros((P, 4), np.float32)
        sxi = np.zeros(P); syi = np.zeros(P)
        for i in range(P):
            a, sx, sy, _ = combine_params2(params[i, 0], params[i, 1], params[i, 2], int(params[i, 3]), 0.0, -cs[0], -cs[1], 0)
            _, x, y, _ = inverse_transform2(a, sx, sy)
            x = min(max(x, -mashi), mashi); y = min(max(y, -mashi), mashi)
            sxi[i], syi[i] = x, y
            tx = search_range(nx, ou, x, xr); ty = search_range(nx, ou, y, yr)
            centres[i] = (cnx + x, cnx + y); win[i] = (tx[0], tx[1], ty[0], ty[1])
        out = align_batch(imgs, cimage, numr, centres, win, ts, False, nthreads)
        sx_sum = sy_sum = 0.0
        for i in range(P):
            a, sx, sy, m = combine_params2(0.0, -sxi[i], -syi[i], 0, out[i, 0], out[i, 1], out[i, 2], int(out[i, 3]))
            params[i] = (a, sx, sy, m)
            sx_sum += sx if m == 0 else -sx
            sy_sum += sy
        history[-1]["peak"] = out[:, 5].copy()
    return params, tavg, history


def ref_free_alignment_2d(images, class_of, refs, ir=1, ou=-1, rs=1, xr=0, yr=0, ts=1, maxit=1, filt=None, nthreads=1):
    """Class-bound reference-free alignment, CPU twin of gpu_isac's ref_free_alignment_2D
    (cuda/gpu_aln_noref.cu:559-782): every particle is matched against the average of its own class
    only (sbj_cid_list, :566-571), the averages are rebuilt from the transformed particles after every
    pass (:771-775) and optionally low-passed with the tangent filter (:777-814).  The per-particle step
    has EMAN2 ormq semantics (ali2d_single_iter: no ring normalisation, clamped shifts), as in ali2d_base
    above.  filt = (cutoff, falloff) or None.  Returns (params [P][4], refs [R][nx][nx], peaks)."""
    images = np.array(images, np.float32)
    class_of = np.asarray(class_of, np.int64)
    refs = np.array(refs, np.float32)
    P, nx = images.shape[0], images.shape[-1]
    R = refs.shape[0]
    if ou == -1:
        ou = nx // 2 - 2
    mask = model_circle(ou, nx)
    numr = numrinit(ir, ou, rs)
    wr = ringwe(numr)
    imgs = np.stack([normalize_mask(im, mask, 0) for im in images])
    params = np.zeros((P, 4))
    cnx = nx // 2 + 1
    mashi = cnx - ou - 2
    peaks = np.zeros(P, np.float32)
    for it in range(int(maxit)):
        cref = np.stack([applyws(frngs(polar2dm(refs[r], float(cnx), float(cnx), numr), numr), numr, wr) for r in range(R)])
        centres = np.zeros((P, 2), np.float32); win = np.zeros((P, 4), np.float32)
        sxi = np.zeros(P); syi = np.zeros(P)
        for i in range(P):
            _, x, y, _ = inverse_transform2(params[i, 0], params[i, 1], params[i, 2])
            x = min(max(x, -mashi), mashi); y = min(max(y, -mashi), mashi)
            sxi[i], syi[i] = x, y
            tx = search_range(nx, ou, x, xr); ty = search_range(nx, ou, y, yr)
            centres[i] = (cnx + x, cnx + y); win[i] = (tx[0], tx[1], ty[0], ty[1])
        out = np.zeros((P, 8), np.float32)
        for r in range(R):
            idx = np.nonzero(class_of == r)[0]
            if idx.size:
                out[idx] = align_batch(imgs[idx], cref[r:r + 1], numr, centres[idx], win[idx], ts, False, nthreads)
        sums = np.zeros((R, nx, nx), np.float32)
        counts = np.zeros(R)
        for i in range(P):
            a, sx, sy, m = combine_params2(0.0, -sxi[i], -syi[i], 0, out[i, 0], out[i, 1], out[i, 2], int(out[i, 3]))
            params[i] = (a, sx, sy, m)
            sums[class_of[i]] += rot_shift2d(imgs[i], a, sx, sy, int(m))
            counts[class_of[i]] += 1
        peaks = out[:, 5].copy()
        for r in range(R):
            if counts[r] > 0:
                refs[r] = sums[r] / np.float32(counts[r])
            if filt is not None:                                 # the reference filters the whole reference batch
                refs[r] = filt_tanl(refs[r], filt[0], filt[1])
    return params, refs, peaks
