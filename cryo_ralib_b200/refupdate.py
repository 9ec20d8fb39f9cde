"""Per-iteration reference update: the rank-0 section of the reference's loop
(test_mref.py:238-286; reference-free twin test_reffree.py:314-372), which there is Python
calling EMAN2 on R small images.  Same here: host-side numpy over the R class sums the device
produced (R * nx^2 floats per iteration -- not on the per-particle hot path).

  fsc(even, odd)        EMData::calc_fourier_shell_correlation   (test_mref.py:254)
  ref_ali2d             fit_tanh -> filt_tanl -> center_2D       (test_mref.py:273-276)
  normalize.mask        (x - mean_mask) / sigma_mask             (test_mref.py:284)
"""
import math
import random

import numpy as np


def _round_half_away(x):
    return np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5)).astype(np.int64)


class FourierGeometry(object):
    """Frequency grids of the half-complex FFT of an [ny][nx] image, cached per size."""
    _cache = {}

    def __new__(cls, ny, nx):
        key = (ny, nx)
        if key not in cls._cache:
            g = object.__new__(cls)
            g.ny, g.nx = ny, nx
            ky = np.where(np.arange(ny) > ny // 2, np.arange(ny) - ny, np.arange(ny))[:, None]
            kx = np.arange(nx // 2 + 1)[None, :]
            nx2, ny2 = nx // 2, ny // 2
            inc = int(max(nx2, ny2) + 0.5)
            argx = 0.5 * np.sqrt((ky * ky / float(ny2 * ny2) + kx * kx / float(nx2 * nx2)).astype(np.float32))
            g.shell = _round_half_away(inc * 2 * argx.astype(np.float64))
            g.inc = inc
            g.use = ((kx > 0) | (ky >= 0)) & (g.shell <= inc)         # skip Friedel mates on the kx=0 column
            g.fy = ky / float(ny)
            g.fx = kx / float(nx)
            g.rad = np.sqrt(g.fx * g.fx + g.fy * g.fy)
            cls._cache[key] = g
        return cls._cache[key]


def fsc(img1, img2):
    """[freq, fsc, n] lists, as sp_statistics.fsc(img1, img2, 1.0) returns them."""
    ny, nx = img1.shape
    g = FourierGeometry(ny, nx)
    F = np.fft.rfft2(img1.astype(np.float64))
    G = np.fft.rfft2(img2.astype(np.float64))
    sh = g.shell[g.use]
    num = np.bincount(sh, (F.real * G.real + F.imag * G.imag)[g.use], g.inc + 1)
    n1 = np.bincount(sh, (F.real ** 2 + F.imag ** 2)[g.use], g.inc + 1)
    n2 = np.bincount(sh, (G.real ** 2 + G.imag ** 2)[g.use], g.inc + 1)
    lr = np.bincount(sh, None, g.inc + 1) * 2.0
    keep = lr > 0
    den = np.sqrt(n1 * n2)
    val = np.where(den > 0, num / np.where(den > 0, den, 1.0), 0.0).astype(np.float32).astype(np.float64)
    idx = np.arange(g.inc + 1)[keep]
    return [list(idx / float(2 * g.inc)), list(val[keep]), list(lr[keep])]


def fsc_mask(img1, img2, mask):
    """sp_statistics.fsc_mask: subtract the in-mask mean, multiply by the mask, then fsc
    (test_reffree.py:708)."""
    m = mask > 0.5
    a = (img1 - np.float32(img1[m].astype(np.float64).mean())) * mask
    b = (img2 - np.float32(img2[m].astype(np.float64).mean())) * mask
    return fsc(a, b)


def amoeba(var, scale, func, ftolerance=1.e-4, xtolerance=1.e-4, itmax=500, data=None):
    """sp_utilities.amoeba (Nelder-Mead maximiser); kept step-for-step because the fitted
    filter parameters depend on its exact path."""
    nvar = len(var)
    nsimplex = nvar + 1
    simplex = [var[:]] + [var[:] for _ in range(nvar)]
    for i in range(nvar):
        simplex[i + 1][i] += scale[i]
    fvalue = [func(s, data=data) for s in simplex]
    iteration = 0
    while True:
        ssworst = ssbest = 0
        for i in range(nsimplex):
            if fvalue[i] > fvalue[ssbest]:
                ssbest = i
            if fvalue[i] < fvalue[ssworst]:
                ssworst = i
        pavg = [sum(simplex[i][j] for i in range(nsimplex) if i != ssworst) / nvar for j in range(nvar)]
        simscale = sum(abs(pavg[i] - simplex[ssworst][i]) / scale[i] for i in range(nvar)) / nvar
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
                    simplex[i] = [0.5 * simplex[ssbest][j] + 0.5 * simplex[i][j] for j in range(nvar)]
                    fvalue[i] = func(simplex[i], data=data)
            pnew = [0.5 * simplex[ssbest][j] + 0.5 * simplex[ssworst][j] for j in range(nvar)]
            fnew = func(pnew, data=data)
        elif fnew >= fvalue[ssbest]:
            pnew2 = [3.0 * pavg[i] - 2.0 * simplex[ssworst][i] for i in range(nvar)]
            fnew2 = func(pnew2, data=data)
            if fnew2 > fnew:
                pnew, fnew = pnew2, fnew2
        simplex[ssworst] = pnew[:]
        fvalue[ssworst] = fnew
        iteration += 1


def fit_tanh(dres, low=0.1, native=True):
    """sp_filter.fit_tanh: (cut-off, fall-off) of the tanh low-pass that best fits 2f/(1+f)."""
    freq = np.array(dres[0], np.float64)
    val = np.array(dres[1], np.float64)
    val = np.where(val == -1.0, -1.0 + 1e-12, val)          # 2f/(1+f) pole; EMAN2 would raise here
    full = 2 * val / (1.0 + val)
    below = np.where(full[1:] < low)[0]
    if below.size:
        val[1 + below[0]:] = 0.0
    full = 2 * val / (1.0 + val)
    fl = -1.0
    for i in range(1, len(freq) - 1):
        if full[i] < 0.5:
            fl = freq[i - 1]
            break
    if fl < 0.0:
        return 0.4, 0.2
    if val[0] < 0.0:
        val[0] = -val[0]
    target = 2 * val / (1.0 + val)

    def cost(args, data=None):
        if args[0] == 0 or args[1] == 0:
            return -float(np.sum(target * 0 + 0.0))
        c = math.pi / 2.0 / args[1] / args[0]
        qt = target - 0.5 * (np.tanh(c * (freq + args[0])) - np.tanh(c * (freq - args[0])))
        return -float(np.sum(qt * qt))

    L = _native() if native else None
    if L is not None:
        # the same simplex in native code (csrc/cra_host.cu: cra_fit_tanh): ~1500 cost evaluations
        f = np.ascontiguousarray(freq, np.float64); t = np.ascontiguousarray(target, np.float64)
        out = np.zeros(4, np.float64)
        if L.cra_fit_tanh(len(f), f.ctypes.data, t.ctypes.data, float(fl), 0.1, 0.05, 0.05, out.ctypes.data) != 0:
            raise RuntimeError(L.cra_last_error().decode())
        return float(out[0]), float(out[1])
    best, _, _ = amoeba([fl, 0.1], [0.05, 0.05], cost)
    return best[0], best[1]


def filt_tanl(img, fl, aa):
    ny, nx = img.shape
    g = FourierGeometry(ny, nx)
    c = math.pi / (2.0 * aa * fl)
    H = 0.5 * (np.tanh(c * (g.rad + fl)) - np.tanh(c * (g.rad - fl)))
    return np.fft.irfft2(np.fft.rfft2(img.astype(np.float64)) * H, s=(ny, nx)).astype(np.float32)


def fshift(img, sx, sy):
    ny, nx = img.shape
    g = FourierGeometry(ny, nx)
    ph = np.exp(-2j * math.pi * (g.fx * sx + g.fy * sy))
    return np.fft.irfft2(np.fft.rfft2(img.astype(np.float64)) * ph, s=(ny, nx)).astype(np.float32)


def phase_cog(img):
    ny, nx = img.shape
    out = []
    for marg, n in ((img.sum(axis=0, dtype=np.float64), nx), (img.sum(axis=1, dtype=np.float64), ny)):
        ang = 2 * math.pi / n * np.arange(n)
        f1 = math.atan2(float(np.dot(np.sin(ang), marg)), float(np.dot(np.cos(ang), marg)))
        if f1 < 0.0:
            f1 += 2 * math.pi
        out.append(f1 / (2 * math.pi / n) + 1.0 - (n // 2 + 1))
    return out


def normalize_mask(img, mask, no_sigma):
    m = mask > 0.5
    v = img[m].astype(np.float64)
    n = v.size
    mean = np.float32(np.float32(v.sum()) / np.float32(n))
    if no_sigma == 0:
        return (img - mean).astype(np.float32)
    sigma = np.float32(math.sqrt(np.float32((np.dot(v, v) - v.sum() ** 2 / n) / (n - 1))))
    return ((img - mean) / sigma).astype(np.float32)


def model_circle(r, nx):
    y, x = np.mgrid[0:nx, 0:nx]
    x2 = np.abs(x - nx // 2).astype(np.float32)
    y2 = np.abs(y - nx // 2).astype(np.float32)
    rr = np.float32(r)
    return ((x2 * x2) / (rr * rr) + (y2 * y2) / (rr * rr) <= 1).astype(np.float32)


def _native():
    """libcryo_ralib.so (cra_fit_tanh); a missing library raises LibraryMissing, there is no silent fallback."""
    from .lib import load_library
    return load_library()


def ref_ali2d(avg, frsc, center, fit=None):
    """sp_user_functions.ref_ali2d.  fit: a (fl, aa) already fitted to this frsc (mref_ali2d hands every class the
    same class-averaged FSC, test_mref.py:258-276, so update_refs fits it once)."""
    fl, aa = fit if fit is not None else fit_tanh(frsc)
    aa = min(aa, 0.2)
    fl = max(min(0.4, fl), 0.12)
    out = filt_tanl(avg, fl, aa)
    cs = [0.0, 0.0]
    if center == 1:
        cs = phase_cog(out)
        out = fshift(out, -cs[0], -cs[1])
    elif center != 0:
        raise NotImplementedError("center methods other than 0/1 (and -1 in the ref-free driver) are off the path")
    return out, cs, (fl, aa)


def update_refs(sums, counts, mask, center=1, reseed=None):
    """New references from the (all-reduced) even/odd class sums.  `reseed(j)` must return the
    replacement image for a class with fewer than 4 members (test_mref.py:244-249)."""
    R = sums.shape[0]
    refs = np.zeros((R,) + sums.shape[2:], np.float32)
    acc = None
    nfsc = 0
    frsc = None
    reseeded = []
    class_fsc = {}
    for j in range(R):
        if counts[j] < 4:
            refs[j] = reseed(j)
            reseeded.append(j)
        else:
            frsc = fsc(sums[j, 0], sums[j, 1])
            class_fsc[j] = [list(frsc[0]), list(frsc[1]), list(frsc[2])]
            refs[j] = (sums[j, 0] + sums[j, 1]) * np.float32(1.0 / float(counts[j]))
            acc = np.array(frsc[1]) if acc is None else acc + np.array(frsc[1])
            nfsc += 1
    if frsc is None:
        raise RuntimeError("every reference vanished (all classes have < 4 members)")
    if acc.sum() != 0:
        frsc[1] = list(acc / float(nfsc))
    info = dict(frsc=[list(frsc[0]), list(frsc[1]), list(frsc[2])], reseeded=reseeded, cs=[], filter=None, class_fsc=class_fsc)
    fit = fit_tanh(frsc)                      # the same FSC curve for every class: one fit
    for j in range(R):
        refs[j], cs, info["filter"] = ref_ali2d(refs[j], frsc, center, fit=fit)
        info["cs"].append(cs)
        refs[j] = normalize_mask(refs[j], mask, 1)
    return refs, info


def make_reseeder(rand_seed, n_global, fetch):
    """Vanished-class rule: a seeded draw over GLOBAL particle indices, so every rank makes the
    same choice (the reference draws from rank 0's share only, test_mref.py:248)."""
    rng = random.Random(rand_seed)
    return lambda j: fetch(rng.randint(0, n_global - 1))
