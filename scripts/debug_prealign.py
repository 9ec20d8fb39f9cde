"""Debug: pre_align_run iteration whose angle differs from the oracle's ormq."""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, ".")
from cryo_ralib_b200 import synth
from cryo_ralib_b200.lib import load_library, AlignConfig, AlignParam
from oracle import oracle as o
o.build()
allp, truth = synth.make_particles(64 + 60, 90, 16, max_shift=3, seed=7)
images = np.ascontiguousarray(allp[:64])
P, nx, ou, xr = 48, 90, 36, 3
mask = o.model_circle(ou, nx); numr = o.numrinit(1, ou, 1); wr = o.ringwe(numr)
imgs = np.stack([o.normalize_mask(im, mask, 0) for im in images[:P]])
L = load_library()
cfg = AlignConfig(P, 1, nx, ou, 256, 1.0, xr, xr)
ptr = L.pre_align_init(P, C.byref(cfg), 0)
par = C.cast(ptr, C.POINTER(AlignParam))
fp = C.POINTER(C.c_float)
def fetch(arrs, which):
    keep = [np.ascontiguousarray(a, np.float32) for a in arrs]
    L.pre_align_fetch((fp * len(keep))(*[d.ctypes.data_as(fp) for d in keep]), len(keep), which)
    return keep
fetch(imgs, b"sbj_batch")
L.reset_shifts(float(xr), 1.0)
cnx = nx // 2 + 1; mashi = cnx - ou - 2
sh = np.zeros((P, 2)); tavg = imgs.mean(axis=0)
for it in range(2):
    fetch([tavg], b"ref_batch")
    cref = o.applyws(o.frngs(o.polar2dm(tavg.astype(np.float32), float(cnx), float(cnx), numr), numr), numr, wr)[None]
    sh = np.clip(sh, -mashi, mashi)
    centres = (cnx + sh).astype(np.float32)
    win = np.zeros((P, 4), np.float32)
    for i in range(P):
        win[i, 0:2] = o.search_range(nx, ou, sh[i, 0], xr); win[i, 2:4] = o.search_range(nx, ou, sh[i, 1], xr)
    want = o.align_batch(imgs, cref, numr, centres, win, 1.0, False, nthreads=8)
    L.pre_align_run(0, P)
    for i in range(P):
        w = want[i]
        d = abs((par[i].angle - w[0] + 180) % 360 - 180)
        if d > 0.71:
            ix, iy = -(sh[i, 0] - par[i].shift_x), -(sh[i, 1] - par[i].shift_y)
            print("it", it, "particle", i, "engine ang %.3f mirror %d shift (%g,%g) | oracle" % (par[i].angle, par[i].mirror, par[i].shift_x, par[i].shift_y), w)
            cx, cy = centres[i, 0] - (sh[i, 0] - par[i].shift_x) * 0 , centres[i, 1]
            # polar centre of the engine's winner: centre + (ix, iy) with r.sx = -ix = sh - new
            ex, ey = centres[i, 0] - (sh[i, 0] - par[i].shift_x), centres[i, 1] - (sh[i, 1] - par[i].shift_y)
            ox, oy = centres[i, 0] + (-w[6]), centres[i, 1] + (-w[7])
            print("   engine centre", ex, ey, "oracle centre", ox, oy)
            for (px, py, tag) in ((ex, ey, "E"), (ox, oy, "O")):
                cur = o.crosrng_ms(cref[0], o.frngs(o.polar2dm(imgs[i], float(px), float(py), numr), numr), numr)
                print("   ", tag, "qn %.6f tot %.3f qm %.6f tmt %.3f; top q lags" % (cur["qn"], cur["tot"], cur["qm"], cur["tmt"]), np.argsort(cur["q"])[-3:][::-1], np.sort(cur["q"])[-3:][::-1], "top t", np.argsort(cur["t"])[-3:][::-1], np.sort(cur["t"])[-3:][::-1])
    sh = np.array([[par[i].shift_x, par[i].shift_y] for i in range(P)], np.float64)
    acc = np.zeros((nx, nx), np.float32)
    for i in range(P):
        a = np.deg2rad(par[i].angle)
        sx = -par[i].shift_x * np.cos(a) - par[i].shift_y * np.sin(a); sy = par[i].shift_x * np.sin(a) - par[i].shift_y * np.cos(a)
        acc += o.rot_shift2d(imgs[i], par[i].angle, sx, sy, int(par[i].mirror))
    tavg = acc / np.float32(P)
