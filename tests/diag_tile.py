"""Row-spectrum error of every shift row against the oracle, per particle, with the windowed image tile of the grouped row
kernel (default) and with CRA_GRP_TILE=0 (whole image: the general kernel at this box size).  usage: python tests/diag_tile.py"""
import sys, os, numpy as np
sys.path.insert(0, ".")
from oracle import oracle
from cryo_ralib_b200 import Engine, synth
from cryo_ralib_b200.lib import SEARCH_DTYPE
nx, ou, ts, xr = 256, 60, 1.0, 3.0     # the geometry of tests/test_gpu_largebox.py::test_windowed_tile_row_spectra_match_oracle
P = 6
allp, _ = synth.make_particles(P + 8, nx, 16, max_shift=2, seed=21)
images = np.ascontiguousarray(allp[:P]); refs = synth.initial_references(allp[P:], 2, per_ref=4, seed=5)
mask = oracle.model_circle(ou, nx); numr = oracle.numrinit(1, ou, 1)
imgs = np.stack([oracle.normalize_mask(im, mask, 0) for im in images])
e = Engine(nx, ou, xr, ts=ts, max_particles=P, max_refs=2, normalize_ring=True)
e.upload_particles(images, subtract_mask_mean=True); e.set_refs(refs)
c = nx // 2 + 1
lo, hi = ou + 2 + xr, nx - 1 - ou - xr
search = np.zeros(P, SEARCH_DTYPE)
search["cx"] = [c, lo, hi, c + 0.5, c - 3.37, hi - 0.25]
search["cy"] = [c, hi, lo, c - 0.5, c + 2.81, lo + 0.75]
search["xl"] = xr; search["xr"] = xr; search["yl"] = xr; search["yr"] = xr
e.align(0, P, search)
row = 0
for p in range(P):
    s = search[p]; errs = []; scales = []
    for iy in range(-3, 4):
        for ix in range(-3, 4):
            got, kern = e.batch_row_spectrum(row)
            cc = oracle.normalize_ring(oracle.polar2dm(imgs[p], float(s["cx"]) + ix, float(s["cy"]) + iy, numr), numr)
            want = oracle.frngs(cc, numr)
            errs.append(np.abs(got - want).max() / np.abs(want).max()); scales.append(np.abs(want).max())
            row += 1
    print("particle", p, "kernel", kern, "max rel err %.2e median %.2e scale %.1f..%.1f" % (max(errs), np.median(errs), min(scales), max(scales)))
