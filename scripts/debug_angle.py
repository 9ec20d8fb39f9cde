"""Debug: particles of config 1 whose angle differs from the oracle although reference / mirror / shift agree."""
import sys
import numpy as np
sys.path.insert(0, ".")
from cryo_ralib_b200 import Engine, synth, alignment as al
from oracle import oracle as o
o.build()
P, R, nx, ou, xr = 10000, 10, 90, 36, 3
images, _ = synth.make_particles(P, nx, 64, max_shift=xr, seed=2025)
refs = synth.initial_references(images, R, seed=99)
mask = o.model_circle(ou, nx); numr = o.numrinit(1, ou, 1)
ids = [1983, 3080, 5]
imgs = np.stack([o.normalize_mask(images[i], mask, 0) for i in ids])
_, cref = o.prepare_refs(refs, mask, numr)
e = Engine(nx, ou, xr, max_particles=len(ids), max_refs=R)
e.upload_particles(images[ids]); e.set_refs(refs)
search, sxi, syi, _ = al.mref_search_request(np.zeros((len(ids), 4)), nx, ou, xr, xr)
res = e.align(0, len(ids), search)
want = o.align_batch(imgs, cref, numr, np.stack([search["cx"], search["cy"]], 1), np.stack([search["xl"], search["xr"], search["yl"], search["yr"]], 1), 1.0, True, 4)
for i in range(len(ids)):
    r, w = res[i], want[i]
    print("particle", ids[i], "engine", r, "oracle", w)
    cx, cy = float(search["cx"][i]) - float(r["sx"]), float(search["cy"][i]) - float(r["sy"])
    c = o.frngs(o.normalize_ring(o.polar2dm(imgs[i], cx, cy, numr), numr), numr)
    cur = o.crosrng_ms(cref[int(r["iref"])], c, numr)
    curve = cur["t"] if int(r["mirror"]) else cur["q"]
    q, t = e.ccf_curves(i, cx, cy, int(r["iref"]))
    dcurve = t if int(r["mirror"]) else q
    lag_e = float(r["ang"]) / 360 * 256; lag_o = float(w[0]) / 360 * 256
    print("  lags engine %.3f oracle %.3f; oracle curve argmax %d max %.6f qn %.6f qm %.6f" % (lag_e, lag_o, int(np.argmax(curve)), curve.max(), cur["qn"], cur["qm"]))
    for L in (int(round(lag_e)) % 256, int(round(lag_o)) % 256):
        print("   around lag", L, "oracle", [float("%.6f" % curve[(L + d) % 256]) for d in range(-3, 4)], "device", [float("%.6f" % dcurve[(L + d) % 256]) for d in range(-3, 4)])
    top = np.argsort(curve)[-4:][::-1]
    print("   oracle top lags", top, curve[top], " device top", np.argsort(dcurve)[-4:][::-1], np.sort(dcurve)[-4:][::-1])
