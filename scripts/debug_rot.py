"""Debug: per-particle rot_shift2D (engine vs oracle) on config-1 iteration-1 parameters."""
import sys
import numpy as np
sys.path.insert(0, ".")
from cryo_ralib_b200 import Engine, synth, alignment as al
from oracle import oracle as o
o.build()
P, R, nx, ou, xr = 2000, 10, 90, 36, 3
images, _ = synth.make_particles(P, nx, 64, max_shift=xr, seed=2025)
refs = synth.initial_references(images, R, seed=99)
mask = o.model_circle(ou, nx); numr = o.numrinit(1, ou, 1)
imgs = np.stack([o.normalize_mask(im, mask, 0) for im in images])
_, cref = o.prepare_refs(refs, mask, numr)
p_new, a_o, pk_o, s_o, c_o = o.mref_iteration(images.copy(), mask, cref, numr, xr, xr, 1, ou, np.zeros((P, 4)), 0, True, o.max_threads())
e = Engine(nx, ou, xr, max_particles=P, max_refs=R)
e.upload_particles(images); e.set_refs(refs)
got = e.transform(0, P, p_new)
worst = []
for i in range(P):
    w = o.rot_shift2d(imgs[i], p_new[i, 0], p_new[i, 1], p_new[i, 2], int(p_new[i, 3]))
    d = np.abs(got[i] - w)
    worst.append((d.max() / np.abs(w).max(), i, int(d.argmax()) // nx, int(d.argmax()) % nx, int((d > 1e-4 * np.abs(w).max()).sum())))
worst.sort(reverse=True)
for x in worst[:12]:
    print("rel %.2e particle %d at (y %d, x %d), pixels over 1e-4: %d, params %s" % (x[0], x[1], x[2], x[3], x[4], p_new[x[1]]))
