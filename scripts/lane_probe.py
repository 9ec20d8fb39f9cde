"""Wall time of cra_align over several row batches with per-kernel timing off (two pipeline lanes, cra_api.cu) --
run once with CRA_LANES=1 (one lane) and once without.  usage: lane_probe.py [P] [R] [rows per batch]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from cryo_ralib_b200 import Engine, synth, alignment as al  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
R = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ROWS = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # rows per batch (0: the engine's default, 2 GB of spectra)
nx, ou, xr = 90, 36, 3
images, _ = synth.make_particles(P, nx, 64, seed=2025)
refs = synth.initial_references(images, R, per_ref=max(1, min(200, P // R)), seed=99)
e = Engine(nx, ou, xr, ts=1.0, max_particles=P, max_refs=R, row_batch=ROWS)
e.upload_particles(images); e.set_refs(refs)
search, sxi, syi, _ = al.mref_search_request(np.zeros((P, 4)), nx, ou, xr, xr)
res0 = None
for it in range(6):
    t = time.perf_counter(); res = e.align(0, P, search); dt = time.perf_counter() - t
    if res0 is None:
        res0 = res
    same = all(np.array_equal(res[k], res0[k]) for k in res.dtype.names)
    if it >= 2:
        print("align %d particles: %.2f ms, %.4e alignments/s, identical to first pass: %s" % (P, 1e3 * dt, e.stats()["alignments"] / dt, same))
