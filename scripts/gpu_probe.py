"""Quick device probe: FP32 peak micro-benchmark + stage timings on a config-2-shaped slice.
usage: gpu_probe.py [P] [R] [nx] [ou] [xr] [ts]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from cryo_ralib_b200 import Engine, synth, alignment as al  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
R = int(sys.argv[2]) if len(sys.argv) > 2 else 50
nx = int(sys.argv[3]) if len(sys.argv) > 3 else 90
ou = int(sys.argv[4]) if len(sys.argv) > 4 else 36
xr = int(sys.argv[5]) if len(sys.argv) > 5 else 3
ts = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
images, _ = synth.make_particles(P, nx, 64, seed=2025)
refs = synth.initial_references(images, R, per_ref=max(1, min(200, P // R)), seed=99)
e = Engine(nx, ou, xr, ts=ts, max_particles=P, max_refs=R)
print("fp32 peak (ffma, ffma2) TFLOP/s:", e.measure_fp32_peak())
e.upload_particles(images)
e.set_refs(refs)
search, sxi, syi, _ = al.mref_search_request(np.zeros((P, 4)), nx, ou, xr, xr)
e.set_timing(True)
for it in range(3):
    t = time.time()
    res = e.align(0, P, search)
    dt = time.time() - t
    st = e.stats()
    st["wall_s"] = dt
    st["align_per_s"] = st["alignments"] / dt
    print(json.dumps(st))
newp = al.compose_result(sxi, syi, res)
t = time.time(); e.zero_sums(); e.accumulate(0, P, newp, res["iref"], 0); print("accumulate s:", time.time() - t)
print("assign histogram:", np.bincount(res["iref"], minlength=R)[:10])
# second iteration: fractional centres and ragged windows, as every iteration after the first sees them
search2, sxi2, syi2, _ = al.mref_search_request(newp, nx, ou, xr, xr)
for it in range(2):
    res2 = e.align(0, P, search2)
    st = e.stats()
    print("iter2", json.dumps(st))
