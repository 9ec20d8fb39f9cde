"""Wall time of whole user-level iterations (mref_ali2d / ali2d_base: alignment + class sums + reference update on
the host) beside the device time of the alignment kernels.  usage: iter_profile.py [mref|reffree] [P] [R]"""
import cProfile
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
from cryo_ralib_b200 import synth  # noqa: E402
from cryo_ralib_b200.mref import mref_ali2d, ali2d_base  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "mref"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
R = int(sys.argv[3]) if len(sys.argv) > 3 else 50
img_d, _ = synth.make_particles(P, 90, 64, max_shift=3, seed=2025, device="cuda:0")
refs = synth.initial_references(img_d, R, seed=99).cpu().numpy()
images = img_d.cpu().numpy()
del img_d
torch.cuda.empty_cache()
t0 = [time.time()]


def tick(it, *a):
    info = a[-1]
    now = time.time()
    st = info.get("stats", {})
    print("iteration %d: wall %.3f s, alignment kernels %.3f s" % (it + 1, now - t0[0], 1e-3 * st.get("ms_total", 0.0)), flush=True)
    t0[0] = now


pr = cProfile.Profile()
pr.enable()
if mode == "mref":
    mref_ali2d(images, refs, ou=36, xr=3, yr=3, ts=1, maxit=3, on_iteration=tick)
else:
    ali2d_base(images, ou=36, xr=3, yr=3, ts=1, maxit=4, on_iteration=tick)
pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(22)
