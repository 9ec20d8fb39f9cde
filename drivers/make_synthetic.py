#!/usr/bin/env python
"""Write a synthetic particle stack + initial references for the drivers."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cryo_ralib_b200 import synth, stackio  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("out_stack"); ap.add_argument("out_refs")
ap.add_argument("--n", type=int, default=10000); ap.add_argument("--nx", type=int, default=90)
ap.add_argument("--refs", type=int, default=10); ap.add_argument("--views", type=int, default=64)
ap.add_argument("--xr", type=int, default=3); ap.add_argument("--seed", type=int, default=2024)
a = ap.parse_args()
images, _ = synth.make_particles(a.n, a.nx, max(a.views, a.refs), max_shift=a.xr, seed=a.seed)
stackio.write_stack(a.out_stack, images)
stackio.write_stack(a.out_refs, synth.initial_references(images, a.refs, seed=99))
print("wrote", a.out_stack, images.shape, "and", a.out_refs)
