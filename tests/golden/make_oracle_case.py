#!/usr/bin/env python
"""Writes tests/golden/oracle_mref_case.npz: a small end-to-end multi-reference case (inputs + the outputs of the
oracle in this repository).  It guards the oracle against drift (tests/test_oracle.py reproduces it on the CPU) and
gives the GPU tests a fixed vector that does not depend on the oracle library being rebuilt.  EMAN2 itself cannot be
run here, so these are NOT EMAN2-generated vectors: parity at the multiref_polar_ali_2d boundary stays unpinned.

    python tests/golden/make_oracle_case.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cryo_ralib_b200 import synth  # noqa: E402
from oracle import oracle as o  # noqa: E402


def main():
    o.build()
    P, R, nx, ou, xr = 16, 4, 48, 16, 2
    allp, _ = synth.make_particles(P + 8 * R, nx, 8, max_shift=xr, seed=31)
    images = np.ascontiguousarray(allp[:P])
    refs = synth.initial_references(allp[P:], R, per_ref=8, seed=7)
    mask = o.model_circle(ou, nx)
    numr = o.numrinit(1, ou, 1)
    _, cref = o.prepare_refs(refs, mask, numr)
    params0 = np.zeros((P, 4))
    params1, assign, peak, sums, counts = o.mref_iteration(images.copy(), mask, cref, numr, xr, xr, 1, ou, params0, 0, True, 1)
    imgs = np.stack([o.normalize_mask(im, mask, 0) for im in images])
    cnx = nx // 2 + 1
    centres = np.full((P, 2), float(cnx), np.float32)
    win = np.full((P, 4), float(xr), np.float32)
    rows = o.align_batch(imgs, cref, numr, centres, win, 1.0, True, 1)
    spec0 = o.frngs(o.normalize_ring(o.polar2dm(imgs[0], cnx + 1.0, cnx - 2.0, numr), numr), numr)
    import random
    new_refs, info = o.update_refs(sums, counts, imgs, mask, 1, random.Random(1000))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_mref_case.npz"),
                        images=images, refs=refs, nx=nx, ou=ou, xr=xr, numr=numr, cref=cref, align_rows=rows,
                        params1=params1, assign=assign, peak=peak, sums=sums, counts=counts, spec0=spec0,
                        new_refs=new_refs, filter=np.array(info["filter"]))
    print("written; assignments", assign, "filter", info["filter"])


if __name__ == "__main__":
    main()
