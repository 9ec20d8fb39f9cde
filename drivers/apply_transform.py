#!/usr/bin/env python
"""Apply 2D alignment parameters to a stack on the GPU -- the downstream step the reference does with
sxtransform2d (notebook/00 cell 4) and its rot_shift_2d_cupy utility (notebook/03 cell 3): every particle
is passed through rot_shift2D(img, alpha, sx, sy, mirror) (EMAN2 rot_scale_trans2D_background semantics,
the same kernel that builds the class sums) and written out; with --averages the per-class averages too.

    python drivers/apply_transform.py stack params.txt out_stack [--averages avg.mrcs] [--no-mask-mean]

params.txt: rows 'idx angle sx sy mirror class' (params.txt of the mref / isac drivers) or
'angle sx sy mirror' (initial2Dparams.txt of the reference-free driver).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _common import ROOT  # noqa: E402,F401


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("stack"); ap.add_argument("params"); ap.add_argument("out_stack")
    ap.add_argument("--averages", default="", help="also write the per-class averages of the transformed particles")
    ap.add_argument("--no-mask-mean", action="store_true", help="do not subtract the in-mask mean first (normalize.mask no_sigma=0)")
    ap.add_argument("--ou", type=int, default=-1, help="mask radius for the mean subtraction (default nx/2-2)")
    ap.add_argument("--gpu", type=int, default=0)
    args = ap.parse_args(argv)
    from cryo_ralib_b200 import Engine, stackio
    images = stackio.read_stack(args.stack)
    params, cls = stackio.read_params(args.params)
    P, nx = images.shape[0], images.shape[-1]
    if params.shape[0] != P:
        raise SystemExit("parameter file has %d rows for %d particles" % (params.shape[0], P))
    ou = args.ou if args.ou != -1 else nx // 2 - 2
    R = int(cls.max()) + 1 if (cls is not None and args.averages) else 1
    e = Engine(nx, ou, 1, max_particles=P, max_refs=R, device=args.gpu)      # no alignment here: the search range is unused
    e.upload_particles(images, subtract_mask_mean=not args.no_mask_mean)
    out = e.transform(0, P, params)
    stackio.write_stack(args.out_stack, out)
    print("wrote", args.out_stack, out.shape)
    if args.averages:
        iref = cls.astype(np.int32) if cls is not None else np.zeros(P, np.int32)
        e.zero_sums()
        e.accumulate(0, P, params, iref, 0)
        sums, counts = e.get_sums()
        avg = (sums[:, 0] + sums[:, 1]) / np.maximum(counts, 1)[:, None, None]
        stackio.write_stack(args.averages, avg.astype(np.float32))
        print("wrote", args.averages, avg.shape, "members per class:", counts.astype(int).tolist())
    e.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
