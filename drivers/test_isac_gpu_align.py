#!/usr/bin/env python
"""Class-bound reference-free 2D alignment on B200 -- a driver for gpu_isac's ref_free_alignment_2D
(cuda/gpu_aln_noref.h:94-109; the reference ships the library entry points but no driver that calls them):
every particle is aligned to the average of its own class only, the class averages are rebuilt on the
device after every pass and optionally low-passed with the tangent filter.

    python drivers/test_isac_gpu_align.py stack classes refstack outdir --ou=36 --xr=3 --yr=3 --ts=1 --maxit=4 \
        --fl=0.2 --aa=0.2

stack / refstack: .npy or MRC float32 stacks; classes: text file, one class index per particle (the stack holds
runs of particles of the same class, as gpu_isac lays it out).  Outputs (rank 0): aqm%03d.mrcs per pass,
params.txt rows 'idx angle sx sy mirror class' (src/utils_ralib.py:31-32), logfile.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _common import add_alignment_flags, first_of, init_distributed, pick_device, Log  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("stack"); ap.add_argument("classes"); ap.add_argument("refstack"); ap.add_argument("outdir")
    add_alignment_flags(ap, reffree=True)
    ap.add_argument("--fl", type=float, default=0.0, help="tangent low-pass cut-off of the class averages (0: no filter)")
    ap.add_argument("--aa", type=float, default=0.2, help="tangent low-pass fall-off")
    args = ap.parse_args(argv)
    from cryo_ralib_b200 import stackio, alignment as al
    from cryo_ralib_b200.lib import load_library
    from cryo_ralib_b200.mref import ref_free_alignment_2d
    if args.gpu_info:
        load_library().print_gpu_info(0)
        return 0
    comm, rank, world, local = init_distributed(args)
    if rank == 0:
        if os.path.exists(args.outdir):
            raise SystemExit("Output directory exists, please change the name and restart the program")
        os.makedirs(args.outdir)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    log = Log(args.outdir, rank)
    P, nx = stackio.stack_shape(args.stack)
    s, e = al.mpi_start_end(P, world, rank)
    images = stackio.read_stack(args.stack, s, e)          # this rank's share only (memory-mapped read)
    refs = stackio.read_stack(args.refstack)
    cls = np.loadtxt(args.classes, dtype=np.int64).reshape(-1)
    if cls.shape[0] != P:
        raise SystemExit("classes file has %d entries for %d particles" % (cls.shape[0], P))
    if cls.min() < 0 or cls.max() >= refs.shape[0]:
        raise SystemExit("class index outside the reference stack")
    xr, ts = first_of(args.xr), first_of(args.ts)
    yr = first_of(args.yr) if first_of(args.yr) != -1 else xr
    ou = args.ou if args.ou != -1 else nx // 2 - 2
    maxit = args.maxit if args.maxit > 0 else 4
    if ou + max(xr, yr) > (nx - 1) // 2:
        raise SystemExit("Shift or radius is too large - particle crosses image boundary")
    filt = (args.fl, args.aa) if args.fl > 0 else None
    log.add("ref_free_alignment_2D: %d particles %dx%d in %d classes, ou=%d xr=%g yr=%g ts=%g maxit=%d filter=%s, %d GPU(s)"
            % (P, nx, nx, refs.shape[0], ou, xr, yr, ts, maxit, filt, world))
    t0 = [time.time()]

    def on_iteration(it, params, info):
        if rank == 0:
            dt = time.time() - t0[0]; t0[0] = time.time()
            log.add("Pass #%4d   %.3f s   mean peak = %15.8e" % (it + 1, dt, float(np.mean(info["peak"]))))

    params, new_refs, hist = ref_free_alignment_2d(images, cls[s:e], refs, ir=args.ir, ou=ou, rs=args.rs, xr=xr, yr=yr,
                                                   ts=ts, maxit=maxit, filt=filt, comm=comm, global_offset=s,
                                                   device=pick_device(args, local), on_iteration=on_iteration)
    if world > 1:
        import torch
        import torch.distributed as dist
        full = np.zeros((P, 4)); full[s:e] = params
        t = torch.from_numpy(full).cuda(); dist.all_reduce(t); params = t.cpu().numpy()
    if rank == 0:
        stackio.write_stack(os.path.join(args.outdir, "aqm%03d.mrcs" % maxit), new_refs)
        stackio.write_params(os.path.join(args.outdir, "params.txt"), params, cls)
        log.add("Finished ref_free_alignment_2D")
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
