#!/usr/bin/env python
"""Reference-free 2D alignment on B200 -- the counterpart of the reference's
test_reffree_gpu_align.py / test_reffree.py (ali2d_base semantics), same arguments and flags:

    python drivers/test_reffree_gpu_align.py stack outdir --ou=36 --xr=3 --yr=3 --ts=1 --maxit=6

Outputs (rank 0): aqc.mrcs / aqf.mrcs (raw and filtered average per iteration), aqfinal.mrc,
initial2Dparams.txt (test_reffree.py:333, :370, :490, :505), logfile.
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
    ap.add_argument("stack"); ap.add_argument("outdir"); ap.add_argument("maskfile", nargs="?")
    add_alignment_flags(ap, reffree=True)
    ap.add_argument("--first_step_only", action="store_true",
                    help="use only entry 0 of the --xr/--yr/--ts lists, as the reference driver does")
    args = ap.parse_args(argv)
    from cryo_ralib_b200 import stackio, alignment as al
    from cryo_ralib_b200.lib import load_library
    from cryo_ralib_b200.mref import ali2d_base
    if args.gpu_info:
        load_library().print_gpu_info(0)
        return 0
    if args.maskfile:
        raise SystemExit("user masks are not on the accelerated path; the default model_circle(ou) mask is used")
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
    from cryo_ralib_b200.mref import search_schedule
    # the whole "--xr 4 2 1 --ts 2 1 0.5" schedule (Sphire ali2d_base); --first_step_only reproduces the
    # reference driver, which pins N_step = 0 (test_reffree.py:310, :686)
    sched = search_schedule(args.xr, args.yr, args.ts)
    if args.first_step_only:
        sched = sched[:1]
    xr = [x for x, _, _ in sched]; yr = [y for _, y, _ in sched]; ts = [t for _, _, t in sched]
    ou = args.ou if args.ou != -1 else nx // 2 - 2
    maxit = args.maxit if args.maxit > 0 else 10
    if ou + max(max(xr), max(yr)) > (nx - 1) // 2:
        raise SystemExit("Shift or radius is too large - particle crosses image boundary")   # test_reffree.py:603
    log.add("ali2d_base: %d particles %dx%d, ir=%d ou=%d rs=%d xr=%s yr=%s ts=%s center=%d maxit=%d per step, %d GPU(s)"
            % (P, nx, nx, args.ir, ou, args.rs, xr, yr, ts, args.center, maxit, world))
    raw, filt = [], []
    t0 = [time.time()]

    def on_iteration(it, params, tavg, info):
        if rank == 0:
            filt.append(tavg.copy())
            dt = time.time() - t0[0]; t0[0] = time.time()
            log.add("Iteration #%4d   %.3f s   Criterion = %15.8e   Average center x = %10.3f y = %10.3f"
                    % (it + 1, dt, info["criterion"], info["cs"][0], info["cs"][1]))

    params, tavg, hist = ali2d_base(images, ir=args.ir, ou=ou, rs=args.rs, xr=xr, yr=yr, ts=ts, center=args.center,
                                    maxit=maxit, comm=comm, total_particles=P, global_offset=s,
                                    device=pick_device(args, local), on_iteration=on_iteration)
    if world > 1:
        import torch
        import torch.distributed as dist
        full = np.zeros((P, 4)); full[s:e] = params
        t = torch.from_numpy(full).cuda(); dist.all_reduce(t); params = t.cpu().numpy()
    if rank == 0:
        if filt:
            stackio.write_stack(os.path.join(args.outdir, "aqf.mrcs"), np.stack(filt))
        stackio.write_stack(os.path.join(args.outdir, "aqfinal.mrc"), tavg[None])
        with open(os.path.join(args.outdir, "initial2Dparams.txt"), "w") as f:
            for p in params:
                f.write("%14.6f %14.6f %14.6f %d\n" % (p[0], p[1], p[2], int(p[3])))
        log.add("Finished ali2d_base")
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
