#!/usr/bin/env python
"""2D multi-reference alignment on B200 -- the counterpart of the reference's
test_mref_gpu_align.py / test_mref.py entry point, same positional arguments and flags:

    python drivers/test_mref_gpu_align.py stack refstack outdir --ou=36 --xr=3 --yr=3 --ts=1 --maxit=6
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 drivers/test_mref_gpu_align.py stack refstack outdir ...

Stacks are .npy or MRC float32 stacks.  Outputs (rank 0): aqm%03d.mrcs per iteration
(test_mref.py:240, :285), drm%03d%04d.txt = the even/odd FSC of every class of that iteration
(test_mref.py:254), params.txt with 'idx angle sx sy mirror class' rows, logfile.
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
    ap.add_argument("stack"); ap.add_argument("refstack"); ap.add_argument("outdir"); ap.add_argument("maskfile", nargs="?")
    add_alignment_flags(ap)
    args = ap.parse_args(argv)
    from cryo_ralib_b200 import stackio, alignment as al
    from cryo_ralib_b200.lib import load_library
    from cryo_ralib_b200.mref import mref_ali2d
    if args.gpu_info:
        load_library().print_gpu_info(0)
        return 0
    if args.maskfile:
        raise SystemExit("user masks are not on the accelerated path; the default model_circle(ou) mask is used")
    if args.function != "ref_ali2d":
        raise SystemExit("only --function=ref_ali2d is supported")
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
    xr = first_of(args.xr); yr = first_of(args.yr) if first_of(args.yr) >= 0 else xr; ts = first_of(args.ts)
    ou = args.ou if args.ou != -1 else nx // 2 - 2
    maxit = args.maxit if args.maxit > 0 else 10
    if ou + max(xr, yr) > (nx - 1) // 2:
        log.add("note: ou + range exceeds (nx-1)//2; windows are clipped per particle by search_range (test_mref.py:195-198)")
    log.add("mref_ali2d_gpu: %d particles %dx%d, %d references, ir=%d ou=%d rs=%d xr=%g yr=%g ts=%g center=%d maxit=%d, %d GPU(s)"
            % (P, nx, nx, refs.shape[0], args.ir, ou, args.rs, xr, yr, ts, args.center, maxit, world))

    t0 = [time.time()]

    def on_iteration(it, params, assign, new_refs, info):
        if rank == 0:
            stackio.write_stack(os.path.join(args.outdir, "aqm%03d.mrcs" % it), new_refs)
            for j, frsc in sorted(info.get("class_fsc", {}).items()):
                stackio.write_fsc(os.path.join(args.outdir, "drm%03d%04d.txt" % (it, j)), frsc)
            dt = time.time() - t0[0]; t0[0] = time.time()
            st = info.get("stats", {})
            log.add("ITERATION #%3d   %.3f s   %.3e alignments/s (whole iteration, wall clock)   filter cut-off %.3f fall-off %.3f"
                    % (it + 1, dt, (st.get("alignments", 0) * world) / max(dt, 1e-9) if st else 0.0,
                       info["filter"][0], info["filter"][1]))
            for j, c in enumerate(info["counts"]):
                log.add("   group #%3d   number of particles = %7d" % (j, int(c)))

    params, assign, new_refs, hist = mref_ali2d(images, refs, ir=args.ir, ou=ou, rs=args.rs, xr=xr, yr=yr, ts=ts,
                                                center=args.center, maxit=maxit, rand_seed=args.rand_seed, comm=comm,
                                                total_particles=P, global_offset=s, device=pick_device(args, local),
                                                on_iteration=on_iteration)
    # gather parameter rows on rank 0 (recv_attr_dict, test_mref.py:304-313)
    if world > 1:
        import torch
        import torch.distributed as dist
        full = np.zeros((P, 5)); full[s:e, :4] = params; full[s:e, 4] = assign
        t = torch.from_numpy(full).cuda(); dist.all_reduce(t); full = t.cpu().numpy()
        params, assign = full[:, :4], full[:, 4].astype(int)
    if rank == 0:
        stackio.write_params(os.path.join(args.outdir, "params.txt"), params, assign)
        log.add("mref_ali2d_gpu finished")
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
