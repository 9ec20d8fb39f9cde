"""N-GPU consistency check (torchrun): mref_ali2d and the class-bound ref_free_alignment_2d sharded over the ranks
(NCCL allreduce of the device-resident class sums) against the same loops on one GPU.  Prints one line per check.
usage: torchrun --nproc-per-node N --master-addr 127.0.0.1 scripts/check_multi_gpu.py"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from cryo_ralib_b200 import synth, alignment as al  # noqa: E402
from cryo_ralib_b200.mref import mref_ali2d, ref_free_alignment_2d, TorchComm  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P, R = 4000, 12
allp, _ = synth.make_particles(P + 240, 90, 24, max_shift=3, seed=77)
images, refs = allp[:P], synth.initial_references(allp[P:], R, per_ref=20, seed=3)
s, e = al.mpi_start_end(P, world, rank)
comm = TorchComm()


def gather(x):
    full = np.zeros((P,) + x.shape[1:], np.float64); full[s:e] = x
    t = torch.from_numpy(full).cuda(); dist.all_reduce(t)
    return t.cpu().numpy()


ok = True
# multi-reference alignment, 3 iterations
p_n, a_n, r_n, _ = mref_ali2d(images[s:e], refs, ou=36, xr=3, yr=3, ts=1, maxit=3, comm=comm, total_particles=P,
                              global_offset=s, device=local)
p_n = gather(p_n); a_n = gather(a_n[:, None].astype(np.float64))[:, 0]
if rank == 0:
    p_1, a_1, r_1, _ = mref_ali2d(images, refs, ou=36, xr=3, yr=3, ts=1, maxit=3, device=local)
    same = float((a_n == a_1).mean()); dr = float(np.abs(r_n - r_1).max() / np.abs(r_1).max())
    good = same > 0.995 and dr < 1e-3
    ok &= good
    print("mref_ali2d    %d GPUs vs 1: assignments equal %.4f, references max rel diff %.2e  %s" % (world, same, dr, "OK" if good else "FAIL"))
# class-bound reference-free alignment, 2 passes with the tangent filter
cls = np.sort(np.random.default_rng(5).integers(0, R, P)).astype(np.int32)
q_n, f_n, _ = ref_free_alignment_2d(images[s:e], cls[s:e], refs, ou=36, xr=2, yr=2, ts=1, maxit=2, filt=(0.25, 0.2), comm=comm,
                                    global_offset=s, device=local)
q_n = gather(q_n)
if rank == 0:
    q_1, f_1, _ = ref_free_alignment_2d(images, cls, refs, ou=36, xr=2, yr=2, ts=1, maxit=2, filt=(0.25, 0.2), device=local)
    differ = np.abs(q_n - q_1).max(axis=1) >= 1e-3
    same = float((~differ).mean()); dr = float(np.abs(f_n - f_1).max() / np.abs(f_1).max())
    # The class sums are float reductions (atomics, then NCCL), so the averages of the first pass differ in the last
    # bits between ANY two runs; a particle sitting on a tie of the second pass may then take the other answer, and that
    # particle moves its class average by up to ~2 max / (class size).  The averages of the classes WITHOUT such a
    # particle must agree to the reduction-order bar.
    touched = np.unique(cls[differ])
    clean = np.setdiff1d(np.arange(R), touched)
    dr_clean = float(np.abs(f_n[clean] - f_1[clean]).max() / np.abs(f_1).max()) if clean.size else 0.0
    bar = 1e-3 + sum(2.0 * int((differ & (cls == c)).sum()) / max(int((cls == c).sum()), 1) for c in touched)
    good = same > 0.995 and dr_clean < 1e-3 and dr < bar
    ok &= good
    print("ref_free_2d   %d GPUs vs 1: parameters equal %.4f (%d particles on a tie, classes %s), references max rel diff %.2e "
          "(classes without such a particle: %.2e)  %s" % (world, same, int(differ.sum()), touched.tolist(), dr, dr_clean, "OK" if good else "FAIL"))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
