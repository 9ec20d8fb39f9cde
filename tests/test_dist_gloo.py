"""World-size-2 run of the iteration orchestrator over gloo on CPU.  The device engine is
replaced by a test double that answers through the CPU oracle (allowed in tests/ only), so what is
exercised is the product's sharding, the single allreduce of class sums + counts, the
rank-independent reseeding broadcast and the redundant reference update."""
import os
import socket

import numpy as np
import pytest


class OracleEngine(object):
    """Engine-shaped test double backed by the oracle."""

    def __init__(self, nx, ou, R, xr, ts):
        from oracle import oracle as o
        self.o, self.nx, self.ou, self.max_refs, self.ts = o, nx, ou, R, ts
        self.mask = o.model_circle(ou, nx)
        self.numr = o.numrinit(1, ou, 1)
        self.device = 0

    def upload_particles(self, images, first=0, subtract_mask_mean=True):
        self.imgs = np.stack([self.o.normalize_mask(im, self.mask, 0) for im in images])

    def set_refs(self, refs, normalize_mask=True):
        o = self.o
        if normalize_mask:
            _, self.cref = o.prepare_refs(refs, self.mask, self.numr)
        else:
            c = float(self.nx // 2 + 1)
            wr = o.ringwe(self.numr)
            self.cref = np.stack([o.applyws(o.frngs(o.polar2dm(r, c, c, self.numr), self.numr), self.numr, wr) for r in refs])

    def align_bound(self, start, stop, search, class_of):
        from cryo_ralib_b200.lib import RESULT_DTYPE
        r = np.zeros(stop - start, RESULT_DTYPE)
        cen = np.stack([search["cx"], search["cy"]], 1)
        win = np.stack([search["xl"], search["xr"], search["yl"], search["yr"]], 1)
        for c in np.unique(class_of):
            idx = np.nonzero(class_of == c)[0]
            out = self.o.align_batch(self.imgs[start:stop][idx], self.cref[c:c + 1], self.numr, cen[idx], win[idx], self.ts, False, 2)
            for k, col in zip(("ang", "sxs", "sys", "mirror", "iref", "peak", "sx", "sy"), range(8)):
                r[k][idx] = out[:, col]
        r["iref"] = class_of
        return r

    def align(self, start, stop, search):
        from cryo_ralib_b200.lib import RESULT_DTYPE
        out = self.o.align_batch(self.imgs[start:stop], self.cref, self.numr, np.stack([search["cx"], search["cy"]], 1),
                                 np.stack([search["xl"], search["xr"], search["yl"], search["yr"]], 1), self.ts, True, 2)
        r = np.zeros(stop - start, RESULT_DTYPE)
        for k, c in zip(("ang", "sxs", "sys", "mirror", "iref", "peak", "sx", "sy"), range(8)):
            r[k] = out[:, c]
        return r

    def zero_sums(self):
        self.sums = np.zeros((self.max_refs, 2, self.nx, self.nx), np.float32)
        self.counts = np.zeros(self.max_refs, np.float32)

    def accumulate(self, start, stop, params, iref, global_offset=0):
        for i in range(stop - start):
            t = self.o.rot_shift2d(self.imgs[start + i], params[i, 0], params[i, 1], params[i, 2], int(params[i, 3]))
            self.sums[iref[i], (global_offset + i) % 2] += t
            self.counts[iref[i]] += 1

    def get_sums(self):
        return self.sums.copy(), self.counts.copy()

    def stats(self):
        return {}

    def close(self):
        pass


def _case():
    from cryo_ralib_b200 import synth
    images, _ = synth.make_particles(22, 64, 3, max_shift=1, seed=31)
    refs = synth.initial_references(images, 3, per_ref=4, seed=1)
    return images, refs


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cryo_ralib_b200 import alignment as al
    from cryo_ralib_b200.mref import mref_ali2d, TorchComm
    images, refs = _case()
    s, e = al.mpi_start_end(images.shape[0], world, rank)
    eng = OracleEngine(64, 28, 3, 1, 1.0)
    p, a, r, h = mref_ali2d(images[s:e], refs, ou=28, xr=1, yr=1, ts=1, maxit=2, comm=TorchComm(),
                            total_particles=images.shape[0], global_offset=s, engine=eng, device_allreduce=False)
    q.put((rank, s, e, p, a, r, [x["counts"] for x in h]))
    dist.destroy_process_group()


def test_two_rank_gloo_equals_single_rank():
    import torch.multiprocessing as mp
    from cryo_ralib_b200.mref import mref_ali2d
    images, refs = _case()
    eng = OracleEngine(64, 28, 3, 1, 1.0)
    p1, a1, r1, h1 = mref_ali2d(images, refs, ou=28, xr=1, yr=1, ts=1, maxit=2, engine=eng)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    params = np.concatenate([g[3] for g in got]); assign = np.concatenate([g[4] for g in got])
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == images.shape[0]
    assert np.array_equal(got[0][5], got[1][5])                      # every rank holds the same references
    assert np.array_equal(assign, a1)
    assert np.allclose(params, p1, atol=1e-6)
    assert np.abs(got[0][5] - r1).max() <= 1e-5 * np.abs(r1).max()   # allreduce order only
    assert all(np.array_equal(c0, c1) for c0, c1 in zip(got[0][6], [x["counts"] for x in h1]))
    assert sum(got[0][6][-1]) == images.shape[0]


def _bound_case():
    images, refs = _case()
    cls = np.array([0] * 9 + [2] * 13, np.int32)                    # runs of classes; class 1 empty; a run spans the rank split
    return images, refs, cls


def _bound_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cryo_ralib_b200 import alignment as al
    from cryo_ralib_b200.mref import ref_free_alignment_2d, TorchComm
    images, refs, cls = _bound_case()
    s, e = al.mpi_start_end(images.shape[0], world, rank)
    eng = OracleEngine(64, 28, 3, 1, 1.0)
    p, r, h = ref_free_alignment_2d(images[s:e], cls[s:e], refs, ou=28, xr=1, yr=1, ts=1, maxit=2, filt=(0.3, 0.2),
                                    comm=TorchComm(), global_offset=s, engine=eng, device_allreduce=False)
    q.put((rank, p, r))
    dist.destroy_process_group()


def test_two_rank_class_bound_alignment_equals_single_rank_and_oracle():
    """ref_free_alignment_2d (gpu_isac's class-bound variant) sharded over two ranks: same parameters as one
    rank, identical references on every rank, and both equal the oracle twin."""
    import torch.multiprocessing as mp
    from oracle import oracle as o
    from cryo_ralib_b200.mref import ref_free_alignment_2d
    images, refs, cls = _bound_case()
    eng = OracleEngine(64, 28, 3, 1, 1.0)
    p1, r1, _ = ref_free_alignment_2d(images, cls, refs, ou=28, xr=1, yr=1, ts=1, maxit=2, filt=(0.3, 0.2), engine=eng,
                                      device_allreduce=False)
    p0, r0, _ = o.ref_free_alignment_2d(images, cls, refs, ou=28, xr=1, yr=1, ts=1, maxit=2, filt=(0.3, 0.2), nthreads=2)
    assert np.allclose(p1, p0, atol=1e-5) and np.abs(r1 - r0).max() <= 1e-5 * np.abs(r0).max()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bound_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert np.array_equal(got[0][2], got[1][2])
    assert np.allclose(np.concatenate([got[0][1], got[1][1]]), p1, atol=1e-5)
    assert np.abs(got[0][2] - r1).max() <= 1e-5 * np.abs(r1).max()
