"""Iteration orchestrators: the B200 counterparts of the reference's mref_ali2d_gpu
(test_mref_gpu_align.py:222-612; CPU semantics test_mref.py:48-315) and
ali2d_base_gpu_isac_CLEAN / ali2d_base (test_reffree.py:111-513, :515-837).

One process per GPU.  Particles are split contiguously with MPI_start_end
(test_mref_gpu_align.py:1384), references are replicated, and the only exchange per
iteration is ONE allreduce of the packed class sums + counts (replaces 2R
reduce_EMData_to_root + R mpi_reduce + R bcast_EMData_to_all, test_mref.py:219-223,
:293-296).  Every rank then runs the deterministic reference update, so no broadcast follows.
"""
import numpy as np

from . import alignment as al
from . import refupdate as ru


class LocalComm(object):
    """Single-process stand-in: rank 0 of 1."""
    rank, world = 0, 1

    def allreduce_device(self, engine):
        return

    def allreduce_host(self, arr):
        return arr

    def fetch_image(self, k, owner_fn, local_get, shape):
        return local_get(k)


class TorchComm(object):
    """torch.distributed plumbing (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.backend = dist.get_backend()

    def allreduce_device(self, engine):
        """In-place sum of the engine's [R][2][nx][nx]+[R] buffer across ranks over NCCL."""
        import torch
        ptr, n = engine.sums_device_ptr()

        class _Buf(object):
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        t = torch.as_tensor(_Buf(), device="cuda:%d" % engine.device)
        self.dist.all_reduce(t)
        torch.cuda.synchronize(engine.device)

    def allreduce_host(self, arr):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if self.backend == "nccl":
            t = t.cuda()
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def fetch_image(self, k, owner_fn, local_get, shape):
        import torch
        owner = owner_fn(k)
        t = torch.from_numpy(local_get(k).copy()) if owner == self.rank else torch.zeros(shape, dtype=torch.float32)
        if self.backend == "nccl":
            t = t.cuda()
        self.dist.broadcast(t, src=owner)
        return t.cpu().numpy()


def _owner_fn(P, world):
    bounds = [al.mpi_start_end(P, world, i) for i in range(world)]

    def owner(k):
        for i, (s, e) in enumerate(bounds):
            if s <= k < e:
                return i
        raise IndexError(k)
    return owner


def _reduce_sums(engine, comm, device_allreduce):
    if comm.world > 1 and device_allreduce:
        comm.allreduce_device(engine)
        return engine.get_sums()
    sums, counts = engine.get_sums()
    if comm.world > 1:
        flat = comm.allreduce_host(np.concatenate([sums.ravel(), counts]))
        sums = flat[:sums.size].reshape(sums.shape)
        counts = flat[sums.size:]
    return sums, counts


def mref_ali2d(images, refs, ir=1, ou=-1, rs=1, xr=0, yr=0, ts=1, center=1, maxit=10, rand_seed=1000,
               params=None, comm=None, total_particles=None, global_offset=0, engine=None,
               device=0, device_allreduce=True, on_iteration=None, upload=True, device_update=None):
    """2D multi-reference alignment of this rank's particles.

    images  [n][nx][nx] float32: this rank's share (global indices global_offset..+n)
    refs    [R][nx][nx] initial references (replicated)
    device_update: run the reference update on the GPU (Engine.update_refs_device: the all-reduced class sums never
    leave the device; default whenever the sums are reduced on the device) instead of host numpy (refupdate.update_refs).
    Returns (params [n][4] alpha,sx,sy,mirror; assign [n]; refs [R][nx][nx]; history)
    """
    from .lib import Engine
    comm = comm or LocalComm()
    n, nx = images.shape[0], images.shape[-1]
    P = total_particles if total_particles is not None else n
    if ou == -1:
        ou = nx // 2 - 2
    R = refs.shape[0]
    own_engine = engine is None
    if own_engine:
        engine = Engine(nx, ou, xr, yr, ts=ts, ir=ir, rs=rs, max_particles=n, max_refs=R,
                        normalize_ring=True, device=device)
    if upload:
        engine.upload_particles(images, subtract_mask_mean=True)      # normalize.mask no_sigma=0, test_mref.py:188
    mask = ru.model_circle(ou, nx)
    refs = np.array(refs, np.float32)
    params = np.zeros((n, 4)) if params is None else np.array(params, np.float64)
    owner = _owner_fn(P, comm.world)
    masked_local = lambda k: ru.normalize_mask(np.asarray(images[k - global_offset], np.float32), mask, 0)
    reseed = ru.make_reseeder(rand_seed, P, lambda k: comm.fetch_image(k, owner, masked_local, (nx, nx)))
    history = []
    assign = np.zeros(n, np.int32)
    if device_update is None:
        device_update = bool(device_allreduce) and hasattr(engine, "update_refs_device")
    for it in range(int(maxit)):
        if it == 0 or not device_update:
            engine.set_refs(refs, normalize_mask=True)                # test_mref.py:170-175
        else:
            engine.prepare_refs(normalize_mask=True)                  # the references are already on the device
        search, sxi, syi, params = al.mref_search_request(params, nx, ou, xr, yr)
        res = engine.align(0, n, search)                              # test_mref.py:200
        params = al.compose_result(sxi, syi, res)                     # test_mref.py:206
        assign = res["iref"].copy()
        engine.zero_sums()
        engine.accumulate(0, n, params, assign, global_offset)        # test_mref.py:210-215
        if device_update:
            if comm.world > 1:
                comm.allreduce_device(engine)                         # test_mref.py:219-223
            last = it == int(maxit) - 1
            refs_d, info = engine.update_refs_device(center, reseed, fetch=bool(on_iteration) or last)   # test_mref.py:238-286
            if refs_d is not None:
                refs = refs_d
            counts = engine.get_sums_counts() if hasattr(engine, "get_sums_counts") else engine.get_sums()[1]
        else:
            sums, counts = _reduce_sums(engine, comm, device_allreduce)   # test_mref.py:219-223
            refs, info = ru.update_refs(sums[:R], counts[:R], mask, center, reseed)   # test_mref.py:238-286
        info.update(counts=counts[:R].copy(), peak=res["peak"].copy(), stats=engine.stats())
        history.append(info)
        if on_iteration:
            on_iteration(it, params, assign, refs, info)
    if own_engine:
        engine.close()
    return params, assign, refs, history


def search_schedule(xr, yr, ts):
    """The reference parses --xr/--yr/--ts with get_input_from_string (test_reffree.py:173-174):
    "4 2 1 1" / [4, 2, 1, 1] / 4 -> one (xr, yr, ts) per step; yr = -1 copies xr, a shorter list repeats
    its last entry.  The shipped driver then pins N_step = 0 (test_reffree.py:310, :686); ali2d_base
    here runs the whole schedule as Sphire's ali2d_base does."""
    def lst(v):
        if isinstance(v, str):
            return [float(x) for x in v.split()]
        try:
            return [float(x) for x in v]
        except TypeError:
            return [float(v)]
    xs, ys, tss = lst(xr), lst(yr), lst(ts)
    if len(ys) == 1 and ys[0] == -1:
        ys = list(xs)
    n = len(xs)
    ys = ys + [ys[-1]] * (n - len(ys))
    tss = tss + [tss[-1]] * (n - len(tss))
    return [(xs[i], ys[i], tss[i]) for i in range(n)]


def ali2d_base(images, ir=1, ou=-1, rs=1, xr=0, yr=0, ts=1, center=-1, maxit=10, comm=None,
               total_particles=None, global_offset=0, engine=None, device=0, device_allreduce=True,
               on_iteration=None):
    """Reference-free alignment (ali2d_base, test_reffree.py:515-837): one reference = the global
    average, ormq semantics (no ring normalisation), shifts clamped, average centred by the mean shift."""
    from .lib import Engine
    comm = comm or LocalComm()
    n, nx = images.shape[0], images.shape[-1]
    P = total_particles if total_particles is not None else n
    if ou == -1:
        ou = nx // 2 - 2
    sched = search_schedule(xr, yr, ts)
    own_engine = engine is None
    if own_engine:
        # sized for the largest window of the schedule: most positions per axis = max(range / step)
        ts_min = min(t for _, _, t in sched)
        kmax = max(int(max(x, y) / t) for x, y, t in sched)
        engine = Engine(nx, ou, kmax * ts_min, kmax * ts_min, ts=ts_min, ir=ir, rs=rs, max_particles=n, max_refs=1,
                        normalize_ring=False, device=device)
    engine.upload_particles(images, subtract_mask_mean=True)          # data[im] -= infomask mean, test_reffree.py:663
    mask = ru.model_circle(ou, nx)
    params = np.zeros((n, 4))
    sx_sum = sy_sum = 0.0
    history = []
    tavg = None
    zeros = np.zeros(n, np.int32)
    total = len(sched) * int(maxit)
    for it in range(total):
        xr, yr, ts = sched[it // int(maxit)]                          # for N_step: for Iter (Sphire ali2d_base)
        engine.set_step(ts)
        engine.zero_sums()
        engine.accumulate(0, n, params, zeros, global_offset)         # sum_oe(data, "a"), test_reffree.py:695
        sums, counts = _reduce_sums(engine, comm, device_allreduce)
        ave1, ave2 = sums[0, 0], sums[0, 1]
        tavg = (ave1 + ave2) / np.float32(P)
        frsc = ru.fsc_mask(ave1, ave2, mask)                          # test_reffree.py:708
        crit = float(np.sum((tavg * tavg)[mask > 0.5]) / float((mask > 0.5).sum()))
        if center == -1:
            tavg, _, filt = ru.ref_ali2d(tavg, frsc, 0)
            cs = [float(sx_sum) / P, float(sy_sum) / P]
            tavg = ru.fshift(tavg, -cs[0], -cs[1])                    # test_reffree.py:741-745
        else:
            tavg, cs, filt = ru.ref_ali2d(tavg, frsc, center)
        if it == total - 1:
            history.append(dict(criterion=crit, cs=cs, filter=filt, tavg=tavg.copy()))
            break
        engine.set_refs(tavg[None], normalize_mask=False)
        search, sxi, syi = al.reffree_search_request(params, cs, nx, ou, xr, yr)
        res = engine.align(0, n, search)                              # ali2d_single_iter -> ormq
        params = al.compose_result(sxi, syi, res)
        loc = np.array([np.sum(np.where(params[:, 3] == 0, params[:, 1], -params[:, 1])), np.sum(params[:, 2])])
        glob = comm.allreduce_host(loc) if comm.world > 1 else loc
        sx_sum, sy_sum = float(glob[0]), float(glob[1])
        info = dict(criterion=crit, cs=cs, filter=filt, tavg=tavg.copy(), peak=res["peak"].copy(), stats=engine.stats())
        history.append(info)
        if on_iteration:
            on_iteration(it, params, tavg, info)
    if own_engine:
        engine.close()
    return params, tavg, history


def ref_free_alignment_2d(images, class_of, refs, ir=1, ou=-1, rs=1, xr=0, yr=0, ts=1, maxit=1, filt=None, comm=None,
                          global_offset=0, engine=None, device=0, device_allreduce=True, on_iteration=None):
    """gpu_isac's class-bound reference-free alignment (ref_free_alignment_2D, cuda/gpu_aln_noref.cu:559-782):
    every particle is aligned to the average of its own class only (class_of = sbj_cid_list), the class
    averages are rebuilt ON THE DEVICE from the transformed particles after every pass and optionally
    low-passed with the tangent filter (filt = (cutoff, falloff)).  ormq semantics as in ali2d_base.
    With N > 1 ranks the particles are sharded, the references replicated and the class sums allreduced,
    so every rank rebuilds the same averages.  Returns (params [n][4], refs [R][nx][nx], history)."""
    from .lib import Engine
    comm = comm or LocalComm()
    n, nx = images.shape[0], images.shape[-1]
    refs = np.array(refs, np.float32)
    class_of = np.ascontiguousarray(class_of, np.int32)
    R = refs.shape[0]
    if ou == -1:
        ou = nx // 2 - 2
    own_engine = engine is None
    if own_engine:
        engine = Engine(nx, ou, xr, yr, ts=ts, ir=ir, rs=rs, max_particles=n, max_refs=R, normalize_ring=False,
                        device=device)
    engine.upload_particles(images, subtract_mask_mean=True)
    engine.set_refs(refs, normalize_mask=False)
    params = np.zeros((n, 4))
    history = []
    for it in range(int(maxit)):
        search, sxi, syi = al.reffree_search_request(params, (0.0, 0.0), nx, ou, xr, yr)
        res = engine.align_bound(0, n, search, class_of)
        params = al.compose_result(sxi, syi, res)
        engine.zero_sums()
        engine.accumulate(0, n, params, class_of, global_offset)
        if device_allreduce:
            if comm.world > 1:
                comm.allreduce_device(engine)
            engine.refs_from_sums(normalize_mask=False)
            if filt is not None:
                engine.filter_refs(filt[0], filt[1], normalize_mask=False)
        else:
            # host exchange (gloo / no NCCL): same arithmetic on the reduced sums, references re-uploaded
            sums, counts = _reduce_sums(engine, comm, False)
            for r in range(R):
                if counts[r] > 0.5:
                    refs[r] = (sums[r, 0] + sums[r, 1]) / np.float32(counts[r])
                if filt is not None:
                    refs[r] = ru.filt_tanl(refs[r], filt[0], filt[1])
            engine.set_refs(refs, normalize_mask=False)
        info = dict(peak=res["peak"].copy(), stats=engine.stats())
        history.append(info)
        if on_iteration:
            on_iteration(it, params, info)
    out_refs = engine.get_refs() if device_allreduce else refs
    if own_engine:
        engine.close()
    return params, out_refs, history
