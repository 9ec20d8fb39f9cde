#!/usr/bin/env python
"""Benchmark of the 2D multi-reference alignment hot path (BASELINE.json metric:
particle x reference x shift alignments/s, and s/iteration).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-gpu]
                    [--config 1..5] [--variant clipped|ou56] [--scaling weak|strong]

A "step" is one iteration of the per-particle section of mref_ali2d (test_mref.py:170-223):
reference preparation, multiref_polar_ali_2d for every particle, rot_shift2D + even/odd class
sums, and (N > 1) one allreduce of the sums.  Default workload at every N: BASELINE.json
configs[1] (100k synthetic 90x90 particles, 50 references, ou=36, xr=yr=3, ts=1, mirror on) PER
GPU (weak scaling).  --config selects the other named configurations (CONFIGS below; numbering
1..5 = BASELINE.json configs[0..4]); --scaling strong shards the configuration's particle count
over the ranks with MPI_start_end exactly as test_mref_gpu_align.py:1384 does.  `value` is timed
with the particle stack resident in HBM; `e2e` runs the same step through the C ABI from pinned
host buffers (H2D of the stack, D2H of parameters and class sums inside the timed region).
`--impl reference` times the CPU oracle port of the reference's EMAN2 path (oracle/, all host
threads) on a bounded sample of the same workload; `--impl reference-gpu` times the reference's
own CUDA library (cuda/gpu_aln_*.cu built for sm_100 by baseline/build_ref_cuda.sh into
baseline/_ref/, driven through its stock entry points) -- a throughput bar, not a parity oracle
(its arithmetic is gpu_isac's).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs (1-based).  mode "mref": multiref_polar_ali_2d semantics (Normalize_ring on);
# "reffree": one reference = the running average, ormq semantics (test_reffree.py:780-783).
CONFIGS = {
    1: dict(P=10000, nx=90, R=10, ou=36, xr=3, yr=3, ts=1.0, nviews=64, mode="mref",
            name="mref 10k synthetic 90x90 particles, 10 refs, ou=36 xr=yr=3 ts=1 (BASELINE configs[0], the CPU driver's case)"),
    2: dict(P=100000, nx=90, R=50, ou=36, xr=3, yr=3, ts=1.0, nviews=64, mode="mref",
            name="mref 100k synthetic 90x90 particles, 50 refs, ou=36 xr=yr=3 ts=1, mirror on (BASELINE configs[1])"),
    3: dict(P=50000, nx=90, R=1, ou=36, xr=3, yr=3, ts=1.0, nviews=64, mode="reffree",
            name="reference-free 50k synthetic 90x90 particles, ou=36 xr=yr=3 ts=1 (BASELINE configs[2])"),
    4: dict(P=200000, nx=128, R=200, ou=60, xr=6, yr=6, ts=1.0, nviews=200, mode="mref",
            name="mref 200k synthetic 128x128 particles, 200 refs, ou=60 xr=yr=6 (BASELINE configs[3]; search_range clips the "
                 "window to <=49 positions as the reference's CPU path does, test_mref.py:195-198)"),
    5: dict(P=100000, nx=90, R=500, ou=36, xr=8, yr=8, ts=0.5, nviews=500, mode="mref", gpus_named=8,
            name="mref shift-grid sweep ts=0.5 xr=yr=8 (1089 positions), 500 refs, 100k synthetic 90x90 particles on 8 GPUs "
                 "(BASELINE configs[4]; particle count assumed, SURVEY 8d)"),
}
VARIANTS = {"ou56": dict(ou=56, name_suffix="; variant ou=56: the full 169-position grid is legal (SURVEY 8d recommendation ii)")}
CFG = CONFIGS[2]
METRIC = "particle x reference x shift alignments/s"


def flops_per_alignment(lcirc, maxrin):
    # contraction 4*lcirc + two inverse FFTs 5*maxrin*log2(maxrin)
    return 4.0 * lcirc + 5.0 * maxrin * np.log2(maxrin)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                    power_w_max=float(max(pw)), samples=len(sm))


def select_config(args):
    cfg = dict(CONFIGS[args.config])
    if args.variant != "clipped":
        v = VARIANTS[args.variant]
        cfg["ou"] = v["ou"]
        cfg["name"] += v["name_suffix"]
    if args.particles:
        cfg["P"] = args.particles
    cfg["S_nominal"] = (2 * int(cfg["xr"] / cfg["ts"]) + 1) * (2 * int(cfg["yr"] / cfg["ts"]) + 1)
    return cfg


def window_of(cfg):
    """(xl, xr, yl, yr) search_range leaves at zero accumulated shift (test_mref.py:195-198)."""
    from cryo_ralib_b200 import alignment as al
    xl, xr_ = al.search_range(cfg["nx"], cfg["ou"], 0.0, cfg["xr"])
    yl, yr_ = al.search_range(cfg["nx"], cfg["ou"], 0.0, cfg["yr"])
    return float(xl), float(xr_), float(yl), float(yr_)


def positions_of(cfg):
    xl, xr_, yl, yr_ = window_of(cfg)
    ts = cfg["ts"]
    return (int(xl / ts) + int(xr_ / ts) + 1) * (int(yl / ts) + int(yr_ / ts) + 1)


def oracle_prepare(images, refs, cfg):
    from oracle import oracle as o
    o.build()
    nx, ou = cfg["nx"], cfg["ou"]
    mask = o.model_circle(ou, nx)
    numr = o.numrinit(1, ou, 1)
    imgs = np.stack([o.normalize_mask(im, mask, 0) for im in images])
    if cfg["mode"] == "reffree":                   # ormq: the average is used as it is (no normalize.mask), test_reffree.py:755
        cnx = float(nx // 2 + 1)
        wr = o.ringwe(numr)
        cref = np.stack([o.applyws(o.frngs(o.polar2dm(np.asarray(r, np.float32), cnx, cnx, numr), numr), numr, wr) for r in refs])
    else:
        _, cref = o.prepare_refs(refs, mask, numr)
    return o, imgs, cref, numr


def oracle_runner(images, refs, cfg, nthreads):
    """Returns run(n) -> seconds for the CPU restatement of Util.multiref_polar_ali_2d on the first n particles."""
    o, imgs, cref, numr = oracle_prepare(images, refs, cfg)
    cnx = cfg["nx"] // 2 + 1
    win1 = np.array(window_of(cfg), np.float32)
    norm_ring = cfg["mode"] == "mref"

    def run(n):
        centres = np.full((n, 2), float(cnx), np.float32)
        win = np.tile(win1, (n, 1))
        t = time.perf_counter()
        o.align_batch(imgs[:n], cref, numr, centres, win, cfg["ts"], norm_ring, nthreads)
        return time.perf_counter() - t
    return run


def oracle_time_sample(images, refs, cfg, nthreads, target_s=15.0):
    """Time the CPU restatement on a bounded sample; returns (alignments/s, n_particles, seconds)."""
    run = oracle_runner(images, refs, cfg, nthreads)
    S = positions_of(cfg)
    n0 = min(len(images), max(nthreads, 4))
    t0 = run(n0)
    n = int(min(len(images), max(n0, n0 * target_s / max(t0, 1e-6))))
    t = run(n) if n > n0 else t0
    return n * S * cfg["R"] / t, n, t


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port; EMAN2 is not installable),
    all host threads, bounded sample per step."""
    if rank != 0:
        return
    from cryo_ralib_b200 import synth
    cfg = select_config(args)
    nth = host_threads()
    S = positions_of(cfg)
    # a sample that costs about 6 s per step: the cost per particle scales with S * (row + R * pair) work
    nsample = int(max(2 * nth, min(4096, 4096 * (49.0 * 50 * 5816) / (S * max(cfg["R"], 8) * (5816 if cfg["nx"] == 90 else 17080)))))
    images, _ = synth.make_particles(nsample, cfg["nx"], min(cfg["nviews"], 64), max_shift=int(cfg["xr"]), seed=2025)
    refs = synth.initial_references(images, cfg["R"], seed=99)
    run = oracle_runner(images, refs, cfg, nth)
    probe = min(nsample, 2 * nth)
    run(min(nsample, nth))                          # thread start-up and first-touch costs stay out of the probe
    t_probe = run(probe)
    n = int(max(probe, min(nsample, probe * 5.0 / max(t_probe, 1e-6))))
    times = []
    for it in range(args.warmup + args.steps):
        dt = run(n)
        if it >= args.warmup:
            times.append(dt)
    tot = sum(times)
    value = n * S * cfg["R"] * len(times) / tot
    sample = "%d of %d particles per step (x%d refs x%d shifts), %d host threads" % (n, cfg["P"], cfg["R"], S, nth)
    line = dict(impl="reference", metric=METRIC, value=value, unit="alignments/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * tot / len(times), higher_is_better=True, scaling=args.scaling,
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=cfg["name"] + "; CPU oracle port on a bounded sample",
                            particles_per_step=n, refs=cfg["R"], shifts=S, nx=cfg["nx"], ou=cfg["ou"]),
                cpu_baseline=dict(value=value, unit="alignments/s", cores=nth, kind="port", sample=sample),
                e2e=dict(value=value, unit="alignments/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                s_per_iteration_extrapolated=cfg["P"] * S * cfg["R"] / value)
    print(json.dumps(line))


def run_reference_gpu(args, rank, world):
    """--impl reference-gpu: the reference's own CUDA library on this box (see baseline/ref_cuda.py)."""
    if rank != 0:
        return
    from baseline import ref_cuda
    cfg = select_config(args)
    print(json.dumps(ref_cuda.bench_line(cfg, args, METRIC)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs, 1-based (default 2 = configs[1])")
    ap.add_argument("--variant", default="clipped", choices=["clipped"] + sorted(VARIANTS), help="config 4: ou=60 with search_range clipping (default) or ou=56 (full grid)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--particles", type=int, default=0, help="override the configuration's particle count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer pass (probing only: the line is then not a bench line)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.impl == "reference-gpu":
        run_reference_gpu(args, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    from cryo_ralib_b200 import Engine, synth, alignment as al, refupdate as ru
    from cryo_ralib_b200.mref import TorchComm, LocalComm
    torch.cuda.set_device(local_rank)
    comm = LocalComm()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = TorchComm()
    cfg = select_config(args)
    nx, R, ou, xr, yr, ts, mode = cfg["nx"], cfg["R"], cfg["ou"], cfg["xr"], cfg["yr"], cfg["ts"], cfg["mode"]
    Ptot = cfg["P"]
    if args.scaling == "strong":
        # the configuration's stack sharded over the ranks exactly as the reference does (test_mref_gpu_align.py:1384)
        goff, gend = al.mpi_start_end(Ptot, world, rank)
        P = gend - goff
        share = "MPI_start_end shard of %d particles over %d GPUs" % (Ptot, world)
    else:
        # weak scaling: a fixed share per GPU.  A configuration named for k GPUs contributes 1/k of its stack per GPU.
        P = Ptot // cfg.get("gpus_named", 1)
        goff = rank * P
        share = "%d particles per GPU%s" % (P, " (1/%d of the configuration's stack: it is named for %d GPUs)" % (cfg["gpus_named"], cfg["gpus_named"]) if cfg.get("gpus_named") else "")
    W = max(args.warmup, 3)

    # synthetic particles, generated on the device (plumbing) -- seed differs per rank
    dev = "cuda:%d" % local_rank
    images_d, _ = synth.make_particles(P, nx, cfg["nviews"], max_shift=int(xr), seed=2025 + rank, device=dev)
    if mode == "reffree":
        refs_t = images_d[:min(P, 2000)].mean(dim=0, keepdim=True)
    else:
        refs_t = synth.initial_references(images_d, R, seed=99)
    if world > 1:                                   # references are replicated: every rank uses rank 0's
        torch.distributed.broadcast(refs_t, src=0)
    refs = refs_t.cpu().numpy()
    host_images = torch.empty((P, nx, nx), dtype=torch.float32, pin_memory=True)
    host_images.copy_(images_d)
    torch.cuda.synchronize()

    eng = Engine(nx, ou, xr, yr, ts=ts, max_particles=P, max_refs=R, normalize_ring=(mode == "mref"), device=local_rank)
    eng.upload_particles_dev(images_d.data_ptr(), P, subtract_mask_mean=True)
    del images_d
    torch.cuda.empty_cache()
    fp32_peak = eng.measure_fp32_peak() if rank == 0 else (0.0, 0.0)
    stream = torch.cuda.ExternalStream(eng.L.cra_stream(eng.h), device=dev)
    params0 = np.zeros((P, 4))
    mask = ru.model_circle(ou, nx)

    # e2e: the stack is uploaded in chunks on the copy stream, each aligned as soon as it landed.  Chunks are whole
    # row batches of the engine (no partial launches in between); the first one is small so the alignment starts early.
    S = positions_of(cfg)
    per_batch = max(1, eng.L.cra_row_batch(eng.h) // S)
    # chunk sizes 1, 1, 2, 4, ... row batches: the host link (~22 GB/s measured here) delivers a chunk no bigger than
    # everything before it while those are aligned (~8 GB/s of images), so only the first batch's upload is exposed
    bounds, s0, nb, done = [], 0, 1, 0
    while s0 < P:
        e0 = min(P, s0 + nb * per_batch)
        bounds.append((s0, e0))
        done += nb
        s0, nb = e0, min(32, done)

    trace = os.environ.get("CRA_BENCH_TRACE") and rank == 0
    zeros_i = np.zeros(P, np.int32)

    def request(params):
        if mode == "reffree":
            search, sxi, syi = al.reffree_search_request(params, (0.0, 0.0), nx, ou, xr, yr)
            return search, sxi, syi, params
        return al.mref_search_request(params, nx, ou, xr, yr)

    def step(resident, params, full=False):
        """One iteration of the per-particle section; returns (new params, assign, stats).  full: followed by
        the reference update (test_mref.py:238-286), i.e. a whole user-level iteration."""
        tr = [("start", time.perf_counter())]
        eng.set_refs(refs, normalize_mask=(mode == "mref"))      # before the stack is queued: a copy from pageable memory waits behind it
        if not resident:
            base = host_images.data_ptr()
            for s, e in bounds:
                eng.upload_particles_async(base + s * nx * nx * 4, e - s, first=s, subtract_mask_mean=True)
            tr.append(("queued", time.perf_counter()))
        search, sxi, syi, params = request(params)
        tr.append(("refs+request", time.perf_counter()))
        if resident:
            res = eng.align(0, P, search)
            st = eng.stats()
        else:
            parts, st = [], None
            for s, e in bounds:
                parts.append(eng.align(s, e, search[s:e]))
                t = eng.stats()
                tr.append(("align %d (kernels %.1f ms)" % (e - s, t["ms_total"]), time.perf_counter()))
                st = t if st is None else {k: st[k] + t[k] for k in st}
            res = np.concatenate(parts)
        tr.append(("align", time.perf_counter()))
        newp = al.compose_result(sxi, syi, res)
        eng.zero_sums()
        eng.accumulate(0, P, newp, res["iref"] if mode == "mref" else zeros_i, goff)
        if world > 1:
            comm.allreduce_device(eng)
        if not resident:
            sums, counts = eng.get_sums()
        tr.append(("sums", time.perf_counter()))
        if full and mode == "mref":
            # the reference update on the device (class sums stay in HBM; the host fits the tangent filter)
            eng.update_refs_device(center=1, reseed=lambda j: host_images[j % P].numpy(), fetch=True)
            tr.append(("update_refs", time.perf_counter()))
        if trace:
            sys.stderr.write("trace %s: " % ("resident" if resident else "e2e") +
                             ", ".join("%s +%.1f" % (n, 1e3 * (t - tr[i][1])) for i, (n, t) in enumerate(tr[1:])) + "\n")
        return newp, res, st

    def timed(resident, nsteps, params, full=False):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        agg = dict(ms_polar=0.0, ms_ccf=0.0, ms_final=0.0, launches=0, alignments=0, rows=0, ccf_launches=0)
        for _ in range(nsteps):
            params, res, st = step(resident, params, full)
            for k in ("ms_polar", "ms_ccf", "ms_final", "launches", "alignments", "rows"):
                agg[k] += st[k]
            agg["ccf_launches"] += st["launches"] // 3
            agg["launches"] += 2 + 1 + (1 if not resident else 0)   # mask-normalise + polar(refs), rot/sum, (mask-normalise on upload)
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, agg, params

    eng.set_timing(True)
    p = params0
    for _ in range(W):
        p, _, _ = step(True, p)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res, agg, p = timed(True, args.steps, p)
    if args.no_e2e:
        ms_e2e, agg_e = float("nan"), agg
    else:
        p, _, _ = step(False, p)                      # one untimed pass through the host-buffer path
        ms_e2e, agg_e, p = timed(False, args.steps, p)
    clocks = sampler.stop() if rank == 0 else None
    # a whole user-level iteration (alignment + class sums + allreduce + reference update), resident stack
    ms_full, _, p = timed(True, max(1, min(args.steps, 2)), p, full=True)
    ms_full /= max(1, min(args.steps, 2))

    # alignments of one step over ALL ranks (strong scaling: the shards differ by a particle)
    aligns_all = float(agg["alignments"]) / args.steps
    if world > 1:
        t = torch.tensor([aligns_all], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t)
        aligns_all = float(t.item())

    if rank == 0:
        value = aligns_all * args.steps / (ms_res * 1e-3)
        e2e = aligns_all * args.steps / (ms_e2e * 1e-3)
        fpa = flops_per_alignment(eng.lcirc, eng.maxrin)
        ccf_s = agg["ms_ccf"] * 1e-3
        ccf_tflops = agg["alignments"] * fpa / ccf_s / 1e12
        polar_gbs = agg["rows"] * eng.lcirc * 4.0 / (agg["ms_polar"] * 1e-3) / 1e9
        polar_flops = agg["rows"] * (45.0 * eng.lcirc + 2.5 * sum(float(l) * np.log2(l) for l in eng.numr[2::3]))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        if args.config != 2:
            prof = {}                                 # the ncu traffic figures were captured on configuration 2
        fp32 = max(fp32_peak)
        # a kernel timed inside a long step: the sustained figure
        tensor_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0))
        tensor_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else "fallback 1400 sustained (B200_PROFILING.md)"
        h2d = P * nx * nx * 4 + R * nx * nx * 4 + P * 24 + P * 20
        d2h = P * 32 + (R * 2 * nx * nx + R) * 4
        polar_share = agg["ms_polar"] / ms_res
        ccf_share = agg["ms_ccf"] / ms_res
        line = dict(metric=METRIC, value=value, unit="alignments/s", n_gpus=world, steps=args.steps, warmup=W,
                    ms_per_step=ms_res / args.steps, higher_is_better=True, scaling=args.scaling, vs_baseline=None,
                    dtype="f32", data="synthetic",
                    config=dict(workload=cfg["name"] + ("; " + share),
                                config_index=args.config, variant=args.variant, particles_per_gpu=P, particles_total=(Ptot if args.scaling == "strong" else P * world),
                                refs=R, nx=nx, ou=ou, xr=xr, yr=yr, ts=ts, shifts=S, shifts_nominal=cfg["S_nominal"],
                                lcirc=eng.lcirc, maxrin=eng.maxrin,
                                l2="inputs larger than L2 (%.1f GB stack + row-batch spectra of up to 2 GB); no flush" % (P * nx * nx * 4 / 1e9)),
                    s_per_iteration=ms_res / args.steps * 1e-3,
                    s_per_full_iteration=ms_full * 1e-3,
                    e2e=dict(value=e2e, unit="alignments/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                             ms_per_step=ms_e2e / args.steps),
                    gpu_launches=int(agg["launches"]),
                    clocks=clocks,
                    roofline=dict(kernel="ccf_tm_kernel (Crosrng_ms contraction on mma.sync split-bf16 x3, W in tensor memory, + inverse FFT + peak search)",
                                  bound="tensor", achieved=ccf_tflops, peak=tensor_peak, unit="TFLOP/s",
                                  frac=ccf_tflops / tensor_peak,
                                  traffic=prof.get("ccf_dram_bytes_per_launch"),
                                  peak_source=tensor_src,
                                  note="algorithmic FP32-semantics flops (4*lcirc + 5*maxrin*log2(maxrin) per alignment, SURVEY 8d) "
                                       "over the measured dense bf16 peak; the contraction needs 3 bf16 MMAs per product (split "
                                       "precision), K <= 36 rules out tcgen05 tiles, and 30 % of the flops are the FP32 inverse FFT: "
                                       "see roofline_fp32_equiv and DESIGN.md 3.2 for the ceilings that actually bind",
                                  flops_per_alignment=fpa, avg_launch_ms=agg["ms_ccf"] / max(agg["ccf_launches"], 1),
                                  share_of_step=ccf_share),
                    roofline_fp32_equiv=dict(kernel="ccf_tm_kernel", bound="fp32", achieved=ccf_tflops, peak=fp32, unit="TFLOP/s",
                                             frac=ccf_tflops / fp32 if fp32 else None,
                                             note="same algorithmic flops over the FFMA micro-benchmark measured in this run "
                                                  "(what an FP32 SIMT implementation could reach at best)"),
                    roofline_composite=dict(kernel="ccf_tm_kernel", bound="tensor + fp32",
                                            floor_ms_per_step=(agg["alignments"] * (4.0 * eng.lcirc * 3.0 / (tensor_peak * 1e12)
                                                               + 5.0 * eng.maxrin * np.log2(eng.maxrin) / (fp32 * 1e12 if fp32 else float("inf")))) * 1e3 / args.steps,
                                            measured_ms_per_step=agg["ms_ccf"] / args.steps,
                                            frac=(agg["alignments"] * (4.0 * eng.lcirc * 3.0 / (tensor_peak * 1e12)
                                                  + 5.0 * eng.maxrin * np.log2(eng.maxrin) / (fp32 * 1e12 if fp32 else float("inf")))) / ccf_s,
                                            note="time floor of the kernel as it computes: the ring contraction's 4*lcirc flops three times "
                                                 "(split-bf16) at the measured dense bf16 peak, plus the inverse FFT's 5*maxrin*log2(maxrin) "
                                                 "flops at the FFMA peak of this run, over the measured kernel time"),
                    roofline_polar=dict(kernel="polar_group_kernel (Polar2Dm + Normalize_ring sums + Frngs)", bound="hbm",
                                        achieved=polar_gbs, peak=hbm_peak, unit="GB/s", frac=polar_gbs / hbm_peak,
                                        traffic=prof.get("polar_dram_bytes_per_launch"),
                                        peak_source="MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                                        share_of_step=polar_share,
                                        fp32=dict(achieved=polar_flops / (agg["ms_polar"] * 1e-3) / 1e12, peak=fp32, unit="TFLOP/s",
                                                  frac=(polar_flops / (agg["ms_polar"] * 1e-3) / 1e12 / fp32) if fp32 else None,
                                                  note="algorithmic flops (45 per sample + 2.5 L log2 L per ring, SURVEY 8d) over the FFMA peak of this run")),
                    stage_ms_per_step=dict(polar=agg["ms_polar"] / args.steps, ccf=agg["ms_ccf"] / args.steps,
                                           finalize=agg["ms_final"] / args.steps))
        if not args.no_cpu_baseline:
            nth = host_threads()
            ns = min(P, 2048)
            v, n, t = oracle_time_sample(host_images[:ns].numpy(), refs, cfg, nth)
            line["cpu_baseline"] = dict(value=v, unit="alignments/s", cores=nth, kind="port",
                                        sample="%d of %d particles x %d refs x %d shifts, %.1f s, oracle/cra_oracle.c (OpenMP)" % (n, P, R, S, t))
        print(json.dumps(line))
    eng.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
