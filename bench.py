#!/usr/bin/env python
"""Benchmark of the 2D multi-reference alignment hot path (BASELINE.json metric:
particle x reference x shift alignments/s, and s/iteration).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one iteration of the per-particle section of mref_ali2d (test_mref.py:170-223):
reference preparation, multiref_polar_ali_2d for every particle, rot_shift2D + even/odd class
sums, and (N > 1) one allreduce of the sums.  Workload at every N: BASELINE.json configs[1]
(100k synthetic 90x90 particles, 50 references, ou=36, xr=yr=3, ts=1, mirror on) PER GPU
(weak scaling).  `value` is timed with the particle stack resident in HBM; `e2e` runs the same
step through the C ABI from pinned host buffers (H2D of the stack, D2H of parameters and class
sums inside the timed region).  `--impl reference` times the CPU oracle port of the reference's
EMAN2 path (oracle/, all host threads) on a bounded sample of the same workload.
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

CFG = dict(P=100000, nx=90, R=50, ou=36, xr=3, yr=3, ts=1.0, nviews=64)
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


def oracle_prepare(images, refs, cfg):
    from oracle import oracle as o
    o.build()
    nx, ou = cfg["nx"], cfg["ou"]
    mask = o.model_circle(ou, nx)
    numr = o.numrinit(1, ou, 1)
    imgs = np.stack([o.normalize_mask(im, mask, 0) for im in images])
    _, cref = o.prepare_refs(refs, mask, numr)
    return o, imgs, cref, numr


def oracle_time_sample(images, refs, cfg, nthreads, target_s=15.0):
    """Time the CPU restatement of Util.multiref_polar_ali_2d on a bounded sample; returns
    (alignments/s, n_particles, seconds)."""
    o, imgs, cref, numr = oracle_prepare(images, refs, cfg)
    S = (2 * int(cfg["xr"] / cfg["ts"]) + 1) * (2 * int(cfg["yr"] / cfg["ts"]) + 1)
    cnx = cfg["nx"] // 2 + 1
    def run(n):
        centres = np.full((n, 2), float(cnx), np.float32)
        win = np.tile(np.array([cfg["xr"], cfg["xr"], cfg["yr"], cfg["yr"]], np.float32), (n, 1))
        t = time.perf_counter()
        o.align_batch(imgs[:n], cref, numr, centres, win, cfg["ts"], True, nthreads)
        return time.perf_counter() - t
    n0 = min(len(imgs), max(2 * nthreads, 8))
    t0 = run(n0)
    n = int(min(len(imgs), max(n0, n0 * target_s / max(t0, 1e-6))))
    t = run(n) if n > n0 else t0
    return n * S * cfg["R"] / t, n, t


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port; EMAN2 is not installable),
    all host threads, bounded sample per step."""
    if rank != 0:
        return
    from cryo_ralib_b200 import synth
    cfg = CFG
    nth = host_threads()
    nsample = 4096
    images, _ = synth.make_particles(nsample, cfg["nx"], cfg["nviews"], max_shift=cfg["xr"], seed=2025)
    refs = synth.initial_references(images, cfg["R"], seed=99)
    o, imgs, cref, numr = oracle_prepare(images, refs, cfg)
    S = 49
    cnx = cfg["nx"] // 2 + 1
    # size a step at ~6 s of wall time
    probe = min(nsample, 2 * nth)
    centres = np.full((nsample, 2), float(cnx), np.float32)
    win = np.full((nsample, 4), float(cfg["xr"]), np.float32)
    t = time.perf_counter(); o.align_batch(imgs[:probe], cref, numr, centres[:probe], win[:probe], cfg["ts"], True, nth)
    t_probe = time.perf_counter() - t
    n = int(max(probe, min(nsample, probe * 6.0 / max(t_probe, 1e-6))))
    times = []
    for it in range(args.warmup + args.steps):
        t = time.perf_counter()
        o.align_batch(imgs[:n], cref, numr, centres[:n], win[:n], cfg["ts"], True, nth)
        dt = time.perf_counter() - t
        if it >= args.warmup:
            times.append(dt)
    tot = sum(times)
    value = n * S * cfg["R"] * len(times) / tot
    sample = "%d of %d particles per step (x%d refs x%d shifts), %d host threads" % (n, cfg["P"], cfg["R"], S, nth)
    line = dict(impl="reference", metric=METRIC, value=value, unit="alignments/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * tot / len(times), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="mref 100k x 90x90, 50 refs, ou=36 xr=yr=3 ts=1 (BASELINE configs[1]); CPU oracle port on a bounded sample",
                            particles_per_step=n, refs=cfg["R"], shifts=S, nx=cfg["nx"], ou=cfg["ou"]),
                cpu_baseline=dict(value=value, unit="alignments/s", cores=nth, kind="port", sample=sample),
                e2e=dict(value=value, unit="alignments/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                s_per_iteration_extrapolated=cfg["P"] * S * cfg["R"] / value)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=CFG["P"], help="particles per GPU (default: BASELINE configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    from cryo_ralib_b200 import Engine, synth, alignment as al
    from cryo_ralib_b200.mref import TorchComm, LocalComm
    torch.cuda.set_device(local_rank)
    comm = LocalComm()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = TorchComm()
    cfg = dict(CFG); cfg["P"] = args.particles
    P, nx, R, ou, xr, yr, ts = cfg["P"], cfg["nx"], cfg["R"], cfg["ou"], cfg["xr"], cfg["yr"], cfg["ts"]
    W = max(args.warmup, 3)

    # synthetic particles, generated on the device (plumbing) -- seed differs per rank
    dev = "cuda:%d" % local_rank
    images_d, _ = synth.make_particles(P, nx, cfg["nviews"], max_shift=xr, seed=2025 + rank, device=dev)
    refs = synth.initial_references(images_d, R, seed=99).cpu().numpy()
    host_images = torch.empty((P, nx, nx), dtype=torch.float32, pin_memory=True)
    host_images.copy_(images_d)
    torch.cuda.synchronize()

    eng = Engine(nx, ou, xr, yr, ts=ts, max_particles=P, max_refs=R, normalize_ring=True, device=local_rank)
    eng.upload_particles_dev(images_d.data_ptr(), P, subtract_mask_mean=True)
    del images_d
    torch.cuda.empty_cache()
    fp32_peak = eng.measure_fp32_peak() if rank == 0 else (0.0, 0.0)
    stream = torch.cuda.ExternalStream(eng.L.cra_stream(eng.h), device=dev)
    goff = rank * P
    params0 = np.zeros((P, 4))

    # e2e: the stack is uploaded in chunks on the copy stream, each aligned as soon as it landed.  Chunks are whole
    # row batches of the engine (no partial launches in between); the first one is small so the alignment starts early.
    S = (2 * int(xr / ts) + 1) * (2 * int(yr / ts) + 1)
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

    def step(resident, params):
        """One iteration of the per-particle section; returns (new params, assign, stats)."""
        tr = [("start", time.perf_counter())]
        eng.set_refs(refs, normalize_mask=True)      # before the stack is queued: a copy from pageable memory waits behind it
        if not resident:
            base = host_images.data_ptr()
            for s, e in bounds:
                eng.upload_particles_async(base + s * nx * nx * 4, e - s, first=s, subtract_mask_mean=True)
            tr.append(("queued", time.perf_counter()))
        search, sxi, syi, params = al.mref_search_request(params, nx, ou, xr, yr)
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
        eng.accumulate(0, P, newp, res["iref"], goff)
        if world > 1:
            comm.allreduce_device(eng)
        if not resident:
            eng.get_sums()
        tr.append(("sums", time.perf_counter()))
        if trace:
            sys.stderr.write("trace %s: " % ("resident" if resident else "e2e") +
                             ", ".join("%s +%.1f" % (n, 1e3 * (t - tr[i][1])) for i, (n, t) in enumerate(tr[1:])) + "\n")
        return newp, res, st

    def timed(resident, nsteps, params):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        agg = dict(ms_polar=0.0, ms_ccf=0.0, ms_final=0.0, launches=0, alignments=0, rows=0, ccf_launches=0)
        for _ in range(nsteps):
            params, res, st = step(resident, params)
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
    p, _, _ = step(False, p)                      # one untimed pass through the host-buffer path
    ms_e2e, agg_e, p = timed(False, args.steps, p)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        aligns_step = agg["alignments"] / args.steps          # this rank; identical on all ranks
        value = world * aligns_step * args.steps / (ms_res * 1e-3)
        e2e = world * aligns_step * args.steps / (ms_e2e * 1e-3)
        fpa = flops_per_alignment(eng.lcirc, eng.maxrin)
        ccf_s = agg["ms_ccf"] * 1e-3
        ccf_tflops = agg["alignments"] * fpa / ccf_s / 1e12
        polar_gbs = agg["rows"] * eng.lcirc * 4.0 / (agg["ms_polar"] * 1e-3) / 1e9
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
        fp32 = max(fp32_peak)
        # a kernel timed inside a long step: the sustained figure
        tensor_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0))
        tensor_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else "fallback 1400 sustained (B200_PROFILING.md)"
        h2d = P * nx * nx * 4 + R * nx * nx * 4 + P * 24 + P * 20
        d2h = P * 32 + (R * 2 * nx * nx + R) * 4
        line = dict(metric=METRIC, value=value, unit="alignments/s", n_gpus=world, steps=args.steps, warmup=W,
                    ms_per_step=ms_res / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32", data="synthetic",
                    config=dict(workload="mref 100k synthetic 90x90 particles, 50 refs, ou=36 xr=yr=3 ts=1, mirror on (BASELINE configs[1]) per GPU",
                                particles_per_gpu=P, refs=R, nx=nx, ou=ou, xr=xr, yr=yr, ts=ts, shifts=49,
                                lcirc=eng.lcirc, maxrin=eng.maxrin, l2="inputs larger than L2 (3.2 GB stack + 2 GB spectra per batch); no flush"),
                    s_per_iteration=ms_res / args.steps * 1e-3,
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
                                  share_of_step=agg["ms_ccf"] / ms_res),
                    roofline_fp32_equiv=dict(kernel="ccf_tm_kernel", bound="fp32", achieved=ccf_tflops, peak=fp32, unit="TFLOP/s",
                                             frac=ccf_tflops / fp32 if fp32 else None,
                                             note="same algorithmic flops over the FFMA micro-benchmark measured in this run "
                                                  "(what an FP32 SIMT implementation could reach at best)"),
                    roofline_polar=dict(kernel="polar_fft_kernel (Polar2Dm + Normalize_ring + Frngs)", bound="hbm",
                                        achieved=polar_gbs, peak=hbm_peak, unit="GB/s", frac=polar_gbs / hbm_peak,
                                        traffic=prof.get("polar_dram_bytes_per_launch"),
                                        peak_source="MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                                        share_of_step=agg["ms_polar"] / ms_res),
                    stage_ms_per_step=dict(polar=agg["ms_polar"] / args.steps, ccf=agg["ms_ccf"] / args.steps,
                                           finalize=agg["ms_final"] / args.steps))
        if not args.no_cpu_baseline:
            nth = host_threads()
            ns = 2048
            v, n, t = oracle_time_sample(host_images[:ns].numpy(), refs, cfg, nth)
            line["cpu_baseline"] = dict(value=v, unit="alignments/s", cores=nth, kind="port",
                                        sample="%d of %d particles x %d refs x 49 shifts, %.1f s, oracle/cra_oracle.c (OpenMP)" % (n, P, R, t))
        print(json.dumps(line))
    eng.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
