"""The reference's OWN CUDA alignment library, timed on the same box ("beat this SIMT kernel
+ cuFFT on the same B200").

baseline/build_ref_cuda.sh compiles /root/reference/cuda/gpu_aln_{common,noref}.cu with the reference's
install.sh:25 command line plus an explicit sm_100 -gencode into baseline/_ref/gpu_aln_pack.so (git-ignored,
shipped to the GPU box).  This module drives that library through its stock entry points in the order the
reference's driver uses (test_mref_gpu_align.py:365-449):

    AlignConfig -> pre_align_size_check (largest batch, power-of-two search) -> pre_align_init ->
    reset_shifts -> per iteration: pre_align_fetch("ref_batch"), per batch: pre_align_fetch("sbj_batch")
    when the stack does not fit one batch, mref_align_run(start, stop)

None of this repository's kernels are on that path.  It is a THROUGHPUT bar, not a parity oracle: the library
is gpu_isac's arithmetic (fixed 256 samples per ring, bilinear texture fetch, ring weight r, no Normalize_ring,
integer shifts only), so its alignments differ from EMAN2's.  The worker runs in a child
process with a time limit because the library answers every CUDA error with exit(1) (gpu_aln_common.cu:89-103).

One deviation from the stock driver, stated in the JSON line: the particle batch is capped so that the CCF table
`258 * n * R * S * 2` stays below 2^31 elements.  The reference sizes batches for the memory of the card; on a
180 GB B200 that gives tables beyond 2^32 floats, past the `unsigned int` offsets of its kernels
(gpu_aln_noref.cu:1014-1015 `batch_table_row_offset`, `batch_table_mirror_offset`).
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "gpu_aln_pack.so")


class AlignConfig(C.Structure):                      # test_mref_gpu_align.py:112-123, gpu_aln_common.h:62-74
    _fields_ = [("sbj_num", C.c_uint), ("ref_num", C.c_uint), ("img_dim", C.c_uint), ("ring_num", C.c_uint),
                ("ring_len", C.c_uint), ("shift_step", C.c_float), ("shift_rng_x", C.c_float), ("shift_rng_y", C.c_float)]


class AlignParam(C.Structure):                       # test_mref_gpu_align.py:125-131
    _fields_ = [("sbj_id", C.c_int), ("ref_id", C.c_int), ("shift_x", C.c_float), ("shift_y", C.c_float),
                ("angle", C.c_float), ("mirror", C.c_bool)]


def build():
    if os.path.isdir("/root/reference/cuda"):
        subprocess.check_call([os.path.join(HERE, "build_ref_cuda.sh")], stdout=subprocess.DEVNULL)
    return SO if os.path.exists(SO) else None


def worker(cfg, nsample, steps, warmup):
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(HERE))
    from cryo_ralib_b200 import synth
    nx, R, ou, xr, ts = cfg["nx"], cfg["R"], cfg["ou"], cfg["xr"], cfg["ts"]
    S = (2 * int(xr / ts) + 1) ** 2
    images_d, _ = synth.make_particles(nsample, nx, min(cfg["nviews"], 64), max_shift=int(xr), seed=2025, device="cuda:0")
    refs = synth.initial_references(images_d, R, seed=99).cpu().numpy().astype(np.float32)
    images = images_d.cpu().numpy().astype(np.float32)
    del images_d
    torch.cuda.empty_cache()
    L = C.CDLL(SO)
    L.pre_align_init.restype = C.c_ulonglong
    L.mref_align_run.restype = C.c_ulonglong
    L.pre_align_size_check.restype = C.c_bool
    fp = C.POINTER(C.c_float)

    def ptrs(a):
        arr = (fp * a.shape[0])()
        for i in range(a.shape[0]):
            arr[i] = a[i].ctypes.data_as(fp)
        return arr

    acfg = AlignConfig(nsample, R, nx, ou, 256, ts, xr, xr)
    limit = 0
    import math
    for split in [2 ** i for i in range(int(math.log(nsample, 2)) + 1)][::-1]:          # test_mref_gpu_align.py:374-378
        acfg.sbj_num = min(limit + split, nsample)
        if L.pre_align_size_check(C.c_uint(nsample), C.byref(acfg), C.c_uint(0), C.c_float(0.9), C.c_bool(False)):
            limit += split
    by_memory = min(limit, nsample)
    cap = int((2 ** 31 - 1) // (258 * R * S * 2))
    batch = max(1, min(by_memory, cap))
    acfg.sbj_num = batch
    L.pre_align_init(C.c_uint(nsample), C.byref(acfg), C.c_uint(0))
    nbatch = (nsample + batch - 1) // batch
    sbj_ptrs = ptrs(images)
    ref_ptrs = ptrs(refs)
    if nbatch == 1:
        L.pre_align_fetch(sbj_ptrs, C.c_uint(nsample), C.c_char_p(b"sbj_batch"))
    L.reset_shifts(C.c_float(xr), C.c_float(ts))
    times = []
    for it in range(warmup + steps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        L.pre_align_fetch(ref_ptrs, C.c_int(R), C.c_char_p(b"ref_batch"))
        for b in range(nbatch):
            s, e = b * batch, min(nsample, (b + 1) * batch)
            if nbatch > 1:
                sub = (fp * (e - s))(*[sbj_ptrs[i] for i in range(s, e)])
                L.pre_align_fetch(sub, C.c_int(e - s), C.c_char_p(b"sbj_batch"))
            L.mref_align_run(C.c_int(s), C.c_int(e))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        if it >= warmup:
            times.append(dt)
    print("REFCUDA " + json.dumps(dict(times=times, batch=batch, by_memory=by_memory, nbatch=nbatch, nsample=nsample, S=S)))
    sys.stdout.flush()
    os._exit(0)                                           # the library's static state is not worth a clean teardown


def bench_line(cfg, args, metric):
    if not os.path.exists(SO):
        return dict(impl="reference-gpu", unavailable="baseline/_ref/gpu_aln_pack.so is missing (baseline/build_ref_cuda.sh builds it where /root/reference exists)")
    nsample = args.particles or 16384
    S = (2 * int(cfg["xr"] / cfg["ts"]) + 1) ** 2
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", json.dumps(cfg), str(nsample),
                              str(args.steps), str(args.warmup)], capture_output=True, text=True, timeout=900)
    except subprocess.TimeoutExpired:
        return dict(impl="reference-gpu", unavailable="the reference library did not finish within 900 s")
    rec = [ln for ln in out.stdout.splitlines() if ln.startswith("REFCUDA ")]
    if out.returncode != 0 or not rec:
        return dict(impl="reference-gpu", unavailable="the reference library failed (rc %d): %s" % (out.returncode, (out.stdout + out.stderr)[-300:].replace("\n", " | ")))
    r = json.loads(rec[-1][8:])
    tot = sum(r["times"])
    value = nsample * cfg["R"] * S * len(r["times"]) / tot
    return dict(impl="reference-gpu", metric=metric, value=value, unit="alignments/s", n_gpus=1, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * tot / len(r["times"]), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=cfg["name"] + "; the reference's own CUDA (gpu_aln_pack.so built for sm_100) on a bounded sample, "
                            "stock entry points pre_align_size_check -> pre_align_init -> pre_align_fetch -> mref_align_run",
                            particles_per_step=nsample, refs=cfg["R"], shifts=S, nx=cfg["nx"], ring_num=cfg["ou"], ring_len=256,
                            batch=r["batch"], batch_by_memory=r["by_memory"], batches=r["nbatch"],
                            note="gpu_isac arithmetic (fixed 256-sample rings, bilinear texture, no Normalize_ring, integer shifts): "
                                 "a throughput bar, not EMAN2 parity; batch capped where the CCF table stays below 2^31 elements "
                                 "(32-bit table offsets in the reference's kernels)"),
                e2e=dict(value=value, unit="alignments/s", h2d_bytes_per_step=int(cfg["R"] * cfg["nx"] ** 2 * 4 + (nsample * cfg["nx"] ** 2 * 4 if r["nbatch"] > 1 else 0)),
                         d2h_bytes_per_step=0))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        worker(json.loads(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    else:
        print(build())
