"""Drop-in check at the behaviour level: ONE ctypes harness, written the way the reference's driver drives its library
(test_mref_gpu_align.py:365-449: AlignConfig -> pre_align_init -> pre_align_fetch x2 -> reset_shifts -> mref_align_run ->
read AlignParam[]), run once with the reference's own CUDA library (oracle/_ref/gpu_aln_pack.so, built unchanged for sm_100 by
`make -C oracle ref`) and once with this repository's libcryo_ralib.so -- a one-line change of the CDLL path
 -- on the same synthetic stack with known poses.

It is NOT a parity test: the reference library is gpu_isac's arithmetic (256 bilinear samples per ring, ring weight r, no
Normalize_ring, integer shifts) while this engine follows EMAN2's.  What it shows is that both
libraries answer the same question through the same ABI with the same conventions: class (reference) chosen, mirror flag,
in-plane angle and the accumulated shift of AlignParam, each against the ground truth and against each other.

usage (GPU box): python scripts/compare_ref_cuda.py [P] [views] [snr] [mref|mref_m|reffree]
mref_m: mref_align_run_m instead (test_mref_cheng_yu_bdb_cuda.py:546-556): additionally the [2R][nx][nx] even / odd class
sums and get_num_ref of the two libraries (class sizes; correlation of the sums and of even - odd per class).
reffree: the reference-free entry point pre_align_run(0, P) instead (test_reffree.py:292-426), with ONE reference.  The
first average of a real reference-free run is a featureless blob (flat correlation landscape: the two arithmetics then
disagree on ties), so the check uses a stack of one view and its noise-free projection as the reference.
Each library runs in its own child process (same symbol names; the reference answers CUDA errors with exit(1))."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_SO = os.path.join(ROOT, "oracle", "_ref", "gpu_aln_pack.so")                 # make -C oracle ref
if not os.path.exists(REF_SO):
    REF_SO = os.path.join(ROOT, "baseline", "_ref", "gpu_aln_pack.so")          # the same build (baseline/build_ref_cuda.sh)
OUR_SO = os.path.join(ROOT, "cryo_ralib_b200", "libcryo_ralib.so")
NX, OU, XR = 90, 36, 3


class AlignConfig(C.Structure):                      # test_mref_gpu_align.py:112-123
    _fields_ = [("sbj_num", C.c_uint), ("ref_num", C.c_uint), ("img_dim", C.c_uint), ("ring_num", C.c_uint),
                ("ring_len", C.c_uint), ("shift_step", C.c_float), ("shift_rng_x", C.c_float), ("shift_rng_y", C.c_float)]


class AlignParam(C.Structure):                       # test_mref_gpu_align.py:125-131
    _fields_ = [("sbj_id", C.c_int), ("ref_id", C.c_int), ("shift_x", C.c_float), ("shift_y", C.c_float),
                ("angle", C.c_float), ("mirror", C.c_bool)]


def worker(so, data, out, mode="mref"):
    d = np.load(data)
    images, refs = np.ascontiguousarray(d["images"], np.float32), np.ascontiguousarray(d["refs"], np.float32)
    P, R = images.shape[0], refs.shape[0]
    L = C.CDLL(so)
    L.pre_align_init.restype = C.c_ulonglong
    L.mref_align_run.restype = C.c_ulonglong
    fp = C.POINTER(C.c_float)
    ptrs = lambda a: (fp * a.shape[0])(*[a[i].ctypes.data_as(fp) for i in range(a.shape[0])])
    cfg = AlignConfig(P, R, NX, OU, 256, 1.0, float(XR), float(XR))
    par = C.cast(C.c_void_p(L.pre_align_init(C.c_uint(P), C.byref(cfg), C.c_uint(0))), C.POINTER(AlignParam))
    L.pre_align_fetch(ptrs(images), C.c_uint(P), C.c_char_p(b"sbj_batch"))
    L.pre_align_fetch(ptrs(refs), C.c_int(R), C.c_char_p(b"ref_batch"))
    L.reset_shifts(C.c_float(XR), C.c_float(1.0))
    rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
    if mode == "reffree":
        L.pre_align_run(C.c_int(0), C.c_int(P))
    elif mode == "mref_m":
        # test_mref_cheng_yu_bdb_cuda.py:546-556: managed [2R][nx][nx] -- R even sums, then R odd sums -- and the class sizes
        L.mref_align_run_m.restype = C.POINTER(C.c_float)
        L.get_num_ref.restype = C.POINTER(C.c_int)
        sp = L.mref_align_run_m(C.c_int(0), C.c_int(P))
        rt.cudaDeviceSynchronize()
        sums = np.ctypeslib.as_array(sp, shape=(2 * R, NX, NX)).copy()
        np.save(out.replace(".npy", "_sums.npy"), sums)
        np.save(out.replace(".npy", "_counts.npy"), np.ctypeslib.as_array(L.get_num_ref(), shape=(R,)).copy())
    else:
        L.mref_align_run(C.c_int(0), C.c_int(P))
    rt.cudaDeviceSynchronize()
    res = np.array([(par[i].ref_id, par[i].shift_x, par[i].shift_y, par[i].angle, int(par[i].mirror)) for i in range(P)], np.float64)
    np.save(out, res)
    sys.stdout.flush()
    os._exit(0)


def circ(a, b):
    return np.abs((a - b + 180.0) % 360.0 - 180.0)


def make_inputs(P, V, snr):
    """-> (images, refs, truth): the stack and the references exactly as they are handed to pre_align_fetch."""
    from cryo_ralib_b200 import synth
    images, truth = synth.make_particles(P, NX, V, max_shift=XR, snr=snr, seed=31)
    # references: the noise-free views at psi = 0, no shift, no mirror (what a converged class average looks like)
    pos, sigma, amp = synth.make_density(NX)
    rots = synth.random_rotations(V, np.random.default_rng(31))
    proj = np.einsum("vij,bj->vbi", rots, pos)[:, :, :2]
    refs = synth.render(NX, proj[:, :, 0], proj[:, :, 1], sigma, amp).astype(np.float32)
    # what the driver does on the host before it fetches (test_mref_gpu_align.py: normalize.mask with model_circle(ou)):
    # particles x - mean_mask, references (x - mean_mask) / sigma_mask
    yy, xx = np.mgrid[0:NX, 0:NX]
    mask = ((xx - NX // 2) ** 2 + (yy - NX // 2) ** 2) <= OU * OU
    images = (images - images[:, mask].mean(axis=1)[:, None, None]).astype(np.float32)
    refs = ((refs - refs[:, mask].mean(axis=1)[:, None, None]) / refs[:, mask].std(axis=1, ddof=1)[:, None, None]).astype(np.float32)
    return images, refs, truth


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    V = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    snr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    mode = sys.argv[4] if len(sys.argv) > 4 else "mref"
    if mode == "reffree":
        V = 1
    tmp = os.path.join(ROOT, "gpurun_out", "cmp")
    os.makedirs(tmp, exist_ok=True)
    images, refs, truth = make_inputs(P, V, snr)
    if mode == "reffree":
        refs = refs[:1]                                                   # one reference (run with views = 1)
    np.savez(os.path.join(tmp, "data.npz"), images=images, refs=refs)
    res = {}
    for tag, so in (("reference", REF_SO), ("this", OUR_SO)):
        if not os.path.exists(so):
            print("%s library missing: %s" % (tag, so)); return 1
        out = os.path.join(tmp, tag + ".npy")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", so, os.path.join(tmp, "data.npz"), out, mode],
                           capture_output=True, text=True, timeout=600)
        if r.returncode != 0 or not os.path.exists(out):
            print("%s library failed (rc %d): %s" % (tag, r.returncode, (r.stdout + r.stderr)[-400:])); return 1
        res[tag] = np.load(out)
    a, b = res["reference"], res["this"]
    rep = dict(mode=mode, particles=P, views=V, snr=snr, nx=NX, ou=OU, xr=XR)
    for tag, x in (("reference", a), ("this", b)):
        rep[tag] = dict(view_recovered=float((x[:, 0] == truth["view"]).mean()), mirror_recovered=float((x[:, 4] == truth["mirror"]).mean()))
    same_cls = a[:, 0] == b[:, 0]
    both = same_cls & (a[:, 4] == b[:, 4])
    da = circ(a[both, 3], b[both, 3])
    ds = np.hypot(a[both, 1] - b[both, 1], a[both, 2] - b[both, 2])
    rep["between"] = dict(same_class=float(same_cls.mean()), same_class_and_mirror=float(both.mean()),
                          angle_diff_deg=dict(median=float(np.median(da)), p90=float(np.percentile(da, 90)), p99=float(np.percentile(da, 99))),
                          shift_diff_px=dict(median=float(np.median(ds)), p90=float(np.percentile(ds, 90)), within_1px=float((ds <= 1.0).mean())))
    if mode == "mref_m":
        ncc = lambda x, y: float((x * y).sum() / np.sqrt((x * x).sum() * (y * y).sum()))
        sa, sb = np.load(os.path.join(tmp, "reference_sums.npy")), np.load(os.path.join(tmp, "this_sums.npy"))
        ca, cb = np.load(os.path.join(tmp, "reference_counts.npy")), np.load(os.path.join(tmp, "this_counts.npy"))
        R = sa.shape[0] // 2
        same_par = [ncc(sa[i], sb[i]) for i in range(2 * R)]
        # even - odd cancels the signal of a class: what is left is the noise of its members, which correlates between the
        # libraries only if both put the same particles into the same half
        dif = [ncc(sa[i] - sa[R + i], sb[i] - sb[R + i]) for i in range(R)]
        rep["sums"] = dict(class_sizes_equal=bool(np.array_equal(ca, cb)), class_sizes=ca.tolist(),
                           ncc_same_half=dict(min=min(same_par), median=float(np.median(same_par))),
                           ncc_even_minus_odd=dict(min=min(dif), median=float(np.median(dif))))
        for f in ("reference_sums.npy", "this_sums.npy", "reference_counts.npy", "this_counts.npy"):
            os.remove(os.path.join(tmp, f))
    print(json.dumps(rep, indent=1))
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "compare_ref_cuda_%s_P%d_V%d_snr%g.json" % (mode, P, V, snr)), "w"), indent=1)
    # the reference library's own answers (ref_id, shift_x, shift_y, angle, mirror per particle): small, kept -- the golden
    # fixture tests/golden/refcuda_mref_outputs.npz is one of these (tests/golden/make_refcuda_case.py)
    np.save(os.path.join(ROOT, "gpurun_out", "refcuda_outputs_%s_P%d_V%d_snr%g.npy" % (mode, P, V, snr)), a)
    for f in ("data.npz", "reference.npy", "this.npy"):       # scratch (the stack is 130 MB at 4096 particles)
        os.remove(os.path.join(tmp, f))
    return 0


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        worker(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "mref")
    else:
        sys.exit(main())
