"""Host-side logic of the product (no GPU): parameter algebra, search windows and the
reference update against the oracle; the C-ABI library loads and exports every declared symbol."""
import copy
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from cryo_ralib_b200 import build
    so = build.build()
    L = C.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", "cryo_ralib.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", hdr)) - {"defined", "sizeof"}
    assert {"cra_create", "cra_align", "cra_accumulate", "pre_align_init", "mref_align_run", "gpu_clear"} <= names
    for n in sorted(names):
        assert hasattr(L, n), "libcryo_ralib.so does not export %s" % n


def test_struct_layouts_match_reference_ctypes_mirrors():
    from cryo_ralib_b200.lib import AlignConfig, AlignParam, CraSearch, CraResult, SEARCH_DTYPE, RESULT_DTYPE
    assert C.sizeof(AlignConfig) == 32          # 5 x u32 + 3 x f32 (gpu_aln_common.h:62-75)
    assert C.sizeof(AlignParam) == 24           # 2 x i32 + 3 x f32 + bool, padded (gpu_aln_common.h:76-83)
    assert AlignParam.mirror.offset == 20
    assert C.sizeof(CraSearch) == SEARCH_DTYPE.itemsize == 24
    assert C.sizeof(CraResult) == RESULT_DTYPE.itemsize == 32


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cryo_ralib_b200 import Engine
    with pytest.raises(RuntimeError):
        Engine(90, 36, 3, max_particles=4, max_refs=2)


def test_missing_library_raises_not_falls_back(tmp_path):
    from cryo_ralib_b200 import lib
    with pytest.raises(lib.LibraryMissing):
        lib.load_library(str(tmp_path / "nope.so"))


def test_numrinit_ringwe_match_oracle(oracle):
    from cryo_ralib_b200 import alignment as al
    for ir, ou, rs in ((1, 36, 1), (1, 60, 1), (1, 29, 1), (2, 40, 2), (1, 5, 1), (3, 100, 3)):
        assert np.array_equal(al.numrinit(ir, ou, rs), oracle.numrinit(ir, ou, rs))
        assert np.array_equal(al.ringwe(al.numrinit(ir, ou, rs)), oracle.ringwe(oracle.numrinit(ir, ou, rs)))


def test_transform_algebra_matches_oracle_and_golden(oracle):
    from cryo_ralib_b200 import alignment as al
    a, sx, sy, m = al.combine_params2(np.array([141.42257753434927]), 0.47458410263061523, 4.216013431549072,
                                      np.array([0]), 0.0, -3, -1, 0)
    assert (a[0], sx[0], sy[0], m[0]) == (141.42257753434927, -2.5254158973693848, 3.2160134315490723, 0)
    a, sx, sy, m = al.inverse_transform2(np.array([141.42257753434927]), -2.5254158973693848, 3.2160134315490723)
    assert (a[0], sx[0], sy[0]) == (218.5774203360948, 0.031129680573940277, 4.0889482498168945)
    rng = np.random.default_rng(0)
    n = 500
    a1, a2 = rng.uniform(0, 360, (2, n)); s = rng.uniform(-6, 6, (4, n)); m1, m2 = rng.integers(0, 2, (2, n))
    got = al.combine_params2(a1, s[0], s[1], m1, a2, s[2], s[3], m2)
    inv = al.inverse_transform2(a1, s[0], s[1], m1)
    for i in range(n):
        w = oracle.combine_params2(a1[i], s[0][i], s[1][i], m1[i], a2[i], s[2][i], s[3][i], m2[i])
        # libm vs numpy atan2 may differ in the last ulp of the angle; translations are float32-exact
        assert abs(got[0][i] - w[0]) < 1e-11 and (got[1][i], got[2][i], got[3][i]) == w[1:]
        w = oracle.inverse_transform2(a1[i], s[0][i], s[1][i], m1[i])
        assert abs(inv[0][i] - w[0]) < 1e-11 and (inv[1][i], inv[2][i], inv[3][i]) == w[1:]


def test_search_request_matches_reference_loop(oracle):
    """test_mref.py:184-198 particle by particle, including the mashi reset and ragged windows."""
    from cryo_ralib_b200 import alignment as al
    rng = np.random.default_rng(1)
    n = 300
    params = np.stack([rng.uniform(0, 360, n), rng.uniform(-9, 9, n), rng.uniform(-9, 9, n), rng.integers(0, 2, n)], 1)
    s, sxi, syi, pr = al.mref_search_request(params, 90, 36, 3, 2)
    for i in range(n):
        _, x, y, _ = oracle.inverse_transform2(params[i, 0], params[i, 1], params[i, 2])
        if abs(x) > 8 or abs(y) > 8:
            x = y = 0.0
            assert np.all(pr[i] == 0)
        tx = oracle.search_range(90, 36, x, 3); ty = oracle.search_range(90, 36, y, 2)
        assert (s["xl"][i], s["xr"][i]) == (np.float32(tx[0]), np.float32(tx[1]))
        assert (s["yl"][i], s["yr"][i]) == (np.float32(ty[0]), np.float32(ty[1]))
        assert s["cx"][i] == np.float32(46 + x) and s["cy"][i] == np.float32(46 + y)
    # composing a zero result returns the inverse of the search offset
    from cryo_ralib_b200.lib import RESULT_DTYPE
    res = np.zeros(n, RESULT_DTYPE)
    newp = al.compose_result(sxi, syi, res)
    assert np.allclose(newp[:, 1], -sxi, atol=1e-6) and np.allclose(newp[:, 2], -syi, atol=1e-6)


def test_native_bookkeeping_equals_numpy_restatement():
    """csrc/cra_host.cu (what the drivers call) against the numpy Transform algebra the golden values pin:
    search requests bit for bit, composed parameters bit for bit except the last ulp of the angle."""
    from cryo_ralib_b200 import alignment as al
    from cryo_ralib_b200.lib import RESULT_DTYPE
    rng = np.random.default_rng(5)
    n = 20000
    params = np.stack([rng.uniform(0, 360, n), rng.uniform(-9, 9, n), rng.uniform(-9, 9, n),
                       rng.integers(0, 2, n).astype(float)], 1)
    params[:50] = 0.0
    a = al.mref_search_request(params, 90, 36, 3, 2, native=True)
    b = al.mref_search_request(params, 90, 36, 3, 2, native=False)
    for f in a[0].dtype.names:
        assert np.array_equal(a[0][f], b[0][f]), f
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    res = np.zeros(n, RESULT_DTYPE)
    res["ang"] = rng.uniform(0, 360, n); res["sxs"] = rng.uniform(-4, 4, n); res["sys"] = rng.uniform(-4, 4, n)
    res["mirror"] = rng.integers(0, 2, n)
    c = al.compose_result(a[1], a[2], res, native=True)
    d = al.compose_result(a[1], a[2], res, native=False)
    assert np.array_equal(c[:, 1:], d[:, 1:])
    assert np.abs(c[:, 0] - d[:, 0]).max() < 1e-12
    # the reference-free prologue (centre shift folded in, shifts clamped instead of reset)
    for cs in ((0.0, 0.0), (0.37, -1.21)):
        e = al.reffree_search_request(params, cs, 90, 36, 3, 2, native=True)
        f = al.reffree_search_request(params, cs, 90, 36, 3, 2, native=False)
        for name in e[0].dtype.names:
            assert np.array_equal(e[0][name], f[0][name]), (cs, name)
        assert np.array_equal(e[1], f[1]) and np.array_equal(e[2], f[2])
        assert np.abs(e[1]).max() <= 8 and np.abs(e[2]).max() <= 8          # mashi = 46 - 36 - 2


def test_mpi_start_end_partitions():
    from cryo_ralib_b200 import alignment as al
    for n, p in ((10000, 8), (100001, 3), (7, 8), (50, 1)):
        b = [al.mpi_start_end(n, p, i) for i in range(p)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(p - 1))


def test_reference_update_matches_oracle(oracle):
    from cryo_ralib_b200 import refupdate as ru, synth
    images, _ = synth.make_particles(80, 64, 3, max_shift=1, snr=1.0, seed=12)
    mask = ru.model_circle(28, 64)
    assert np.array_equal(mask, oracle.model_circle(28, 64))
    R = 4
    sums = np.zeros((R, 2, 64, 64), np.float32); counts = np.zeros(R)
    for i, im in enumerate(images):
        r = i % 3                                  # class 3 stays empty -> reseeded
        sums[r, i % 2] += im; counts[r] += 1
    masked = np.stack([oracle.normalize_mask(im, mask, 0) for im in images])
    import random
    want, winfo = oracle.update_refs(sums.copy(), counts.copy(), masked, mask, 1, random.Random(1000))
    got, ginfo = ru.update_refs(sums.copy(), counts.copy(), mask, 1, ru.make_reseeder(1000, 80, lambda k: masked[k]))
    assert ginfo["reseeded"] == [3] and list(winfo["reseeded"].keys()) == [3]
    assert np.allclose(ginfo["filter"], winfo["filter"], rtol=1e-9)
    assert np.allclose(np.array(ginfo["cs"]), np.array(winfo["cs"]), atol=1e-9)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    f = ru.fsc(sums[0, 0], sums[0, 1]); g = oracle.fsc(sums[0, 0], sums[0, 1])
    assert f[0] == g[0] and f[2] == g[2] and np.allclose(f[1], g[1], atol=1e-7)
    assert ru.fit_tanh(copy.deepcopy(f)) == pytest.approx(oracle.fit_tanh(copy.deepcopy(g)), rel=1e-9)


def test_native_fit_tanh_follows_the_python_simplex():
    """cra_fit_tanh (csrc/cra_host.cu) against refupdate.fit_tanh's interpreted amoeba (sp_filter.fit_tanh ->
    sp_utilities.amoeba) on random FSC curves, and the oracle's own restatement on one of them."""
    from cryo_ralib_b200 import refupdate as ru
    from oracle import oracle as o
    rng = np.random.default_rng(3)
    freq = np.arange(46) / 90.0
    for trial in range(40):
        fc = rng.uniform(0.08, 0.35); w = rng.uniform(0.02, 0.15)
        val = np.clip(1 / (1 + np.exp((freq - fc) / w * 4)) + rng.normal(0, 0.03, 46), -0.2, 0.999)
        fr = [list(freq), list(val), [1] * 46]
        a = ru.fit_tanh(fr, native=True); b = ru.fit_tanh(fr, native=False)
        assert abs(a[0] - b[0]) <= 1e-9 and abs(a[1] - b[1]) <= 1e-9, (trial, a, b)
    c = o.fit_tanh(fr)
    assert abs(a[0] - c[0]) <= 1e-6 and abs(a[1] - c[1]) <= 1e-6


def test_update_refs_fits_the_shared_fsc_once(monkeypatch):
    """Every class gets the same class-averaged FSC (test_mref.py:258-276): one fit per update, same references."""
    from cryo_ralib_b200 import refupdate as ru, synth
    images, _ = synth.make_particles(60, 64, 8, seed=4)
    R = 5
    sums = np.zeros((R, 2, 64, 64), np.float32); counts = np.zeros(R, np.float32)
    for i in range(60):
        sums[i % R, i % 2] += images[i]; counts[i % R] += 1
    mask = ru.model_circle(28, 64)
    calls = []
    real = ru.fit_tanh
    monkeypatch.setattr(ru, "fit_tanh", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    refs, info = ru.update_refs(sums, counts, mask, center=1)
    assert len(calls) == 1
    per_class = np.stack([ru.normalize_mask(ru.ref_ali2d((sums[j, 0] + sums[j, 1]) * np.float32(1.0 / counts[j]), info["frsc"], 1)[0], mask, 1)
                          for j in range(R)])
    assert np.array_equal(per_class, refs)


def test_params_files_round_trip(tmp_path):
    """params.txt ('idx angle sx sy mirror class', src/utils_ralib.py:31-32) and initial2Dparams.txt rows."""
    from cryo_ralib_b200 import stackio
    rng = np.random.default_rng(2)
    p = np.stack([rng.uniform(0, 360, 9), rng.uniform(-4, 4, 9), rng.uniform(-4, 4, 9), rng.integers(0, 2, 9)], 1)
    c = rng.integers(0, 5, 9)
    f = str(tmp_path / "params.txt")
    stackio.write_params(f, p, c, first_index=3)
    q, d = stackio.read_params(f)
    assert np.allclose(q, p, atol=1e-6) and np.array_equal(d, c)
    g = str(tmp_path / "initial2Dparams.txt")
    with open(g, "w") as fh:
        for r in p:
            fh.write("%14.6f %14.6f %14.6f %d\n" % (r[0], r[1], r[2], int(r[3])))
    q, d = stackio.read_params(g)
    assert np.allclose(q, p, atol=1e-6) and d is None
