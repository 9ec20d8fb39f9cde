"""Boxes beyond what fits into shared memory (through the C ABI, against the CPU oracle).

The image tile of the row kernels, of rot_shift2D and the three spectra of the reference update live in shared memory
for the named configurations (90 and 128 pixels).  Larger boxes take the same kernels with the image taps read from
global memory (general row kernel: boxes beyond ~170 pixels; rot_shift2D: beyond ~238) and the transforms of the
reference update in a global scratch buffer (beyond ~136 pixels); maxrin = 1024 (ou > 81) takes the shared-memory
CCF kernel.  Where the ring set is small against the frame the grouped row kernel stages a WINDOW of the image around
the particle's search window instead of the whole image (any frame size).  Same bars as the named geometries
(tests/test_gpu_configs.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PEAK_RTOL = 1e-4
TIE_BAND = 2e-5


def _setup(oracle, nx, ou, P, R):
    from cryo_ralib_b200 import synth
    allp, _ = synth.make_particles(P + 6 * R, nx, 16, max_shift=2, seed=11)
    images = np.ascontiguousarray(allp[:P])
    refs = synth.initial_references(allp[P:], R, per_ref=6, seed=5)
    mask = oracle.model_circle(ou, nx)
    numr = oracle.numrinit(1, ou, 1)
    imgs = np.stack([oracle.normalize_mask(im, mask, 0) for im in images])
    return images, refs, mask, numr, imgs


@pytest.mark.parametrize("nx,ou,xr", [(160, 72, 2), (192, 80, 2), (256, 100, 1), (256, 60, 3), (128, 40, 3)])
def test_large_box_alignment(oracle, nx, ou, xr):
    """Two iterations (zero parameters, then the composed ones: fractional centres, ragged windows)."""
    from cryo_ralib_b200 import Engine, alignment as al
    P, R = 12, 5
    images, refs, mask, numr, imgs = _setup(oracle, nx, ou, P, R)
    maxrin = int(numr[-1])
    _, cref = oracle.prepare_refs(refs, mask, numr)
    e = Engine(nx, ou, xr, ts=1.0, max_particles=P, max_refs=R, normalize_ring=True)
    e.upload_particles(images); e.set_refs(refs)
    for j in range(R):                                           # the reference spectra themselves
        got, want = e.ref_spectrum(j), cref[j]
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max(), (nx, j)
    params = np.zeros((P, 4))
    for it in range(2):
        search, sxi, syi, params = al.mref_search_request(params, nx, ou, xr, xr)
        res = e.align(0, P, search)
        centres = np.stack([search["cx"], search["cy"]], 1)
        win = np.stack([search["xl"], search["xr"], search["yl"], search["yr"]], 1)
        want = oracle.align_batch(imgs, cref, numr, centres, win, 1.0, True, nthreads=oracle.max_threads())
        same = (res["iref"] == want[:, 4].astype(int)) & (res["mirror"] == want[:, 3].astype(int)) \
            & (res["sx"] == want[:, 6]) & (res["sy"] == want[:, 7])
        rel = np.abs(res["peak"] - want[:, 5]) / np.maximum(np.abs(want[:, 5]), 1e-30)
        dang = np.abs((res["ang"] - want[:, 0] + 180.0) % 360.0 - 180.0)
        assert not ((~same) & (rel >= TIE_BAND)).any(), (nx, it, np.where(~same)[0], rel[~same])
        assert rel[same].max() <= PEAK_RTOL, (nx, it, rel[same].max())
        assert (dang[same] <= 0.5 * 360.0 / maxrin + 1e-3).sum() >= same.sum() - 1, (nx, it, dang)
        params = al.compose_result(sxi, syi, res)
    e.close()


@pytest.mark.parametrize("nx,ou", [(192, 80), (256, 100)])
def test_large_box_class_sums_and_reference_update(oracle, nx, ou):
    """rot_shift2D + even/odd sums, then the device reference update (FSC, tangent filter, centring) on them."""
    from cryo_ralib_b200 import Engine
    P, R = 24, 3
    images, refs, mask, numr, imgs = _setup(oracle, nx, ou, P, R)
    rng = np.random.default_rng(5)
    params = np.stack([rng.uniform(0, 360, P), rng.uniform(-5, 5, P), rng.uniform(-5, 5, P), rng.integers(0, 2, P)], 1)
    params[0] = [0, 0, 0, 0]; params[1] = [90, 2, -3, 1]
    assign = (np.arange(P) % R).astype(np.int32)
    e = Engine(nx, ou, 1, ts=1.0, max_particles=P, max_refs=R, normalize_ring=True)
    e.upload_particles(images); e.set_refs(refs)
    e.zero_sums(); e.accumulate(0, P, params, assign, global_offset=3)
    sums, counts = e.get_sums()
    want = np.zeros((R, 2, nx, nx), np.float64)
    for i in range(P):
        want[assign[i], (3 + i) % 2] += oracle.rot_shift2d(imgs[i], *params[i])
    assert np.array_equal(counts[:R], np.bincount(assign, minlength=R))
    assert np.abs(sums[:R] - want).max() <= 1e-5 * np.abs(want).max()
    wrefs, winfo = oracle.update_refs(sums[:R], counts[:R].astype(np.float64), imgs, mask, 1, None)
    got, ginfo = e.update_refs_device(center=1, reseed=None)
    for j, cur in ginfo["class_fsc"].items():
        w = oracle.fsc(sums[j, 0], sums[j, 1], 1.0)
        assert np.abs(np.array(cur[1]) - np.array(w[1])).max() <= 2e-5, (nx, j)
    assert np.allclose(ginfo["filter"], winfo["filter"], rtol=2e-4), (ginfo["filter"], winfo["filter"])
    err = np.abs(got - wrefs).max() / np.abs(wrefs).max()
    assert err <= 5e-5, (nx, err)
    e.close()


@pytest.mark.parametrize("nx,ou,ts", [(128, 40, 1.0), (128, 40, 0.5), (256, 60, 1.0), (98, 30, 1.0)])
def test_windowed_tile_row_spectra_match_oracle(oracle, nx, ou, ts):
    """The grouped row kernel with a WINDOWED image tile (ring set small against the frame, or a frame too large for
    shared memory): every shift row against the oracle's Polar2Dm + Normalize_ring + Frngs; centres at the extremes the
    windowed tile admits (its origin is clamped to the frame there), integer / half-integer centres (pixel-boundary
    samples), ragged windows.  nx = 98 has lines that are not 16-byte granular: the tile is loaded cooperatively."""
    import os
    from cryo_ralib_b200 import Engine, synth
    from cryo_ralib_b200.lib import SEARCH_DTYPE
    if os.environ.get("CRA_CCF") == "simt" or os.environ.get("CRA_POLAR") == "general":
        pytest.skip("diagnostic switch selects the general row kernel")
    P, xr = 6, 3.0 if ts >= 1 else 1.5
    allp, _ = synth.make_particles(P + 8, nx, 16, max_shift=2, seed=21)
    images = np.ascontiguousarray(allp[:P]); refs = synth.initial_references(allp[P:], 2, per_ref=4, seed=5)
    mask = oracle.model_circle(ou, nx)
    numr = oracle.numrinit(1, ou, 1)
    imgs = np.stack([oracle.normalize_mask(im, mask, 0) for im in images])
    e = Engine(nx, ou, xr, ts=ts, max_particles=P, max_refs=2, normalize_ring=True)
    e.upload_particles(images, subtract_mask_mean=True); e.set_refs(refs)
    c = nx // 2 + 1
    lo, hi = ou + 2 + xr, nx - 1 - ou - xr                       # centres whose whole window stays clear of the frame
    search = np.zeros(P, SEARCH_DTYPE)
    search["cx"] = [c, lo, hi, c + 0.5, c - 3.37, hi - 0.25]
    search["cy"] = [c, hi, lo, c - 0.5, c + 2.81, lo + 0.75]
    search["xl"] = [xr, xr, xr, xr, xr / 3, 0]; search["xr"] = [xr, xr, xr, xr / 3, xr, xr]
    search["yl"] = [xr, xr, xr, 0, xr, xr];     search["yr"] = [xr, xr, xr, xr, xr / 3, xr]
    e.align(0, P, search)
    row = 0
    for p in range(P):
        s = search[p]
        for iy in range(-int(s["yl"] / ts), int(s["yr"] / ts) + 1):
            for ix in range(-int(s["xl"] / ts), int(s["xr"] / ts) + 1):
                got, kern = e.batch_row_spectrum(row)
                assert kern == 1, "the grouped row kernel should have handled this batch"
                cc = oracle.normalize_ring(oracle.polar2dm(imgs[p], float(s["cx"]) + ix * ts, float(s["cy"]) + iy * ts, numr), numr)
                want = oracle.frngs(cc, numr)
                scale = np.abs(want).max()
                # 2e-5 of the maximum for centred particles as everywhere else; a window pushed against the frame's corner
                # sees mostly background, its largest coefficient is ~half as large and the float error of 17k-sample
                # rings the same (the general kernel measures 2.6e-5 on the same rows: tests/diag_tile.py)
                assert np.abs(got - want).max() <= (2e-5 if p in (0, 3, 4) else 4e-5) * scale, (nx, p, ix, iy, np.abs(got - want).max() / scale)
                row += 1
    e.close()
