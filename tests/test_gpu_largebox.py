"""Boxes beyond what fits into shared memory (through the C ABI, against the CPU oracle).

The image tile of the row kernels, of rot_shift2D and the three spectra of the reference update live in shared memory
for the named configurations (90 and 128 pixels).  Larger boxes take the same kernels with the image taps read from
global memory (general row kernel: boxes beyond ~170 pixels; rot_shift2D: beyond ~238) and the transforms of the
reference update in a global scratch buffer (beyond ~136 pixels); maxrin = 1024 (ou > 81) takes the shared-memory
CCF kernel.  Same bars as the named geometries (tests/test_gpu_configs.py)."""
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


@pytest.mark.parametrize("nx,ou,xr", [(160, 72, 2), (192, 80, 2), (256, 100, 1)])
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
