"""Drop-in check at the ABI level (scripts/compare_ref_cuda.py): ONE ctypes harness in the reference driver's call order,
run with the reference's own CUDA library (oracle/_ref/gpu_aln_pack.so: the sm_100 build of the reference's sources, a
built artefact that travels to the GPU box) and with libcryo_ralib.so, each in a child process, on the same synthetic stack
with known poses.  Not parity (gpu_isac's arithmetic against EMAN2's): the two libraries must agree on the class, the mirror
flag, the accumulated shift and the angle CONVENTIONS of AlignParam.  Skipped where the reference library was not built."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode,views", [("mref", 12), ("mref_m", 12), ("reffree", 1)])
def test_same_harness_same_answers(mode, views):
    if not any(os.path.exists(os.path.join(ROOT, d, "_ref", "gpu_aln_pack.so")) for d in ("oracle", "baseline")):
        pytest.skip("oracle/_ref/gpu_aln_pack.so not built (`make -C oracle ref` needs the reference's sources)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "compare_ref_cuda.py"), "1024", str(views), "1.0", mode],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-500:] + out.stderr[-500:]
    rep = json.loads(out.stdout[out.stdout.index("{"):])
    b = rep["between"]
    if mode != "reffree":
        assert rep["reference"]["view_recovered"] >= 0.99 and rep["this"]["view_recovered"] >= 0.99
        assert b["same_class"] >= 0.99
    if mode == "mref_m":
        # mref_align_run_m: [2R][nx][nx] even sums then odd sums + get_num_ref.  The reference library transforms with a
        # bilinear texture fetch, this one with EMAN2's quadratic interpolation, so the sums correlate instead of agreeing
        # to digits; even - odd cancels a class's signal and correlates only if both libraries put the same particles into
        # the same half (it would be NEGATIVE with the halves swapped)
        sm = rep["sums"]
        assert sm["class_sizes_equal"]
        assert sm["ncc_same_half"]["min"] >= 0.98 and sm["ncc_even_minus_odd"]["min"] >= 0.4
    assert rep["this"]["mirror_recovered"] >= 0.99
    assert b["same_class_and_mirror"] >= 0.99
    assert b["angle_diff_deg"]["p99"] <= 0.5 * 360.0 / 256          # half a ring sample
    assert b["shift_diff_px"]["within_1px"] >= 0.99 and b["shift_diff_px"]["median"] == 0.0
