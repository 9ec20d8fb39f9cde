"""Golden vectors produced by THE REFERENCE ITSELF: the answers of the reference's own CUDA library
(/root/reference/cuda/gpu_aln_{common,noref}.cu, compiled unchanged for sm_100 by baseline/build_ref_cuda.sh) for a
deterministic synthetic stack, obtained through its stock entry points (scripts/compare_ref_cuda.py: pre_align_init ->
pre_align_fetch -> reset_shifts -> mref_align_run -> AlignParam[]).  Run on a GPU box:

    python tests/golden/make_refcuda_case.py          # -> tests/golden/refcuda_mref_outputs.npz, refcuda_reffree_outputs.npz

The inputs are NOT stored: scripts/compare_ref_cuda.py: make_inputs(P, V, snr) regenerates them bit for bit from seeds
(numpy Generator).  The fixture holds the generation parameters and, per particle, (ref_id, shift_x, shift_y, angle,
mirror) as the library left them in AlignParam[].  tests/test_oracle.py checks the CPU oracle against it (no GPU): at this
noise level the reference library and the EMAN2-semantics oracle must choose the same class, mirror flag and integer shift
and agree on the angle to a fraction of a ring sample -- the reference library's arithmetic differs (256 bilinear samples
per ring, no Normalize_ring), so this pins the CONVENTIONS and the discrete answers of the oracle, not its last digits."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
P, V, SNR = 512, 12, 1.0

if __name__ == "__main__":
    import shutil
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    # mref_align_run with 12 references; pre_align_run (the reference-free entry point) with one reference on a one-view stack
    # ... and mref_align_run on a stack ten times noisier, where the two arithmetics no longer agree on every particle
    for mode, views, snr, name in (("mref", V, SNR, "refcuda_mref_outputs.npz"), ("reffree", 1, SNR, "refcuda_reffree_outputs.npz"),
                                   ("mref", V, 0.1, "refcuda_mref_snr0.1_outputs.npz")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "compare_ref_cuda.py"), str(P), str(views), str(snr), mode],
                           cwd=ROOT, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
        a = np.load(os.path.join(ROOT, "gpurun_out", "refcuda_outputs_%s_P%d_V%d_snr%g.npy" % (mode, P, views, snr)))
        np.savez(os.path.join(HERE, name), particles=P, views=views, snr=snr, nx=90, ou=36, xr=3, ts=1.0,
                 ref_id=a[:, 0].astype(np.int32), shift_x=a[:, 1].astype(np.float32), shift_y=a[:, 2].astype(np.float32),
                 angle=a[:, 3].astype(np.float32), mirror=a[:, 4].astype(np.int32))
        shutil.copy(os.path.join(HERE, name), os.path.join(ROOT, "gpurun_out", name))
        print("wrote", name, ":", P, "particles")
