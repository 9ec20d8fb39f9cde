"""Golden vectors produced by THE REFERENCE ITSELF: the answers of the reference's own CUDA library
(/root/reference/cuda/gpu_aln_{common,noref}.cu, compiled unchanged for sm_100 by baseline/build_ref_cuda.sh) for a
deterministic synthetic stack, obtained through its stock entry points (scripts/compare_ref_cuda.py: pre_align_init ->
pre_align_fetch -> reset_shifts -> mref_align_run -> AlignParam[]).  Run on a GPU box:

    python tests/golden/make_refcuda_case.py          # -> tests/golden/refcuda_mref_outputs.npz

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
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "compare_ref_cuda.py"), str(P), str(V), str(SNR), "mref"],
                       cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    a = np.load(os.path.join(ROOT, "gpurun_out", "refcuda_outputs_mref_P%d_V%d_snr%g.npy" % (P, V, SNR)))
    np.savez(os.path.join(HERE, "refcuda_mref_outputs.npz"), particles=P, views=V, snr=SNR, nx=90, ou=36, xr=3, ts=1.0,
             ref_id=a[:, 0].astype(np.int32), shift_x=a[:, 1].astype(np.float32), shift_y=a[:, 2].astype(np.float32),
             angle=a[:, 3].astype(np.float32), mirror=a[:, 4].astype(np.int32))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    import shutil
    shutil.copy(os.path.join(HERE, "refcuda_mref_outputs.npz"), os.path.join(ROOT, "gpurun_out", "refcuda_mref_outputs.npz"))
    print("wrote refcuda_mref_outputs.npz:", P, "particles")
