#!/bin/bash
# The three drivers end to end on small synthetic stacks (run under gpurun, optionally --gpus 2): outputs present,
# class sizes add up, drm / aqm files written, apply_transform against the oracle.
set -e
D=gpurun_out/drv; rm -rf $D; mkdir -p $D
N=${1:-1}
python drivers/make_synthetic.py --n 2000 --nx 90 --refs 8 $D/stack.npy $D/refs.npy > /dev/null
if [ "$N" -gt 1 ]; then RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"; else RUN=python; fi
$RUN drivers/test_mref_gpu_align.py $D/stack.npy $D/refs.npy $D/mref --ou=36 --xr=3 --yr=3 --ts=1 --maxit=3 2>&1 | tail -2
ls $D/mref | tr '\n' ' '; echo
python - <<PY
import sys, glob, numpy as np
sys.path.insert(0, ".")
from cryo_ralib_b200 import stackio
p, c = stackio.read_params("$D/mref/params.txt")
assert p.shape == (2000, 4) and len(c) == 2000 and set(np.unique(c)) <= set(range(8))
aq = sorted(glob.glob("$D/mref/aqm*.mrcs")); dr = sorted(glob.glob("$D/mref/drm*.txt"))
assert len(aq) == 3 and len(dr) >= 3 * 6, (len(aq), len(dr))
r = stackio.read_stack(aq[-1]); assert r.shape == (8, 90, 90) and np.isfinite(r).all()
f = np.loadtxt(dr[0]); assert f.shape == (46, 3) and abs(f[0, 1] - 1.0) < 1e-3 and f[-1, 0] == 0.5
print("mref driver ok:", len(aq), "aqm stacks,", len(dr), "drm curves, class sizes", np.bincount(c, minlength=8))
PY
$RUN drivers/test_reffree_gpu_align.py $D/stack.npy $D/reffree --ou=36 --xr="2 1" --ts="1 0.5" --maxit=2 2>&1 | tail -1
ls $D/reffree | tr '\n' ' '; echo
python drivers/apply_transform.py $D/stack.npy $D/mref/params.txt $D/aligned.mrcs --averages $D/avg.mrcs --ou 36 | tail -1
rm -f $D/stack.npy $D/aligned.mrcs
