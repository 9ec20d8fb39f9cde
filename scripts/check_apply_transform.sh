#!/bin/bash
# drivers/apply_transform.py end to end on a small synthetic set, against the oracle's rot_shift2d (run under gpurun)
set -e
D=gpurun_out/drv2; mkdir -p $D
python drivers/make_synthetic.py --n 300 --nx 90 --refs 6 $D/stack.npy $D/refs.npy > /dev/null
rm -rf $D/out
python drivers/test_mref_gpu_align.py $D/stack.npy $D/refs.npy $D/out --ou=36 --xr=3 --yr=3 --maxit=2 2>&1 | tail -1
python drivers/apply_transform.py $D/stack.npy $D/out/params.txt $D/aligned.mrcs --averages $D/avg.mrcs --ou 36
python - <<PY
import sys, numpy as np
sys.path.insert(0, ".")
from cryo_ralib_b200 import stackio
from oracle import oracle as o
st = stackio.read_stack("$D/stack.npy"); al = stackio.read_stack("$D/aligned.mrcs")
p, c = stackio.read_params("$D/out/params.txt")
mask = o.model_circle(36, 90)
worst = 0
for i in (0, 17, 299):
    w = o.rot_shift2d(o.normalize_mask(st[i], mask, 0), p[i, 0], p[i, 1], p[i, 2], int(p[i, 3]))
    worst = max(worst, float(np.abs(al[i] - w).max() / np.abs(w).max()))
print("apply_transform vs oracle rot_shift2d: max rel diff %.2e" % worst)
PY
rm -f $D/stack.npy $D/aligned.mrcs
