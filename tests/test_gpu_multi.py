"""1-GPU vs N-GPU equality (last bullet) as a test: scripts/check_multi_gpu.py under torchrun on every
GPU of the box (NCCL): mref_ali2d and the class-bound loop sharded over N GPUs against the same loops on one GPU.
Skipped on a single-GPU box (the driver's `-m gpu` run); the N=2 / N=8 logs of the round are kept under profiles/."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_sharded_loops_equal_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(root, "scripts", "check_multi_gpu.py")]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout)
    log = os.path.join(root, "gpurun_out")
    os.makedirs(log, exist_ok=True)
    open(os.path.join(log, "multi_gpu_check_n%d.log" % n), "w").write(out.stdout + out.stderr[-2000:])
    assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-2000:]
    assert out.stdout.count("OK") >= 2 and "FAIL" not in out.stdout
