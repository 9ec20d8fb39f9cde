"""Shared plumbing of the CLI drivers: the reference's flag surface (test_mref_gpu_align.py:1141-1159,
test_reffree.py:849-873), one process per GPU via torch.distributed (torchrun), logging."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def first_of(value):
    """The reference takes --xr/--yr/--ts as strings of lists and uses entry 0 (test_reffree.py:173-174)."""
    return float(str(value).split()[0])


def add_alignment_flags(p, reffree=False):
    p.add_argument("--ir", type=int, default=1, help="inner radius for rotational correlation > 0 (set to 1)")
    p.add_argument("--ou", type=int, default=-1, help="outer radius for rotational correlation < nx/2-1 (set to nx/2-2)")
    p.add_argument("--rs", type=int, default=1, help="step between rings in rotational correlation > 0 (set to 1)")
    p.add_argument("--xr", type=str, default="0", help="range for translation search in x direction")
    p.add_argument("--yr", type=str, default="-1", help="range for translation search in y direction (default: same as xr)")
    p.add_argument("--ts", type=str, default="1", help="step of translation search in both directions")
    p.add_argument("--center", type=int, default=-1 if reffree else 1, help="-1 average centering (ref-free), 0 none, 1 phase centre of gravity")
    p.add_argument("--maxit", type=int, default=10 if not reffree else 0, help="maximum number of iterations")
    p.add_argument("--CTF", action="store_true", help="accepted for compatibility; CTF is forced off (test_mref_gpu_align.py:303-308)")
    p.add_argument("--snr", type=float, default=1.0, help="accepted for compatibility")
    p.add_argument("--function", type=str, default="ref_ali2d", help="reference preparation function (only ref_ali2d)")
    p.add_argument("--rand_seed", type=int, default=1000, help="random seed for reseeding vanished references")
    p.add_argument("--MPI", action="store_true", help="accepted for compatibility; parallelism comes from torchrun")
    p.add_argument("--gpu_devices", type=str, default="", help="comma-separated CUDA device ids (default: LOCAL_RANK)")
    p.add_argument("--gpu_info", action="store_true", help="print GPU information and exit")


def resolve_device(args, local):
    """CUDA device of this rank: entry LOCAL_RANK of --gpu_devices (test_mref_gpu_align.py:1155), else LOCAL_RANK."""
    if getattr(args, "gpu_devices", ""):
        ids = [int(x) for x in args.gpu_devices.split(",") if x != ""]
        return ids[local % len(ids)]
    return local


def init_distributed(args=None):
    """Returns (comm, rank, world, device index).  The device is resolved BEFORE the process group is bound, so the
    NCCL communicator, torch's current device and the engine all sit on the same GPU (--gpu_devices included)."""
    from cryo_ralib_b200.mref import LocalComm, TorchComm
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = resolve_device(args, local)
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
        return TorchComm(), rank, world, device
    return LocalComm(), 0, 1, device


def pick_device(args, device):
    """Kept for the drivers' call sites: init_distributed(args) already resolved the device."""
    return device


class Log(object):
    def __init__(self, outdir, rank):
        self.f = open(os.path.join(outdir, "logfile"), "a") if rank == 0 else None

    def add(self, msg):
        if self.f:
            line = time.strftime("%Y-%m-%d %H:%M:%S :: ") + msg
            print(line)
            self.f.write(line + "\n")
            self.f.flush()
