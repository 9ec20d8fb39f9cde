import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    o.lib()
    return o


@pytest.fixture(scope="session")
def small_set():
    """64 synthetic 90x90 particles + 10 initial references (config-1 geometry).  The references are averages
    of OTHER particles of the same data set (a particle that is part of a reference matches it at angle 0, shift 0,
    which would leave the angular search untested)."""
    from cryo_ralib_b200 import synth
    allp, truth = synth.make_particles(64 + 60, 90, 16, max_shift=3, seed=7)
    refs = synth.initial_references(allp[64:], 10, per_ref=6, seed=5)
    truth = {k: v[:64] for k, v in truth.items()}
    return np.ascontiguousarray(allp[:64]), refs, truth


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
