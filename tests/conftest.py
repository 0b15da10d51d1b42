import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def sim_engine():
    """Engine bound to the fiber-based CPU emulator build of the CUDA sources (kernel-logic tests)."""
    from tests.cusim.build_sim import build_sim
    from fabolas_b200 import Engine
    eng = Engine(device="cpu", lib_path=build_sim())
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def gpu_engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from fabolas_b200.build import build
    build()
    from fabolas_b200 import Engine
    eng = Engine("cuda:0")
    yield eng
    eng.close()


def rel_err(a, ref, axis=None):
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = np.max(np.abs(ref), axis=axis, keepdims=axis is not None)
    return float(np.max(np.abs(a - ref) / np.maximum(scale, 1e-300)))
