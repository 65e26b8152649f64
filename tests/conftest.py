import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def meshes():
    return dict(np.load(os.path.join(GOLDEN, "meshes.npz")))


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "golden_cfg1.npz")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def pd():
    """The product's pyDeform module on a CUDA device (fails loudly without one)."""
    import torch
    from meshode_b200 import capi
    from meshode_b200 import pyDeform as m
    capi.require_device()
    assert torch.cuda.is_available()
    return m
