import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def b200lib():
    """The product library, built in-tree if needed (nvcc cross-compiles
    without a GPU).  No fallback: a build failure fails the test."""
    from openshadinglanguage_b200 import build
    build.build()
    import openshadinglanguage_b200 as ob
    return ob


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test running without a CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")
