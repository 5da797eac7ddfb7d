import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def hb():
    """Initialised native library on cuda:0 (gpu tests only)."""
    import torch

    assert torch.cuda.is_available(), "gpu-marked test running without a GPU"
    from hirest_b200 import _lib

    return _lib.init(0)
