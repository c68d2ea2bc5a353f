import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_lib():
    """The loaded libb2f_cuda.so on a CUDA device; GPU tests must not silently skip the native
    path, so a missing library is an error, not a skip."""
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    from back2future_b200 import _lib
    return _lib.load()
