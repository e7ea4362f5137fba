import os
import sys

import pytest

# tests/test_gpu_local_images.py runs several images of a decomposition as contexts of ONE process on ONE device; an
# image's kernels wait on the device for its neighbours', so their streams must not share a hardware work queue (the
# default of 8 connections is fewer than the 3 streams x 3 contexts of those tests).  Read when CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; see oracle/mpdata_oracle.hpp)."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def has_gpu():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        cudart = None
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; GPU tests fail (not skip) when it cannot be loaded."""
    import icar_b200
    icar_b200.lib()
    return icar_b200
