import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


@pytest.fixture(scope="session")
def oracle():
    """C restatement of the reference (oracle/wr_oracle.c) -- the checker."""
    from oracle import build_oracle
    from oracle.binding import Restatement
    build_oracle.build_restatement()
    return Restatement()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled into oracle/_ref (skips if it did not travel)."""
    from oracle import build_oracle
    from oracle.binding import Reference
    if os.path.isdir("/root/reference/src"):
        build_oracle.build_ref()
    if not Reference.available("strict"):
        pytest.skip("oracle/_ref not built")
    return Reference("strict")


@pytest.fixture(scope="session")
def product_lib():
    """libwaverange_b200.so; built with nvcc if stale (cross-compiles without a GPU)."""
    from waverange_b200 import build
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        build.build()
    assert os.path.exists(build.LIB), "libwaverange_b200.so missing and nvcc unavailable"
    return build.LIB


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    return torch


@pytest.fixture()
def codec(product_lib, torch_cuda):
    from waverange_b200.api import Codec
    c = Codec(device=0)
    yield c
    c.close()
