import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))      # parity_util.py (shared helpers of the GPU tests)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """The product library, built in-tree (nvcc cross-compiles without a GPU)."""
    from eventclip_b200 import build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_dev(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a -m gpu test was selected but no CUDA device is visible; there is no CPU fallback")
    return torch.device("cuda", 0)

