import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session")
def built_lib():
    """libmwrt.so, built in-tree (nvcc cross-compiles without a GPU)."""
    from mwr_fast_forward_operators_and_lbls_b200 import build
    return build.build()
