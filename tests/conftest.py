import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand with oracle/Makefile."""
    from oracle import oracle as O
    O.build()
    O.set_threads(min(8, O.max_threads()))
    return O


@pytest.fixture(scope="session")
def b2lib():
    """libb2icp.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from icpslam_b200 import build as B
    B.build()
    from icpslam_b200 import registration as R
    R.load_library()
    return R
