import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GX3_FIXTURE = os.path.join(ROOT, "tests", "golden", "gx3_grid.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure); built on demand with gcc."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def evp_lib():
    """libevp_b200.so, built on demand with nvcc (cross-compiles without a GPU)."""
    from cice4_b200 import build, evp
    build.build()
    return evp.load_library()
