import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpss():
    """The ctypes harness over libgpss.so; GPU tests fail loudly (no skip, no fallback) if it cannot run."""
    import gp_ss_ak_b200 as G
    G.load_library()
    return G
