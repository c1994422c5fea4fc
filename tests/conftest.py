import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """The UNMODIFIED reference (oracle/_ref); skip when neither the .so nor /root/reference exists."""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    refbind.lib()
    return refbind


@pytest.fixture(scope="session")
def hs():
    from tests.hostsim import hsbind
    hsbind.lib()
    return hsbind


@pytest.fixture(scope="session")
def gpu_lib():
    import csc_b200
    csc_b200.lib()
    if not csc_b200.available():
        pytest.fail("csc_b200 library loaded but no CUDA device is usable (no CPU fallback exists)")
    return csc_b200
