import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "host_sim")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """CPU restatement of the reference (the checker)."""
    import pyoracle
    pyoracle.build()
    return pyoracle.load()


@pytest.fixture(scope="session")
def sim():
    """Device headers compiled for the host: kernel logic under test without a GPU."""
    import pysim
    return pysim.load()


@pytest.fixture(scope="session")
def gpu():
    """libhc_b200.so on cuda:0 through the C ABI.  Fails loudly if the extension is missing."""
    import __graft_entry__ as ge
    from hcb200 import lib
    if not os.path.exists(lib.LIB_PATH):
        ge.build()
    return lib.load(0)
