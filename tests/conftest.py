import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def tg():
    """the product package (ctypes over libtristan_gpu.so); builds the library if needed"""
    import __graft_entry__ as g
    g.build()
    import tristan_mp_pu_master_densdecomp_b200 as pkg
    return pkg
