import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge

    return ge.load_package()


@pytest.fixture(scope="session")
def orc():
    import __graft_entry__ as ge

    po = ge.load_oracle()
    if not os.path.exists(po.ORACLE_SO):
        po.build()
    return po


@pytest.fixture(scope="session")
def oracle(orc):
    return orc.Oracle()


@pytest.fixture(scope="session")
def ref(orc):
    """the unmodified reference (oracle/_ref/libsdslref.so); skip where it was never built"""
    if not orc.ref_available():
        pytest.skip("oracle/_ref/libsdslref.so not built")
    return orc.Ref()
