import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import agref
    return agref


@pytest.fixture(scope="session")
def agpkg():
    """The product package (autograd.jl_b200/, imported as agb200).  Needs the built .so."""
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    return ge.load_package()


@pytest.fixture()
def hostlib(agpkg):
    """Product host logic running on the numpy stand-in for the C ABI (tests/fake_abi.py)."""
    import fake_abi
    fake, restore = fake_abi.install(agpkg)
    try:
        yield agpkg
    finally:
        restore()


@pytest.fixture(scope="session")
def ag(agpkg):
    """The product package bound to cuda:0 (gpu tests only)."""
    agpkg.init(0)
    return agpkg
