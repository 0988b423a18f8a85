import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def oiio():
    """The product package with libb2r.so loaded (built if missing)."""
    import openimageio_b200 as pkg
    if not os.path.exists(os.path.join(ROOT, "openimageio_b200", "libb2r.so")):
        import __graft_entry__ as g
        g.build()
    pkg.lib()
    return pkg
