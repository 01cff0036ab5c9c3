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
    """The CPU oracle (test infrastructure; built with gcc on first use)."""
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    """A GPU context through the C ABI.  Fails (does not skip) when the CUDA library cannot be used."""
    import ransac_b200
    c = ransac_b200.Context(0)
    yield c
    c.close()
