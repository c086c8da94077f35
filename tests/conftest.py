import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only)."""
    from oracle import oracle as orc

    orc.build()
    return orc


@pytest.fixture(scope="session")
def engine():
    """One engine on cuda:0.  No fallback: without the built library or a B200 this errors out."""
    import rbqoc_b200

    eng = rbqoc_b200.Engine(0)
    yield eng
    eng.close()
