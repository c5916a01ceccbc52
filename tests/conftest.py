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
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def nab():
    """The product's host layer; building the library is a CPU-only step (nvcc cross-compiles)."""
    from nalgebra_b200 import build as nab_build
    nab_build.build()
    import nalgebra_b200
    return nalgebra_b200
