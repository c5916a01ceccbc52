import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a B200: skip them (instead of failing with NA_ECUDA -- the library has no CPU fallback) when
    none is present, so a plain `pytest tests` is green on a CPU box.  NAB_GPU_TESTS=1 forces them to run."""
    if os.environ.get("NAB_GPU_TESTS") == "1" or _have_gpu():
        return
    skip = pytest.mark.skip(reason="no sm_100 device: libnalgebra_b200 has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
