import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# host-logic tests place buffers in host memory when no GPU is present; kernels still refuse to run
os.environ.setdefault("RENDERTOY_B200_POOL_BYTES", str(64 * 1024 * 1024))
try:
    import torch
    HAS_CUDA = torch.cuda.is_available()
except Exception:  # pragma: no cover
    HAS_CUDA = False
if not HAS_CUDA:
    os.environ.setdefault("RENDERTOY_B200_HOST_BUFFERS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if HAS_CUDA:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def ren():
    import rendering
    return rendering
