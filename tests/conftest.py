import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def rand5():
    return dict(np.load(os.path.join(GOLDEN, "rand5.npz")))


@pytest.fixture(scope="session")
def fit3():
    return dict(np.load(os.path.join(GOLDEN, "fit3.npz")))


@pytest.fixture(scope="session")
def fit5():
    return dict(np.load(os.path.join(GOLDEN, "fit5.npz")))


@pytest.fixture(autouse=True)
def _gpu_state_is_clean(request):
    """Every GPU test starts and ends on torch's default stream with no pending CUDA error: a test (or a library path) that
    leaks a stream context or a sticky error would otherwise fail a LATER test, far from its cause."""
    import torch
    if "gpu" not in request.keywords or not torch.cuda.is_available():
        yield
        return
    assert torch.cuda.current_stream() == torch.cuda.default_stream(), "a previous test left a side stream current"
    yield
    torch.cuda.synchronize()
    assert torch.cuda.current_stream() == torch.cuda.default_stream(), f"{request.node.name} left a side stream current"
