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
