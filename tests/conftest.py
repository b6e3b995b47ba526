import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
STATE_KEYS = ("Theta", "Beta", "Gamma_shp", "Gamma_rte", "Lambda_shp", "Lambda_rte", "k_rte", "t_rte")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


@pytest.fixture(scope="session")
def golden_full():
    return np.load(os.path.join(GOLDEN, "toy_full.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_odd():
    return np.load(os.path.join(GOLDEN, "odd_full.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_svi():
    return np.load(os.path.join(GOLDEN, "toy_svi.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_pf():
    return np.load(os.path.join(GOLDEN, "toy_partial_fit.npz"), allow_pickle=False)
