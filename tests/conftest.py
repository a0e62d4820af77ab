import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_index():
    with open(os.path.join(GOLDEN, "index.json")) as fh:
        return json.load(fh)


def golden_case(cid):
    meta = golden_index()[cid]
    data = np.load(os.path.join(GOLDEN, cid + ".npz"))
    return meta, data


def rel_l1(a, b):
    """Relative L1 difference per variable (the north-star parity norm): sum|a-b| / sum|b| over the grid."""
    axes = tuple(range(a.ndim - 1))
    den = np.sum(np.abs(b), axis=axes)
    num = np.sum(np.abs(a - b), axis=axes)
    return np.where(den > 0, num / np.where(den > 0, den, 1), num)


@pytest.fixture(scope="session")
def hostsim_lib():
    """The host-simulated build of the kernel sources (g++): test infrastructure, never used by the package."""
    from astrea_b200 import _native, build
    return _native.bind(build.build(hostsim=True))


@pytest.fixture(scope="session")
def device_lib_path():
    from astrea_b200 import build
    return build.build(hostsim=False)
