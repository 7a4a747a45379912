import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU test files listed in tests/UNVERIFIED_GPU.txt were written after the round's GPU budget was spent and have
    never run on a device: they are collected as non-strict xfail (a failure is reported as xfailed, a pass as xpassed)
    so that they cannot turn the verified suite red.  The list is deleted as soon as a GPU run has confirmed them."""
    path = os.path.join(ROOT, "tests", "UNVERIFIED_GPU.txt")
    if not os.path.exists(path):
        return
    names = {line.strip() for line in open(path) if line.strip() and not line.startswith("#")}
    for item in items:
        if os.path.basename(str(item.fspath)) in names:
            item.add_marker(pytest.mark.xfail(strict=False, reason="never run on a GPU yet (tests/UNVERIFIED_GPU.txt)"))


@pytest.fixture(scope="session")
def bc03():
    """BC03lr SSP template as float32 (rubix_b200/templates/bc03lr_f32.npz, made by tools/make_golden.py)."""
    d = np.load(os.path.join(ROOT, "rubix_b200", "templates", "bc03lr_f32.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def muse_wave():
    return np.load(os.path.join(GOLDEN, "muse_wave.npy"))


@pytest.fixture(scope="session")
def tng_subset():
    d = np.load(os.path.join(GOLDEN, "tng50_subset.npz"))
    return {k: d[k] for k in d.files}
