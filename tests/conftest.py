"""pytest configuration: registers the `gpu` marker and puts the product package (flat modules that
mirror the reference's `src/` layout) and the oracle on sys.path."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neural-implicit-queries_b200")
GOLD = os.path.join(ROOT, "tests", "golden")

for p in (os.path.join(ROOT, "oracle"), PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def sample_params(name):
    """The four sample MLPs of the reference, shipped as a derived fixture (tests/golden/mlps.npz)."""
    with np.load(os.path.join(GOLD, "mlps.npz")) as d:
        return {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith(name + "/")}


@pytest.fixture(scope="session")
def mlps():
    return {n: sample_params(n) for n in ("fox", "bunny", "hammer", "birdcage_occ")}
