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


# ---------------------------------------------------------------------------------------------------
# near-tie / exclusion accounting: north_star says decisions inside the tolerance band are "counted and reported".
# Tests call parity_report(name, **counts); the lines are printed in the terminal summary (visible with -q) and, when
# gpurun_out/ exists, also written to gpurun_out/parity_report.json.
# ---------------------------------------------------------------------------------------------------
_PARITY_REPORT = []


def parity_report(name, **counts):
    _PARITY_REPORT.append((name, {k: (v.item() if hasattr(v, "item") else v) for k, v in counts.items()}))


def pytest_terminal_summary(terminalreporter):
    if not _PARITY_REPORT:
        return
    terminalreporter.write_sep("-", "parity report: compared / flagged (near-tie band) / mismatching counts per test")
    for name, c in _PARITY_REPORT:
        terminalreporter.write_line(f"{name}: " + ", ".join(f"{k}={v}" for k, v in c.items()))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        import json
        try:
            with open(os.path.join(out_dir, "parity_report.json"), "w") as f:
                json.dump([{"test": n, **c} for n, c in _PARITY_REPORT], f, indent=1)
        except OSError:
            pass
