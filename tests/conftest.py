import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
import utils  # noqa: E402,F401  (registers gcs_admm_b200 and the root-level shim)

GOLDEN = os.path.join(ROOT, "tests", "golden")
ALL_PROBLEMS = ["test1", "test2", "test3", "test_autogen1", "test_autogen2",
                "benchmark1", "benchmark2", "benchmark3", "benchmark4"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """Problem + stored reference run from tests/golden/<name>.npz
    (made by tools/export_golden.py from the reference's test_data and pickles)."""
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"), allow_pickle=False)
    keys = [k if k in ("s", "t") else int(k) for k in d["keys"].tolist()]
    As = {k: d[f"A_{i}"] for i, k in enumerate(keys)}
    bs = {k: d[f"b_{i}"] for i, k in enumerate(keys)}
    return As, bs, int(d["n"]), d, keys


@pytest.fixture(scope="session")
def golden():
    return load_golden
