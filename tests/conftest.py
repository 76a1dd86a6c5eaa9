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


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: without one they are skipped (plain `pytest` on a CPU box stays green) instead of
    failing in gcsadmm_create with "no CUDA device"."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        from gcs_admm_b200 import lib
        ndev = lib.load().gcsadmm_device_count()
    except Exception:
        ndev = 0
    if ndev == 0:
        skip = pytest.mark.skip(reason="no CUDA device (gpu tests run on the B200 box)")
        for it in gpu_items:
            it.add_marker(skip)


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


def build_emu():
    """Compiles gcs-admm_b200/csrc/emulate.cpp (the kernel sources with GCS_EMULATE: test infrastructure only) and returns
    the path of libgcsemu.so."""
    import subprocess
    csrc = os.path.join(ROOT, "gcs-admm_b200", "csrc")
    so, src = os.path.join(csrc, "libgcsemu.so"), os.path.join(csrc, "emulate.cpp")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        base = ["g++", "-O2", "-fPIC", "-shared", "-o", so, src]
        if subprocess.call(base[:2] + ["-fopenmp"] + base[2:], stderr=subprocess.DEVNULL) != 0:
            subprocess.check_call(base)
    return so
