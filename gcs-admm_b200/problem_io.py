"""Problem files: load the reference's ``test_data/<name>.py`` format and write it.

Format (reference ``test_data/benchmark1.py``, emitted by reference
``test_generator.py:23-79``): an importable module exposing ``As`` / ``bs`` (dicts
keyed ``"s"``, ``"t"``, ``0..k-1`` of ``(m, n)`` / ``(m,)`` arrays) and ``n``;
``s, t, N, M`` are present but ignored by the solver scripts
(reference ``admm_solver_v3.py:47-48``).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_test_file(name, test_data_dir=None):
    """Import ``<test_data_dir>/<name>.py`` exactly as reference
    ``admm_solver_v3.py:42-51`` does and return ``(As, bs, n)``.

    Raises ModuleNotFoundError when the file does not exist (the CLI turns that into
    the reference's message + exit code 1)."""
    import utils  # noqa: F401  root-level shim: the problem files import it by that name
    test_data_dir = test_data_dir or os.path.join(_ROOT, "test_data")
    path = os.path.join(test_data_dir, f"{name}.py")
    if not os.path.exists(path):
        raise ModuleNotFoundError(f"No module named '{name}'")
    spec = importlib.util.spec_from_file_location(f"_gcs_problem_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.As, mod.bs, int(mod.n)


def write_test_file(filename, As, bs, s=None, t=None, N=None, M=None, header=None):
    """Write a problem in the reference's module format (own layout, same names).

    ``As``/``bs`` hold the regions keyed by int plus optionally ``"s"``/``"t"``; when
    ``s``/``t`` points are given their boxes are emitted through
    ``convert_pt_to_polytope(pt, eps=1e-6)`` like the reference's generator does."""
    keys = sorted(k for k in As if isinstance(k, (int, np.integer)))
    lines = []
    if header:
        lines.append('"""' + header.strip() + '"""')
    lines += ["import os", "import sys", "", "import numpy as np", "",
              "sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))",
              "from utils import convert_pt_to_polytope, visualize_results", ""]

    def arr(a):
        return repr(np.asarray(a, dtype=float).tolist())     # repr(float) round-trips exactly
    if s is not None:
        lines.append(f"s = np.array({arr(s)})")
        lines.append(f"t = np.array({arr(t)})")
        lines.append("A_s, b_s = convert_pt_to_polytope(s, eps=1e-6)")
        lines.append("A_t, b_t = convert_pt_to_polytope(t, eps=1e-6)")
    else:
        for k in ("s", "t"):
            lines.append(f"A_{k} = np.array({arr(As[k])})")
            lines.append(f"b_{k} = np.array({arr(bs[k])})")
    lines.append("")
    lines.append("regions = [")
    for k in keys:
        lines.append(f"    (np.array({arr(As[k])}), np.array({arr(bs[k])})),")
    lines.append("]")
    lines.append("")
    lines.append('As = {"s": A_s, "t": A_t}')
    lines.append('bs = {"s": b_s, "t": b_t}')
    lines.append("for _k, (_A, _b) in enumerate(regions):")
    lines.append("    As[_k] = _A")
    lines.append("    bs[_k] = _b")
    lines.append("")
    lines.append("n = regions[0][0].shape[1]")
    lines.append("")
    lines.append("# rounding hints (unused by the solvers, kept for format compatibility)")
    lines.append(f"N = {int(N) if N is not None else max(1, len(keys) // 5)}")
    lines.append(f"M = {int(M) if M is not None else max(1, 2 * len(keys) // 5)}")
    lines.append("")
    with open(filename, "w") as fh:
        fh.write("\n".join(lines))
