"""BASELINE config 4: a batch of independent start/goal queries on one region set (default: benchmark4's 40 regions).

Every query is the region set plus its own (s, t) drawn by rejection sampling inside two distinct random regions
(``np.random.default_rng(seed)``).  benchmark4's region graph has components of 29/4/3/3/1 regions, so with
``feasible_only=False`` about 46 % of the pairs have no s-t path at all (their ADMM runs to ``max_it`` and their flows are
meaningless); ``feasible_only=True`` (default) draws the second region from the first one's component.
Queries are packed block-diagonally by ``graph.pack_batch`` and solved in one handle per GPU with per-problem
residuals / rho / stop — no communication; rank r of w takes queries r::w.
"""
from __future__ import annotations

import numpy as np

from .graph import build_graph, convert_pt_to_polytope
from .problem_io import load_test_file

__all__ = ["make_queries", "region_components"]


def region_components(As, bs):
    """component id of every region of a problem (keys other than 's' / 't'), from the overlap graph"""
    keys = [k for k in As if not isinstance(k, str)]
    A2 = {k: As[k] for k in keys}
    b2 = {k: bs[k] for k in keys}
    V, E, _, _ = build_graph(A2, b2)
    parent = {k: k for k in keys}

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    for a, b in E:
        parent[find(a)] = find(b)
    return {k: find(k) for k in keys}


def make_queries(n, seed=1, problem="benchmark4", feasible_only=True):
    """-> list of (As, bs) dicts, one per query, in the reference's problem format"""
    As, bs, _ = load_test_file(problem)
    regions = [k for k in As if not isinstance(k, str)]
    comp = region_components(As, bs) if feasible_only else None
    lo, hi = -25.0, 25.0
    rng = np.random.default_rng(seed)

    def sample(k):
        A, b = As[k], bs[k]
        while True:
            p = rng.uniform(lo, hi, size=2)
            if np.all(A @ p <= b - 1e-3):
                return p
    out = []
    while len(out) < n:
        a, c = rng.choice(len(regions), size=2, replace=False)
        if comp is not None and comp[regions[a]] != comp[regions[c]]:
            continue
        s, t = sample(regions[a]), sample(regions[c])
        Aq, bq = dict(As), dict(bs)
        Aq["s"], bq["s"] = convert_pt_to_polytope(s)
        Aq["t"], bq["t"] = convert_pt_to_polytope(t)
        out.append((Aq, bq))
    return out
