"""Self-consistency certificate for configurations no stored reference run exists for (SURVEY.md section 8c fallback list).

For a graph too large for the reference (``admm_solver_v3.py`` needs dense O(|E|^2) matrices) the ADMM answer is bracketed:

* lower bound: the straight-line distance |t - s| — no path can be shorter (costs are path lengths, reference
  ``GCS_utils.py:184-211``);
* the relaxed cost of the ADMM iterates (``compute_cost``): a lower bound of the best path's cost up to the residual;
* upper bounds: the cost of concrete s-t paths, each optimised over its waypoints by the path-restricted convex program
  (reference ``GCS_utils.py:17-89``; here the Drake-free classic solver on the chain of the path's regions):
  (a) the path that follows the largest relaxed flow out of every vertex (what the reference's y_e-proportional
  rounding walk finds with the highest probability), (b) the Dijkstra path on the overlap graph with
  centroid-distance weights — independent of the ADMM.

``ok`` = lower <= relaxed (1 + tol) and relaxed <= upper (1 + tol): the relaxation value sits between the trivial lower
bound and a feasible path.  ``gap`` = (best upper - relaxed) / best upper, the integrality + convergence gap.
"""
from __future__ import annotations

import time

import numpy as np

__all__ = ["certificate", "flow_path", "dijkstra_path", "path_cost"]


def _region_dicts(g, verts):
    A = {}
    b = {}
    names = {g.src: "s", g.dst: "t"}
    for v in verts:
        k = names.get(int(v), int(v))
        A[k] = g.polyA[g.poly_off[v]:g.poly_off[v + 1]]
        b[k] = g.polyb[g.poly_off[v]:g.poly_off[v + 1]]
    return A, b


def flow_path(g, y_e, max_len=None):
    """greedy walk from the source along the largest relaxed flow to an unvisited head; None if it strands"""
    out_edges = [[] for _ in range(g.nV)]
    for e in range(g.nE):
        out_edges[int(g.edge_tail[e])].append(e)
    path, seen, cur = [g.src], {g.src}, g.src
    limit = max_len or g.nV
    while cur != g.dst and len(path) <= limit:
        cand = [(y_e[e], int(g.edge_head[e])) for e in out_edges[cur] if int(g.edge_head[e]) not in seen and y_e[e] > 1e-9]
        if not cand:
            return None
        cur = max(cand)[1]
        seen.add(cur)
        path.append(cur)
    return path if cur == g.dst else None


def dijkstra_path(g):
    """shortest s-t path in the overlap graph, edge weight = distance of the regions' interior points plus 1 % of the head's
    distance from the straight line s-t (on a lattice of regions every monotone staircase has the same centroid length; the
    extra term picks the one that hugs the line)"""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra
    c = g.interior_points()
    d = c[g.dst] - c[g.src]
    nrm = float(np.linalg.norm(d))
    off = np.abs((c[:, 0] - c[g.src, 0]) * d[1] - (c[:, 1] - c[g.src, 1]) * d[0]) / nrm if nrm > 0 else np.zeros(g.nV)
    w = np.linalg.norm(c[g.edge_tail] - c[g.edge_head], axis=1) + 0.01 * off[g.edge_head] + 1e-9
    M = csr_matrix((w, (g.edge_tail.astype(np.int64), g.edge_head.astype(np.int64))), shape=(g.nV, g.nV))
    dist, pred = dijkstra(M, directed=True, indices=g.src, return_predecessors=True)
    if not np.isfinite(dist[g.dst]):
        return None
    path, cur = [g.dst], g.dst
    while cur != g.src:
        cur = int(pred[cur])
        path.append(cur)
    return path[::-1]


def path_cost(g, path):
    """-> (length, pure_chain).  Optimal length of the piecewise-linear curve through the path's regions (the reference's
    1e-4 per edge is NOT included), from the classic relaxation on the path's region set.  When those regions overlap only
    consecutively (``pure_chain``: the sub-graph has exactly one s-t path) the relaxation is tight and the value is the cost
    of a feasible path, i.e. an upper bound of the optimum; otherwise it is only the relaxation over that region set."""
    from .classic import solve_classic
    A, b = _region_dicts(g, path)
    res = solve_classic(A, b, 2, round_solution=False)
    if res["status"] != "optimal":
        return None, False
    length = float(sum(np.linalg.norm(z[:2] - z[2:]) for z in res["z_v_sol"].values()))
    return length, len(res["E"]) == 2 * (len(path) - 1)


def certificate(g, z_v, z_e, tol=1e-3, max_chain=1500):
    t0 = time.perf_counter()
    c = g.interior_points()
    s_pt, t_pt = c[g.src], c[g.dst]
    lower = float(np.linalg.norm(t_pt - s_pt))
    length = float(np.sum(np.linalg.norm(z_v[:, :2] - z_v[:, 2:], axis=1)))
    relaxed = length + 1e-4 * float(np.sum(z_e[:, 4]))
    out = {"lower_bound_straight_line": lower, "relaxed_cost": relaxed, "relaxed_length": length}
    uppers = {}
    for name, path in (("flow_path", flow_path(g, z_e[:, 4])), ("dijkstra_path", dijkstra_path(g))):
        if path is None or len(path) > max_chain:
            out[name] = None if path is None else {"vertices": len(path), "cost": None, "note": "chain too long for the comparator"}
            continue
        cost, pure = path_cost(g, path)
        out[name] = {"vertices": len(path), "length": cost, "pure_chain": pure}
        if cost is not None and pure:
            uppers[name] = cost
    if uppers:
        best = min(uppers.values())
        out["best_upper_bound"] = best
        out["gap"] = (best - length) / best
        out["ok"] = bool(lower <= length * (1 + tol) and length <= best * (1 + tol))
    else:
        out["ok"] = bool(lower <= length * (1 + tol))
    # with every equality and consensus constraint met exactly the relaxed length cannot be below |t - s| (the segments telescope);
    # a deficit measures how far the iterate still is from consensus in GLOBAL coordinates
    out["relaxed_length_minus_lower_bound_rel"] = (length - lower) / lower
    out["seconds"] = time.perf_counter() - t0
    return out
