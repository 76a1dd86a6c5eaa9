"""Synthetic 2-D GCS problems.

* ``grid_problem(G, ...)`` — the scalable benchmark family (BASELINE.json configs
  3 and 5): G x G chamfered boxes overlapping their 4-neighbours, source in region (0,0), target in
  region (G-1,G-1).  Everything is produced as packed arrays (no per-region Python objects), so 10^6
  regions take seconds.
* ``generate_test_2D(...)`` — Drake-free restatement of reference ``test_generator.py:16-171``
  (Latin-hypercube seeds, radius = spacing_factor x nearest-seed distance, convex hull of a random subset
  of lattice points inside the radius); scipy ``ConvexHull`` / ``cKDTree`` replace ``VPolytope`` /
  ``HPolyhedron`` / dense ``cdist``.  Writes the reference's problem-file format.
"""
from __future__ import annotations

import numpy as np

from .graph import PackedGraph, build_graph_packed, convert_pt_to_polytope

__all__ = ["grid_problem", "grid_packed_graph", "generate_test_2D", "packed_to_dicts"]


def grid_problem(G, overlap=0.1, chamfer=(0.25, 0.40), jitter=0.0, seed=0):
    """Packed polytopes of the G x G grid problem.

    Region (i, j) is the box [i-o, i+1+o] x [j-o, j+1+o] with its four corners cut by the lines
    +-x +-y <= const at a distance c ~ U(chamfer) from the corner (m = 8 rows).  Vertex order:
    's', 't', then regions row-major (the reference's key order "s", "t", 0, 1, ...).
    Returns (poly_off, A, b, s_pt, t_pt).
    """
    rng = np.random.default_rng(seed)
    nR = G * G
    ii, jj = np.divmod(np.arange(nR), G)
    o = overlap
    x0 = ii - o + (rng.uniform(-jitter, jitter, nR) if jitter else 0.0)
    x1 = ii + 1 + o + (rng.uniform(-jitter, jitter, nR) if jitter else 0.0)
    y0 = jj - o + (rng.uniform(-jitter, jitter, nR) if jitter else 0.0)
    y1 = jj + 1 + o + (rng.uniform(-jitter, jitter, nR) if jitter else 0.0)
    c = rng.uniform(chamfer[0], chamfer[1], size=(nR, 4))
    r = np.sqrt(0.5)
    A = np.zeros((nR, 8, 2))
    b = np.zeros((nR, 8))
    A[:, 0] = (1, 0); b[:, 0] = x1
    A[:, 1] = (-1, 0); b[:, 1] = -x0
    A[:, 2] = (0, 1); b[:, 2] = y1
    A[:, 3] = (0, -1); b[:, 3] = -y0
    A[:, 4] = (r, r); b[:, 4] = r * (x1 + y1 - c[:, 0])
    A[:, 5] = (-r, r); b[:, 5] = r * (-x0 + y1 - c[:, 1])
    A[:, 6] = (r, -r); b[:, 6] = r * (x1 - y0 - c[:, 2])
    A[:, 7] = (-r, -r); b[:, 7] = r * (-x0 - y0 - c[:, 3])
    s_pt = np.array([0.5, 0.5])
    t_pt = np.array([G - 0.5, G - 0.5])
    As_, bs_ = convert_pt_to_polytope(s_pt)
    At_, bt_ = convert_pt_to_polytope(t_pt)
    Aall = np.concatenate([As_, At_, A.reshape(-1, 2)])
    ball = np.concatenate([bs_, bt_, b.reshape(-1)])
    off = np.concatenate([[0, 4, 8], 8 + 8 * np.arange(1, nR + 1)]).astype(np.int64)
    return off, Aall, ball, s_pt, t_pt


def grid_packed_graph(G, **kw):
    """Grid problem -> PackedGraph (graph discovered by the generic overlap builder)."""
    off, A, b, s_pt, t_pt = grid_problem(G, **kw)
    tail, head = build_graph_packed(off, A, b)
    return PackedGraph(off, A, b, tail, head, 0, 1)


def packed_to_dicts(off, A, b):
    """(poly_off, A, b) with 's','t' first -> the reference's ``As``/``bs`` dicts."""
    keys = ["s", "t"] + list(range(off.shape[0] - 3))
    As = {k: A[off[i]:off[i + 1]] for i, k in enumerate(keys)}
    bs = {k: b[off[i]:off[i + 1]] for i, k in enumerate(keys)}
    return As, bs


def generate_test_2D(filename, low_bound, high_bound, resolution, spacing_factor, num_sets, seed=None):
    """Reference ``test_generator.py:16-171`` without Drake; returns (As, bs, s, t) and writes ``filename``
    (None to skip writing)."""
    from scipy.spatial import ConvexHull, cKDTree
    from scipy.stats.qmc import LatinHypercube
    from .problem_io import write_test_file
    rng = np.random.default_rng(seed)
    gsz = int((high_bound - low_bound) / resolution)
    xs = np.linspace(low_bound, high_bound, gsz)
    X, Y = np.meshgrid(xs, xs)
    grid_points = np.stack([X.ravel(), Y.ravel()], axis=1)
    seeds = LatinHypercube(d=2, optimization="lloyd", seed=rng).random(n=num_sets)
    seeds = (high_bound - low_bound) * seeds + low_bound
    dist, _ = cKDTree(seeds).query(seeds, k=2)
    radius = dist[:, 1] * spacing_factor
    tree = cKDTree(grid_points)
    As, bs = {}, {}
    for i, (sd, rad) in enumerate(zip(seeds, radius)):
        hull = None
        frac = 0.3
        while hull is None:
            close = grid_points[tree.query_ball_point(sd, rad)]
            k = int(frac * len(close))
            if len(close) >= 3 and k > 3:
                pts = close[rng.choice(len(close), size=k, replace=False)]
                try:
                    hcand = ConvexHull(pts)
                    if hcand.volume > 1e-5:
                        hull = hcand
                except Exception:
                    hull = None
            if hull is None:
                rad *= 1.05
                frac = 0.1
        eq = hull.equations                     # rows [a, c] with a.x + c <= 0, |a| = 1
        As[i] = eq[:, :2].copy()
        bs[i] = -eq[:, 2].copy()

    def sample_in(A, b):
        while True:
            p = rng.uniform(low_bound, high_bound, size=2)
            if np.all(A @ p <= b):
                return p
    a, c = rng.choice(num_sets, size=2, replace=False)
    s_pt, t_pt = sample_in(As[a], bs[a]), sample_in(As[c], bs[c])
    A_s, b_s = convert_pt_to_polytope(s_pt)
    A_t, b_t = convert_pt_to_polytope(t_pt)
    As = {"s": A_s, "t": A_t, **As}
    bs = {"s": b_s, "t": b_t, **bs}
    if filename:
        write_test_file(filename, As, bs, s=s_pt, t=t_pt, N=int(num_sets / 5), M=int(2 * num_sets / 5),
                        header=f"generated 2-D GCS problem: {num_sets} regions in [{low_bound}, {high_bound}]^2\n")
    return As, bs, s_pt, t_pt
