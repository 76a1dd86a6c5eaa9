"""Host graph layer vs the reference's behaviour (reference utils.py:31-82)."""
import numpy as np
import pytest
from scipy.optimize import linprog

from conftest import ALL_PROBLEMS, load_golden
from gcs_admm_b200.graph import build_graph, pack_graph, polygon_vertices, convert_pt_to_polytope, delta

SIZES = {"benchmark1": (6, 12), "benchmark2": (10, 28), "benchmark3": (22, 76), "benchmark4": (42, 94)}


def lp_overlap(A1, b1, A2, b2):
    r = linprog(np.zeros(A1.shape[1]), A_ub=np.vstack([A1, A2]), b_ub=np.concatenate([b1, b2]),
                bounds=[(None, None)] * A1.shape[1], method="highs")
    return r.status == 0


@pytest.mark.parametrize("name", ALL_PROBLEMS)
def test_edges_match_lp_feasibility(name):
    """The reference decides overlap with one LP per ordered pair; same edge list, same order."""
    As, bs, n, _, keys = load_golden(name)
    V, E, I_in, I_out = build_graph(As, bs)
    assert V == keys
    expect = [(v1, v2) for v1 in V for v2 in V if v1 != v2 and lp_overlap(As[v1], bs[v1], As[v2], bs[v2])]
    assert E == expect
    for v in V:
        assert I_out[v] == [e for e in E if e[0] == v]
        assert I_in[v] == [e for e in E if e[1] == v]
    if name in SIZES:
        assert (len(V), len(E)) == SIZES[name]


def test_half_edge_layout():
    As, bs, n, _, keys = load_golden("benchmark1")
    V, E, I_in, I_out = build_graph(As, bs)
    g = pack_graph(As, bs, V, E)
    assert g.nV == 6 and g.nE == 12 and g.H == 24
    for vi, v in enumerate(V):
        hs = range(g.he_off[vi], g.he_off[vi + 1])
        edges = [E[g.he_edge[h]] for h in hs]
        assert edges == I_in[v] + I_out[v]            # reference admm_solver_v3.py:105-116 order
        assert [int(g.he_out[h]) for h in hs] == [0] * len(I_in[v]) + [1] * len(I_out[v])
    for e in range(g.nE):
        assert g.he_owner[g.edge_he_tail[e]] == g.edge_tail[e] and g.he_out[g.edge_he_tail[e]] == 1
        assert g.he_owner[g.edge_he_head[e]] == g.edge_head[e] and g.he_out[g.edge_he_head[e]] == 0
    c = g.interior_points()
    off = g.poly_off
    for v in range(g.nV):
        assert np.all(g.polyA[off[v]:off[v + 1]] @ c[v] < g.polyb[off[v]:off[v + 1]])


def test_point_box_and_delta():
    A, b = convert_pt_to_polytope(np.array([2.0, 1.0]))
    assert A.shape == (4, 2) and np.allclose(b, [2 + 1e-6, 1 + 1e-6, -2 + 1e-6, -1 + 1e-6])
    assert delta('s', 's') == 1 and delta('t', 't') == 1 and delta('s', 't') == 0 and delta(0, 0) == 0
    P = polygon_vertices(A, b)
    assert P.shape == (4, 2)


def test_touching_and_disjoint_squares():
    sq = lambda x0, y0: (np.array([[-1., 0], [1, 0], [0, -1], [0, 1]]), np.array([-x0, x0 + 1, -y0, y0 + 1]))
    As, bs = {}, {}
    for k, (x, y) in enumerate([(0, 0), (1, 0), (2.5, 0), (0.5, 0.5)]):
        As[k], bs[k] = sq(x, y)
    V, E, _, _ = build_graph(As, bs)
    # 0-1 touch along an edge (closed sets overlap, as in the LP test), 2 is isolated from 0
    assert (0, 1) in E and (1, 0) in E and (0, 3) in E and (1, 3) in E
    assert (0, 2) not in E and (2, 0) not in E
