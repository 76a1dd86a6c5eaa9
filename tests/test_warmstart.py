"""Dual warm start of the perf mode (gcs_admm_b200/warmstart.py): the cost-to-go field over the portal graph and the invariants the
start has to respect so that the fixed point of the iteration is not shifted."""
import numpy as np
import pytest

from conftest import load_golden
from gcs_admm_b200 import perf, warmstart
from gcs_admm_b200.generator import grid_packed_graph
from gcs_admm_b200.graph import pack_graph


def test_portal_field_on_an_open_grid_is_the_straight_line_distance():
    g = grid_packed_graph(12)
    Jd, gd = warmstart.cost_to_go(g, "dijkstra")
    Je, ge = warmstart.cost_to_go(g, "euclid")
    live = ((g.he_flags[g.edge_he_head] | g.he_flags[g.edge_he_tail]) & 2) == 0
    far = live & (Je > 3.0)
    # portal-to-portal paths are staircases through the midpoints of the overlaps: never shorter than the straight line, and not
    # much longer (an edge that points away from the target has to come back: no immediate U-turn)
    assert np.all(Jd[far] >= Je[far] - 1e-9) and np.all(Jd[far] <= 1.1 * Je[far] + 2.5)
    assert np.median(Jd[far] / Je[far]) < 1.12
    assert np.allclose(np.linalg.norm(gd[far], axis=1), 1.0)
    assert np.mean(np.sum(gd[far] * ge[far], axis=1)) > 0.3          # staircase directions (next portal) vs the straight line: same half-plane on average
    # the field decreases along its own gradient direction: J(next portal) < J(portal)
    assert Jd[g.edge_head == g.dst].max() < 1.0


@pytest.mark.parametrize("frames", ["local", "global"])
def test_dual_start_keeps_the_invariant_of_the_consensus_step(frames):
    """B' mu_tail + mu_head = 0 (B = identity in global frames) is conserved by the iteration, so the start must satisfy it;
    forced-zero edges start with zero duals on both sides; the tail's first point has no dual (it is a free copy at the head)"""
    g = pack_graph(*load_golden("benchmark4")[:2])
    T = perf.perf_tables(g, frames=frames)
    d = T["edge_delta"]
    mu = warmstart.dual_start(g, d, rho=3.0)
    mt, mh = mu[g.edge_he_tail], mu[g.edge_he_head]
    inv = mt + mh
    if d is not None:
        inv[:, 4] -= np.sum(d * mt[:, 2:4], axis=1)
    assert np.max(np.abs(inv)) < 1e-12
    assert np.all(mt[:, :2] == 0) and np.all(mh[:, :2] == 0)
    dead = ((g.he_flags[g.edge_he_head] | g.he_flags[g.edge_he_tail]) & 2) != 0
    assert dead.any() and np.all(mt[dead] == 0) and np.all(mh[dead] == 0)
    assert np.isfinite(mu).all()
    # scaled duals: halving rho doubles them
    assert np.allclose(warmstart.dual_start(g, d, rho=1.5), 2 * mu)


def test_unreachable_edges_get_zero_duals():
    """benchmark4's region graph is not strongly connected: edges that cannot reach the target must not carry inf / nan"""
    g = pack_graph(*load_golden("benchmark4")[:2])
    J, grad = warmstart.cost_to_go(g, "dijkstra")
    assert np.isfinite(J).all() and np.isfinite(grad).all()


def test_dual_start_sliced_to_a_partition_keeps_the_invariant_in_local_indices():
    """multi-GPU time-to-residual run (dist_bench.time_to_residual): every rank takes mu_global[lp.global_he]; for every local edge
    whose two half-edges are owned the invariant must hold with the rank's OWN indices and its own slice of edge_delta"""
    from gcs_admm_b200.partition import partition_vertices, split_graph
    g = grid_packed_graph(10)
    T = perf.perf_tables(g, frames="local")
    mu = warmstart.dual_start(g, T["edge_delta"], rho=3.0)
    for lp in split_graph(g, partition_vertices(g, 3), 3):
        ml = mu[np.asarray(lp.global_he, dtype=np.int64)]
        Tl = perf.local_tables(T, lp)
        d = Tl["edge_delta"]
        assert np.allclose(Tl["edge_cent"], T["edge_cent"][np.asarray(lp.global_edges, dtype=np.int64)])
        nH = int(lp.he_off[-1])
        both = (lp.edge_he_tail < nH) & (lp.edge_he_head < nH)
        assert both.any() and (~both).any()                      # interior edges and cut edges
        mt, mh = ml[lp.edge_he_tail[both]], ml[lp.edge_he_head[both]]
        inv = mt + mh
        inv[:, 4] -= np.sum(d[both] * mt[:, 2:4], axis=1)
        assert np.max(np.abs(inv)) < 1e-12
        # a cut edge: the owned side carries the same dual as in the global array
        cut_t = (lp.edge_he_tail < nH) & ~both
        ge = np.asarray(lp.global_edges, dtype=np.int64)[cut_t]
        assert np.allclose(ml[lp.edge_he_tail[cut_t]], mu[g.edge_he_tail[ge]])
