"""Pins the CPU oracles against the reference's own stored runs
(benchmark_data/admm_solver_v3_benchmark{1..4}.pkl, exported to tests/golden/)."""
import numpy as np
import pytest

from conftest import load_golden
from gcs_admm_b200.graph import pack_graph

# MOSEK-vs-exact noise in the stored residual sequences is ~1e-4 relative
GOLD = {"benchmark1": 39, "benchmark2": 100, "benchmark3": 508, "benchmark4": 465}


@pytest.mark.parametrize("name", list(GOLD))
def test_c_oracle_replays_reference_run(name):
    from c_oracle import COracle
    As, bs, n, d, keys = load_golden(name)
    g = pack_graph(As, bs)
    o = COracle(g)
    o.run()
    info = o.info()
    assert info["inner_fail"] == 0
    assert info["opt"] and info["it"] == int(d["v3_iterations"]) == GOLD[name]      # same stopping iteration
    rho, pri, dual = o.history()
    assert np.all(rho == d["v3_rho_seq"])                                             # rho never adapts in the stored runs
    scale = max(1.0, float(np.max(d["v3_pri_res_seq"])))
    assert np.max(np.abs(pri - d["v3_pri_res_seq"])) < 1e-4 * scale
    assert np.max(np.abs(dual - d["v3_dual_res_seq"])) < 1e-4 * scale
    assert abs(o.cost() - float(d["v3_cost"])) <= 1e-4 * float(d["v3_cost"])        # north-star tolerance
    _, _, y_v = o.solution()
    assert np.max(np.abs(y_v - d["v3_y_v"])) < 2e-3


def test_numpy_oracle_replays_benchmark1():
    """Independent literal restatement (all 9+9d variables per vertex, dense IPM)."""
    from admm_v3_oracle import OracleADMM
    As, bs, n, d, keys = load_golden("benchmark1")
    g = pack_graph(As, bs)
    o = OracleADMM(g).run()
    assert o.opt and o.it == 39
    assert np.max(np.abs(np.array(o.pri_seq) - d["v3_pri_res_seq"])) < 5e-4
    assert np.max(np.abs(np.array(o.dual_seq) - d["v3_dual_res_seq"])) < 5e-4
    assert abs(o.cost() - float(d["v3_cost"])) <= 1e-4 * float(d["v3_cost"])


def test_c_oracle_matches_literal_numpy_oracle_per_iteration():
    """Reduced (null-space) formulation in C == literal formulation in numpy, iterate by iterate."""
    from admm_v3_oracle import OracleADMM
    from c_oracle import COracle
    As, bs, n, d, keys = load_golden("test3")
    g = pack_graph(As, bs)
    a, b = OracleADMM(g), COracle(g)
    for _ in range(8):
        a.step()
        b.step(1)
        xc, mu, z = b.state()
        # two different IPMs stopped at ~1e-9 gap agree to ~1e-5 on the iterates
        assert np.max(np.abs(xc - a.xc)) < 5e-5
        assert np.max(np.abs(z - a.z)) < 5e-5
        assert np.max(np.abs(mu - a.mu)) < 2e-4


def test_fixed_point_is_classic_relaxation_optimum():
    """Run past the reference's loose stop: the ADMM fixed point is the convex relaxation's
    optimum, which the reference's classic_solver pickles hold (3.000398 for benchmark1)."""
    from c_oracle import COracle
    As, bs, n, d, keys = load_golden("benchmark1")
    g = pack_graph(As, bs)
    o = COracle(g)
    o.step(400, check_stop=False)
    rho, pri, dual = o.history()
    assert pri[-1] < 1e-6 and dual[-1] < 1e-6
    assert abs(o.cost() - float(d["classic_cost"])) < 1e-5
