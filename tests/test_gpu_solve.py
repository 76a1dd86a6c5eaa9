"""GPU: the public entry points (solve(), the CLI) and size-independent properties at BASELINE sizes."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu


def test_solve_entry_point_benchmark1():
    from gcs_admm_b200.solver import solve
    As, bs, n, d, keys = load_golden("benchmark1")
    res = solve(As, bs, n, seed=0)
    assert res["iterations"] == 39 and res["converged"]
    assert abs(res["cost"] - float(d["v3_cost"])) <= 1e-4 * float(d["v3_cost"])
    assert len(res["pri_res_seq"]) == 40 and res["pri_res_seq"][0] == 0.0 and res["rho_seq"][0] == 1.0
    assert res["path"][0] == "s" and res["path"][-1] == "t" and set(res["path"]) in ({'s', 0, 1, 2, 't'}, {'s', 0, 3, 2, 't'})
    gold_len = sum(np.linalg.norm(x[:2] - x[2:]) for x, y in zip(d["v3_x_v_rounded"], d["v3_y_v_rounded"]) if y > 0.5)
    assert abs(res["final_cost"] - gold_len) < 1e-4 * gold_len
    assert list(res["y_v_sol"].keys()) == keys and list(res["y_e_sol"].keys()) == res["E"]


def test_cli_writes_reference_pickle(tmp_path):
    env = dict(os.environ)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "admm_solver_v3.py"), "--test_file=benchmark2", "--show_plot=False", "--seed=0"],
                         capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=600)
    assert out.returncode == 0, out.stderr
    for landmark in ("Running ADMM Solver v3 on benchmark2", "V: ['s', 't', 0", "E: [", "it = 100/1000", "BREAKING FOR OPT",
                     "Total solve time:", "Cost before rounding:", "POST-ROUNDING", "x_v_rounded=", "y_v_rounded="):
        assert landmark in out.stdout, landmark
    d = pickle.load(open(os.path.join(ROOT, "benchmark_data", "admm_solver_v3_benchmark2.pkl"), "rb"))
    gold = load_golden("benchmark2")[3]
    assert d["iterations"] == 100 and d["ADMM"] is True
    assert abs(d["cost"] - float(gold["v3_cost"])) <= 1e-4 * float(gold["v3_cost"])
    assert len(d["pri_res_seq"]) == 101 and isinstance(d["rho_seq"], np.ndarray)
    bad = subprocess.run([sys.executable, os.path.join(ROOT, "admm_solver_v3.py"), "--test_file=nope"], capture_output=True, text=True, timeout=120)
    assert bad.returncode == 1 and "Error: Test file 'nope' not found" in bad.stdout


def test_small_grid_converges_to_a_sane_path():
    """Run a 5x5 grid towards small residuals (ADMM at rho ~ 1 converges slowly, reference report section V):
    relaxation cost ~ straight-line distance (the straight segment is feasible here), rounded path cost >=
    relaxation cost, path connects s to t through overlapping regions."""
    from gcs_admm_b200.generator import grid_problem, packed_to_dicts
    from gcs_admm_b200.solver import solve
    off, A, b, s_pt, t_pt = grid_problem(5)
    As, bs = packed_to_dicts(off, A, b)
    res = solve(As, bs, 2, max_it=3000, abs_stop=1, abs_tol=5e-4, seed=0)
    assert not res["diverged"] and max(res["pri_res_seq"][-1], res["dual_res_seq"][-1]) < 2e-3
    straight = float(np.linalg.norm(t_pt - s_pt))
    assert abs(res["cost"] - straight) < 5e-3 * straight
    assert res["final_cost"] >= straight - 1e-6 and res["final_cost"] < 1.5 * straight
    p = res["path"]
    assert p[0] == "s" and p[-1] == "t" and all((a, c) in set(res["E"]) for a, c in zip(p[:-1], p[1:]))


def test_properties_at_full_size():
    """100k-vertex grid (the BASELINE metric config), a few iterations: exact identities that do not need an
    oracle — z is the average of the two copies, mu accumulates z - xc, reruns are bitwise reproducible."""
    from gcs_admm_b200.generator import grid_packed_graph
    from gcs_admm_b200.lib import Solver
    g = grid_packed_graph(316)
    assert g.nV == 99858 and g.nE == 398164
    outs = []
    for rep in range(2):
        s = Solver(g)
        s.step(6)
        xc, mu, z, rho, it = s.state()
        outs.append((xc, mu, z))
        if rep == 0:
            assert it == 6 and np.all(np.isfinite(xc)) and np.all(np.isfinite(mu))
            assert np.array_equal(z, 0.5 * (xc[g.edge_he_tail] + xc[g.edge_he_head]))
            mu_prev = mu.copy()
            s.step(1)
            xc2, mu2, z2, _, _ = s.state()
            assert np.allclose(mu2, mu_prev + (z2[g.he_edge] - xc2), rtol=0, atol=1e-15)
            st = s.status()
            r = z2[g.he_edge] - xc2
            assert abs(st["pri_res"] - np.sqrt(np.sum(r * r))) <= 1e-9 * max(1.0, st["pri_res"])
        s.close()
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))


@pytest.mark.parametrize("G,burn", [(100, 110), (316, 326)])
def test_bench_workload_matches_oracle_per_iteration(G, burn):
    """The bench configurations themselves (BASELINE config 3 and the metric's 100k-vertex grid), pinned to the oracle: after the
    bench's burn-in on the GPU the state is injected into the C oracle and both advance two iterations, glued."""
    from c_oracle import COracle, use_all_cores
    from gcs_admm_b200.generator import grid_packed_graph
    from gcs_admm_b200.lib import Solver
    use_all_cores()
    g = grid_packed_graph(G)
    s = Solver(g, max_it=1000, eps_abs=0.0, eps_rel=0.0)
    s.step(burn)
    xc, mu, z, rho, it = s.state()
    assert s.status()["skipped"] < burn * g.nV          # the cold-start wave has passed: vertex programs are real solves now
    o = COracle(g, max_it=1000, eps_abs=0.0, eps_rel=0.0)
    o.set_state(xc, mu, z, rho, it)
    for k in range(2):
        s.step(1)
        o.step(1)
        xg, mg, zg, rho_g, it_g = s.state()
        xo, mo, zo = o.state()
        # 1e-3 = the north-star waypoint tolerance: at this size a few vertex programs have nearly flat directions, where the two
        # interior-point solvers (tolerance 1e-8 / 1e-9) stop up to 3e-4 apart; the residual norms below agree to 1e-4
        assert np.max(np.abs(xg - xo)) < 1e-3 and np.max(np.abs(zg - zo)) < 1e-3, (G, k)
        assert rho_g == o.info()["rho"]
        s.set_state(xo, mo, zo, rho=rho_g, it=it_g)
    _, p1, d1 = s.history()
    _, p2, d2 = o.history()
    assert abs(p1[-1] - p2[-1]) < 1e-4 * max(1.0, p2[-1]) and abs(d1[-1] - d2[-1]) < 1e-4 * max(1.0, d2[-1])
    assert s.status()["inner_fail"] <= 1e-4 * s.status()["iterations"] * g.nV      # solves that ended at the fp64 noise floor instead of the tolerance
    s.close()


def test_divergence_is_a_status_not_an_exception():
    """reference :662-664 / :679-681: non-finite iterates break the loop, the script still reports and pickles.  Here:
    run() returns normally with diverged = 1 and the last iterates can be read back."""
    from gcs_admm_b200.graph import pack_graph
    from gcs_admm_b200.lib import Solver
    As, bs, n, d, keys = load_golden("benchmark2")
    g = pack_graph(As, bs)
    s = Solver(g)
    s.step(3)
    xc, mu, z, rho, it = s.state()
    mu[0, 0] = np.nan
    s.set_state(xc, mu, z, rho, it)
    st = s.run()
    assert st["diverged"] == 1 and st["converged"] == 0 and st["iterations"] == it + 1
    x_v, z_v, y_v, z_e = s.solution()
    r, p, dl = s.history()
    assert len(p) == it + 2 and not np.isfinite(p[-1])
    s.close()


def test_over_relaxed_consensus_step_vs_numpy():
    """outer_alpha != 1 (perf-mode option): z and mu use alpha xc + (1 - alpha) z_old, the primal residual the true xc"""
    from gcs_admm_b200.graph import pack_graph
    from gcs_admm_b200.lib import Solver
    g = pack_graph(*load_golden("benchmark4")[:2])
    rng = np.random.default_rng(2)
    xc, mu0, z0 = rng.normal(size=(g.H, 5)), rng.normal(size=(g.H, 5)), rng.normal(size=(g.nE, 5))
    al = 1.7
    s = Solver(g, outer_alpha=al, frac=0.0)
    s.set_state(xc, mu0, z0, rho=1.0, it=0)
    s.edge_update()
    s.control()
    xc1, mu1, z1, rho, it = s.state()
    hat = al * xc + (1 - al) * z0[g.he_edge]
    z_ref = 0.5 * (hat[g.edge_he_tail] + hat[g.edge_he_head])
    assert np.allclose(z1, z_ref, rtol=0, atol=1e-15)
    assert np.allclose(mu1, mu0 + z_ref[g.he_edge] - hat, rtol=0, atol=1e-14)
    st = s.status()
    assert abs(st["pri_res"] - np.sqrt(np.sum((z_ref[g.he_edge] - xc) ** 2))) < 1e-12 * st["pri_res"]
    s.close()
