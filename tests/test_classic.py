"""Host: Drake-free classic solver (reference classic_solver.py) against the optimum stored in the
reference's pickles (tests/golden, keys classic_*), and as the comparator of the ADMM fixed point."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden
from gcs_admm_b200.classic import solve_classic


@pytest.mark.parametrize("name,tol", [("benchmark1", 1e-8), ("benchmark2", 1e-7), ("benchmark3", 1e-4), ("benchmark4", 1e-4)])
def test_classic_optimum_matches_reference_pickle(name, tol):
    """north_star tolerance: cost within 1e-4 relative.  (benchmark3/4: the stored relaxation value is ABOVE the cost of
    its own rounded path + edge penalties, i.e. the reference's solver stopped early; ours is below it and tight.)"""
    As, bs, n, d, keys = load_golden(name)
    r = solve_classic(As, bs, n, seed=0)
    gold = float(d["classic_cost"])
    assert r["status"] == "optimal"
    assert abs(r["cost"] - gold) <= tol * gold
    # relaxation is a lower bound of the integral solution it rounds to (+ the 1e-4 per-edge penalty on its path)
    n_edges = len(r["path"]) - 1
    assert r["cost"] <= r["final_cost"] + 1e-4 * n_edges + 1e-7
    # rounded solution = the reference's rounded solution
    gold_on = {k for k, y in zip(keys, d["classic_y_v_rounded"]) if y > 0.5}
    gold_len = sum(np.linalg.norm(x[:2] - x[2:]) for x, y in zip(d["classic_x_v_rounded"], d["classic_y_v_rounded"]) if y > 0.5)
    assert abs(r["final_cost"] - gold_len) <= 1e-6 * gold_len
    # same curve; the vertex labels can differ where a stretch lies in several overlapping regions (benchmark2: s-5-0 vs s-2-0,
    # benchmark3: 4-t vs 4-18-t) — see the tie policy in gcs_admm_b200.rounding.rounding
    from path_utils import gold_path, hausdorff, polyline
    gpath, gx, _ = gold_path(As, d, keys, "classic")
    assert hausdorff(polyline(r["x_v_rounded"], r["path"]), polyline(gx, gpath)) <= 1e-3
    if name in ("benchmark1", "benchmark4"):
        assert r["path"] == gpath
    # flows: conservation and bounds
    y_e, y_v = r["y_e_sol"], r["y_v_sol"]
    assert all(-1e-9 <= y <= 1 + 1e-9 for y in y_e.values())
    assert abs(y_v["s"] - 1) < 1e-8 and abs(y_v["t"] - 1) < 1e-8


def test_classic_is_the_admm_fixed_point_value():
    """The C oracle's ADMM run long on benchmark1 reaches the classic optimum (same check as test_oracle_golden, but
    against OUR classic solver instead of the stored number)."""
    As, bs, n, d, keys = load_golden("benchmark1")
    r = solve_classic(As, bs, n, round_solution=False)
    assert abs(r["cost"] - 3.000398) < 1e-6


def test_classic_cli_writes_reference_pickle(tmp_path):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "classic_solver.py"), "--test_file=benchmark1", "--show_plot=False", "--seed=0"],
                         capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert out.returncode == 0, out.stderr
    for landmark in ("Running Classic Solver on benchmark1", "V: ['s', 't', 0", "E: [", "Beginning MICP Solve.", "Solve Time:", "Solved using:",
                     "Optimal Cost Pre-rounding (Path Length): 3.0003", "POST-ROUNDING", "x_v_rounded=", "y_v_rounded="):
        assert landmark in out.stdout, landmark
    rec = pickle.load(open(os.path.join(ROOT, "benchmark_data", "classic_solver_benchmark1.pkl"), "rb"))
    assert rec["ADMM"] is False and "iterations" not in rec and abs(rec["cost"] - 3.000398) < 1e-5
    bad = subprocess.run([sys.executable, os.path.join(ROOT, "classic_solver.py"), "--test_file=nope"], capture_output=True, text=True, timeout=120)
    assert bad.returncode == 1 and "Error: Test file 'nope' not found" in bad.stdout


def test_classic_on_a_generated_grid():
    from gcs_admm_b200.generator import grid_problem, packed_to_dicts
    off, A, b, s_pt, t_pt = grid_problem(4)
    As, bs = packed_to_dicts(off, A, b)
    r = solve_classic(As, bs, 2, seed=0)
    straight = float(np.linalg.norm(t_pt - s_pt))
    assert r["status"] == "optimal" and straight - 1e-7 <= r["cost"] <= straight + 1e-4 * len(r["E"])
    assert straight - 1e-6 <= r["final_cost"] < 1.5 * straight          # the sampled walks need not contain the diagonal
