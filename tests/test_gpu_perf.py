"""GPU: `perf` mode of K1 (inexact x-update, K closed-form splitting iterations).  Its trajectory is not the
reference's; it is validated (a) against the CPU emulation of the same kernel source, iterate by iterate, and
(b) at convergence against the classic relaxation optimum stored in the reference's pickles."""
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden
from gcs_admm_b200.graph import pack_graph

pytestmark = pytest.mark.gpu


def _cost(s):
    x_v, z_v, y_v, z_e = s.solution()
    return float(np.sum(np.linalg.norm(z_v[:, :2] - z_v[:, 2:], axis=1)) + 1e-4 * np.sum(z_e[:, 4]))


from path_utils import gold_path as _gold_path, hausdorff as _hausdorff, polyline as _polyline  # noqa: E402


PERF_GATE = ["test1", "test2", "test3", "test_autogen1", "test_autogen2", "benchmark1", "benchmark2", "benchmark3", "benchmark4"]
# vertex labels of the rounded path are compared where the optimum curve has ONE labelling; benchmark3's stored v3 and
# classic results already label the last stretch differently (4-14-t vs 4-18-t: the segment lies in all three regions),
# so there the gate is the curve itself
SAME_LABELS = {"benchmark1", "benchmark4"}
# benchmark3's relaxation is loose (s->8 0.53 / s->6 0.47 and more splits downstream): with the reference's N=5 / M=20 the
# seeded walk reaches the optimal curve for seeds 1, 3, 4 but not 0 or 2 (the reference's own unseeded run was one lucky
# draw); the gate searches wider there so that its result does not depend on the seed
ROUND_KW = {"benchmark3": dict(N=20, M=100)}


@pytest.mark.parametrize("name", PERF_GATE)
def test_perf_mode_parity_gate(name):
    """The contract of the perf mode (north_star tolerances) under its defined stop rule max(pri, dual) < PERF_ABS_TOL:
    relaxed cost within 1e-4 relative of the classic optimum; rounded result = the reference's stored one (final cost
    within 1e-4 relative — measured ~1e-8 —, waypoints within 1e-3 as a curve and at the terminals, vertex path identical)."""
    from gcs_admm_b200.classic import solve_classic
    from gcs_admm_b200.solver import solve
    As, bs, n, d, keys = load_golden(name)
    from gcs_admm_b200.solver import PERF_ABS_TOL
    res = solve(As, bs, n, mode="perf", seed=0, rounding_kw=ROUND_KW.get(name))
    assert res["converged"] and not res["diverged"], res["status"]
    assert max(res["status"]["pri_res"], res["status"]["dual_res"]) < PERF_ABS_TOL      # (scaled down for problems shorter than 3: solver.perf_abs_tol)
    ours = solve_classic(As, bs, n, seed=0)                       # Drake-free classic_solver (reference classic_solver.py:47-171)
    assert ours["status"] == "optimal"
    assert abs(res["cost"] - ours["cost"]) <= 1e-4 * ours["cost"], (res["cost"], ours["cost"])
    if "classic_cost" in d.files:
        assert abs(res["cost"] - float(d["classic_cost"])) <= 1e-4 * float(d["classic_cost"]), (res["cost"], float(d["classic_cost"]))
    P = _polyline(res["x_v_rounded"], res["path"])
    if "v3_y_v_rounded" in d.files:
        for tag in ("v3", "classic"):
            gpath, gx, gcost = _gold_path(As, d, keys, tag)
            assert abs(res["final_cost"] - gcost) <= 1e-6 * gcost, (tag, res["final_cost"], gcost)
            assert _hausdorff(P, _polyline(gx, gpath)) <= 1e-3, tag
            for v in ("s", "t"):
                assert np.max(np.abs(np.asarray(res["x_v_rounded"][v]) - gx[v])) <= 1e-3
            if name in SAME_LABELS:       # elsewhere several labellings describe the optimal curve (equal cost, Hausdorff distance ~1e-6)
                assert res["path"] == gpath, (tag, res["path"], gpath)
    else:                                                         # no stored run: our classic solver's rounded result
        assert abs(res["final_cost"] - ours["final_cost"]) <= 1e-6 * max(1.0, ours["final_cost"])
        assert _hausdorff(P, _polyline(ours["x_v_rounded"], ours["path"])) <= 1e-3
    if name == "test2":                                           # the one known answer in the reference (test_data/test2.py:47-52)
        assert res["path"] == ["s", 0, 1, "t"]
        assert np.allclose(res["x_v_rounded"][0], [0, 1, 0.95, 0.05], atol=1e-3) and np.allclose(res["x_v_rounded"][1], [0.95, 0.05, 1.9, 1], atol=1e-3)


@pytest.mark.parametrize("name", ["benchmark1", "benchmark2", "benchmark3", "benchmark4"])
def test_parity_mode_rounds_gpu_flows_to_the_stored_result(name):
    """parity mode (the reference's trajectory and stop rule), flows from the GPU: rounding gives the stored v3 result"""
    from gcs_admm_b200.solver import solve
    As, bs, n, d, keys = load_golden(name)
    res = solve(As, bs, n, seed=0, rounding_kw=ROUND_KW.get(name))
    assert res["iterations"] == int(d["v3_iterations"])
    gpath, gx, gcost = _gold_path(As, d, keys, "v3")
    assert abs(res["final_cost"] - gcost) <= 1e-6 * gcost
    assert _hausdorff(_polyline(res["x_v_rounded"], res["path"]), _polyline(gx, gpath)) <= 1e-3
    if name in SAME_LABELS:
        assert res["path"] == gpath


@pytest.mark.parametrize("name,K,adapt,frames,oa", [("benchmark4", 3, False, "global", 1.0), ("benchmark3", 1, True, "global", 1.0), ("test_autogen2", 2, True, "global", 1.0),
                                                    ("benchmark4", 1, True, "local", 1.0), ("benchmark2", 2, False, "local", 1.0),
                                                    ("benchmark4", 1, True, "local", 1.7), ("benchmark3", 1, False, "global", 1.5)])
def test_perf_kernel_equals_cpu_emulation(name, K, adapt, frames, oa):
    """the CUDA kernel against the same source compiled for the host, iterate by iterate — with the rho adaptation on
    (adapt=True: the lam / mu rescale paths are exercised) and off"""
    import test_perf_mode as T
    from gcs_admm_b200.lib import Solver
    g = pack_graph(*load_golden(name)[:2])
    a = T.EmuPerfADMM(T.load_emu(), g, K=K, adapt=adapt, frames=frames, outer_alpha=oa)
    s = Solver(g, frac=1.0 if adapt else 0.0, max_it=1000, use_graph=0, outer_alpha=oa).enable_perf(inner_iters=K, frames=frames)
    for it in range(60):
        a.step()
        s.step(1)
        xc, mu, z, rho, k = s.state()
        assert rho == a.rho, it
        assert np.max(np.abs(xc - a.xc)) < 1e-9 and np.max(np.abs(z - a.z)) < 1e-9 and np.max(np.abs(mu - a.ms * a.mu)) < 1e-9, it
    t, tn = s.perf_state()
    assert np.max(np.abs(t - a.tstate)) < 1e-9 and np.max(np.abs(tn - a.tn)) < 1e-9
    x_v, z_v, y_v, z_e = s.solution()
    assert np.max(np.abs(z_v - a.z_v)) < 1e-9 and np.max(np.abs(y_v - a.y_v)) < 1e-9
    _, pri, dual = s.history()
    assert np.allclose(pri, a.pri, rtol=1e-9, atol=1e-12) and np.allclose(dual, a.dual, rtol=1e-9, atol=1e-12)
    s.close()


@pytest.mark.parametrize("name", ["benchmark1", "benchmark2", "benchmark4", "test_autogen1"])
def test_local_frames_reach_the_same_fixed_point(name):
    """local frames (perf-mode option): the same optimisation problem in other coordinates — same relaxed cost (1e-4 relative)
    and same rounded curve as the global-frame run, and invariant under a translation of the whole problem"""
    from gcs_admm_b200.solver import solve
    As, bs, n, d, keys = load_golden(name)
    ref = solve(As, bs, n, mode="perf", seed=0)
    loc = solve(As, bs, n, mode="perf", seed=0, frames="local", abs_tol=2e-5)
    assert loc["converged"]
    assert abs(loc["cost"] - ref["cost"]) <= 1e-4 * ref["cost"]
    assert abs(loc["final_cost"] - ref["final_cost"]) <= 1e-6 * ref["final_cost"]
    assert _hausdorff(_polyline(loc["x_v_rounded"], loc["path"]), _polyline(ref["x_v_rounded"], ref["path"])) <= 1e-3
    shift = np.array([250.0, -170.0])
    bs2 = {k: bs[k] + As[k] @ shift for k in As}
    mov = solve(As, bs2, n, mode="perf", seed=0, frames="local", abs_tol=2e-5)
    assert abs(mov["iterations"] - loc["iterations"]) <= 0.02 * loc["iterations"] + 2 and abs(mov["cost"] - loc["cost"]) <= 1e-6 * loc["cost"]


def test_perf_state_round_trip_resumes_the_run():
    """get_state + get_perf_state -> a fresh handle -> the same iterates as the uninterrupted run (checkpoint / resume, and what
    the end-to-end bench leg uploads)"""
    from gcs_admm_b200.lib import Solver
    g = pack_graph(*load_golden("benchmark4")[:2])
    s = Solver(g, max_it=1000, frac=0.0).enable_perf(inner_iters=1)
    s.step(100)
    xc, mu, z, rho, it = s.state()
    t, tn = s.perf_state()
    s.step(50)
    ref = s.state()
    s.close()
    s2 = Solver(g, max_it=1000, frac=0.0).enable_perf(inner_iters=1)
    s2.set_state(xc, mu, z, rho, it)
    s2.set_perf_state(t, tn)
    s2.step(50)
    got = s2.state()
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[2], ref[2]) and got[4] == ref[4]
    s2.close()


def test_grid_fixed_point_equals_our_classic_solver():
    """north_star: configurations without a stored reference run are compared with classic_solver — here OUR Drake-free one."""
    from gcs_admm_b200.classic import solve_classic
    from gcs_admm_b200.generator import grid_problem, packed_to_dicts
    from gcs_admm_b200.lib import Solver
    off, A, b, s_pt, t_pt = grid_problem(6)
    As, bs = packed_to_dicts(off, A, b)
    ref = solve_classic(As, bs, 2, round_solution=False)
    assert ref["status"] == "optimal"
    # local frames: in global coordinates the grid's many zero-cost 2-cycles keep circulating for > 300 000 iterations
    # (their decay is slowed by the |position|^2 "mass" of the perspective variables); centred frames reach the fixed point in ~22 000
    s = Solver(pack_graph(As, bs), max_it=400000, abs_stop=1, abs_tol=3e-5, frac=0.0, check_every=64).enable_perf(inner_iters=1, frames="local")
    st = s.run()
    assert st["converged"] and st["iterations"] < 60000
    assert abs(_cost(s) - ref["cost"]) <= 1e-4 * ref["cost"]
    s.close()


def test_batched_perf_queries_equal_individual_solves():
    """perf mode on a block-diagonal batch: every problem keeps its own residuals / rho / stop and equals its stand-alone run
    (the class tables of the batch are a superset of each problem's, so the arithmetic per vertex is identical)."""
    from gcs_admm_b200.graph import pack_batch
    from gcs_admm_b200.lib import Solver
    names = ["benchmark1", "benchmark2", "test3", "benchmark4"]
    graphs = [pack_graph(*load_golden(n)[:2]) for n in names]
    big = pack_batch(graphs)
    sb = Solver(big, max_it=3000).enable_perf(inner_iters=2)
    stb = sb.run(3000)
    x_v, z_v, y_v, z_e = sb.solution()
    for p, g in enumerate(graphs):
        s = Solver(g, max_it=3000).enable_perf(inner_iters=2)
        st = s.run(3000)
        ps = sb.problem_status(p)
        assert ps["iterations"] == st["iterations"] and ps["converged"] == st["converged"], (names[p], ps, st)
        xs, zs, ys, es = s.solution()
        v0, v1, e0, e1 = big.prob_voff[p], big.prob_voff[p + 1], big.prob_eoff[p], big.prob_eoff[p + 1]
        assert np.allclose(zs, z_v[v0:v1], rtol=0, atol=1e-12) and np.allclose(es, z_e[e0:e1], rtol=0, atol=1e-12)
        s.close()
    sb.close()


def test_inner_residual_is_part_of_the_absolute_stop_rule():
    """perf mode: GcsStatus.inner_res = |(M u + m0) - c| over all (point, flow) pairs (how far the vertex programs' own cone
    constraints are from being met).  It is produced by a second variant of K1 that gcsadmm_run switches to once the consensus
    residuals are within 4x of abs_tol (-1 = not computed in the last iteration), and must be below abs_tol too for the stop."""
    from gcs_admm_b200.lib import Solver
    As, bs, n, d, keys = load_golden("benchmark2")
    g = pack_graph(As, bs)
    s = Solver(g, max_it=400000, abs_stop=1, abs_tol=3e-5, frac=0.0, check_every=64).enable_perf(inner_iters=1)
    s.step(50)
    assert s.status()["inner_res"] == -1.0                                  # throughput variant: not computed, and the run cannot stop
    st = s.run()
    assert st["converged"] == 1 and 0.0 <= st["inner_res"] < 3e-5 and max(st["pri_res"], st["dual_res"]) < 3e-5
    s.close()
    e = Solver(g)
    e.step(3)
    assert e.status()["inner_res"] == 0.0                                   # exact mode: the vertex programs are solved to 1e-8
    e.close()


def test_reference_definition_residuals_in_local_frames_vs_numpy():
    """local frames: the check variant of the edge kernel evaluates the primal residual in GLOBAL coordinates (the reference's
    definition :598) — compared with numpy from the state of the stop iteration.  abs_tol is huge, so the run stops at the first
    iteration the check variant runs (iteration check_every + 1)."""
    from gcs_admm_b200.lib import Solver
    from gcs_admm_b200 import perf
    g = pack_graph(*load_golden("benchmark4")[:2])
    T = perf.perf_tables(g, frames="local")
    s = Solver(g, abs_stop=1, abs_tol=1e9, check_every=8, frac=0.0, max_it=1000).enable_perf(inner_iters=1, tables=T)
    st = s.run(64)
    assert st["converged"] == 1 and st["iterations"] == 9 and st["inner_res"] >= 0.0
    xc, mu, z, rho, it = s.state()
    s.close()
    d, cu = T["edge_delta"], T["edge_cent"]
    cw = cu - d
    xt, xh = xc[g.edge_he_tail], xc[g.edge_he_head]
    bz = z.copy()
    bz[:, 2:4] -= d * z[:, 4:5]
    rt, rh = bz - xt, z - xh                                   # local residuals of the two copies
    assert abs(np.sqrt(np.sum(rt ** 2) + np.sum(rh ** 2)) - st["pri_res"]) < 1e-10 * max(1.0, st["pri_res"])
    gt = rt.copy(); gt[:, 0:2] += rt[:, 4:5] * cu; gt[:, 2:4] += rt[:, 4:5] * cu          # every slot of the tail's copy: the tail's frame
    gh = rh.copy(); gh[:, 0:2] += rh[:, 4:5] * cu; gh[:, 2:4] += rh[:, 4:5] * cw          # head: tail's point in the tail's frame, own point in its own
    ref = np.sqrt(np.sum(gt ** 2) + np.sum(gh ** 2))
    assert abs(ref - st["pri_res_ref"]) < 1e-10 * max(1.0, ref)
    assert st["pri_res_ref"] > st["pri_res"]                   # benchmark4's regions are far from the origin: flow mismatches are amplified


@pytest.mark.parametrize("G", [12, 16, 24, 32, 64])
def test_grid_fixed_point_equals_classic_optimum(G):
    """the scalable benchmark family at sizes the host interior-point comparator can still finish (fixture:
    tests/golden/grid_classic.json, made by tools/gen_grid_classic_golden.py): the perf-mode ADMM with the accelerated
    configuration of the time-to-residual run (local frames, rho0 = 3, over-relaxed consensus step, dual warm start), iterated to
    1e-5, has the classic relaxation's optimal cost to 1e-4 relative"""
    import json
    from gcs_admm_b200.generator import grid_packed_graph
    from gcs_admm_b200.lib import Solver
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "grid_classic.json")))
    if str(G) not in gold:
        pytest.skip(f"no classic optimum stored for G = {G}")
    g = grid_packed_graph(G)
    s = Solver(g, max_it=2_000_000, abs_stop=1, abs_tol=1e-5, check_every=256, frac=100 / 2_000_000, rho0=3.0, outer_alpha=1.7)
    s.enable_perf(inner_iters=1, frames="local").warm_start("dijkstra", rho=3.0)
    st = s.run()
    x_v, z_v, y_v, z_e = s.solution()
    s.close()
    cost = float(np.sum(np.linalg.norm(z_v[:, :2] - z_v[:, 2:], axis=1)) + 1e-4 * np.sum(z_e[:, 4]))
    assert st["converged"] == 1 and max(st["pri_res"], st["dual_res"], st["inner_res"]) < 1e-5
    assert abs(cost - gold[str(G)]["cost"]) <= 1e-4 * gold[str(G)]["cost"], (cost, gold[str(G)]["cost"], st["iterations"])
