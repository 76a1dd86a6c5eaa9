"""GPU: `perf` mode of K1 (inexact x-update, K closed-form splitting iterations).  Its trajectory is not the
reference's; it is validated (a) against the CPU emulation of the same kernel source, iterate by iterate, and
(b) at convergence against the classic relaxation optimum stored in the reference's pickles."""
import numpy as np
import pytest

from conftest import load_golden
from gcs_admm_b200.graph import pack_graph

pytestmark = pytest.mark.gpu


def _cost(s):
    x_v, z_v, y_v, z_e = s.solution()
    return float(np.sum(np.linalg.norm(z_v[:, :2] - z_v[:, 2:], axis=1)) + 1e-4 * np.sum(z_e[:, 4]))


@pytest.mark.parametrize("name,K,iters,tol", [("benchmark1", 3, 500, 1e-5), ("benchmark2", 3, 3000, 1e-4), ("benchmark4", 3, 6000, 2e-2),
                                              ("benchmark1", 1, 1500, 1e-5), ("benchmark2", 1, 8000, 1e-4), ("benchmark4", 1, 15000, 2e-2)])
def test_perf_mode_converges_to_classic_optimum(name, K, iters, tol):
    from gcs_admm_b200.lib import Solver
    As, bs, n, d, keys = load_golden(name)
    s = Solver(pack_graph(As, bs), max_it=iters + 10, eps_abs=0.0, eps_rel=0.0).enable_perf(inner_iters=K)
    s.step(iters)
    st = s.status()
    assert not st["diverged"] and np.isfinite(st["pri_res"])
    assert abs(_cost(s) - float(d["classic_cost"])) <= tol * float(d["classic_cost"])
    s.close()


def test_perf_kernel_equals_cpu_emulation():
    import test_perf_mode as T
    import ctypes as C, os, subprocess
    so = os.path.join(T.CSRC, "libgcsemu.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-o", so, os.path.join(T.CSRC, "emulate.cpp")])
    lib = C.CDLL(so)
    lib.gcsemu_vertex_update_perf_all.restype = C.c_int
    lib.gcsemu_vertex_update_perf_all.argtypes = [C.c_int, C.c_int, T._ip, T._dp, T._dp, T._ip, T._ip, T._bp, T._bp, T._dp, T._dp, T._dp, T._dp, T._dp, T._dp, T._dp,
                                                  C.c_double, C.c_double, C.c_int, T._ip, T._ip, T._dp, T._ip, T._dp, T._dp, C.c_int, C.c_double, C.c_double]
    lib.gcsemu_perf_state_stride.restype = C.c_int
    from gcs_admm_b200.lib import Solver
    g = pack_graph(*load_golden("benchmark4")[:2])
    a = T.EmuPerfADMM(lib, g, K=3)
    s = Solver(g, frac=0.0).enable_perf(inner_iters=3)          # frac = 0: no rho adaptation, like the emulation driver
    for it in range(30):
        a.step()
        s.step(1)
        xc, mu, z, rho, k = s.state()
        assert np.max(np.abs(xc - a.xc)) < 1e-9 and np.max(np.abs(z - a.z)) < 1e-9, it
    s.close()


def test_perf_mode_rounds_to_the_reference_path():
    """After convergence the flows of the inexact mode round to the same vertex path as the reference's stored run."""
    from gcs_admm_b200.graph import build_graph
    from gcs_admm_b200.lib import Solver
    from gcs_admm_b200.rounding import rounding
    As, bs, n, d, keys = load_golden("benchmark4")
    V, E, I_in, I_out = build_graph(As, bs)
    s = Solver(pack_graph(As, bs, V, E), max_it=3010, eps_abs=0.0, eps_rel=0.0).enable_perf(inner_iters=3)
    s.step(3000)
    _, _, _, z_e = s.solution()
    y_e = {e: float(z_e[i, 4]) for i, e in enumerate(E)}
    cost, x_r, y_r, path = rounding(y_e, V, E, I_out, As, bs, n, rng=0, return_path=True)
    gold_on = {k for k, y in zip(keys, d["v3_y_v_rounded"]) if y > 0.5}
    assert set(path) == gold_on
    s.close()


def test_grid_fixed_point_equals_our_classic_solver():
    """north_star: configurations without a stored reference run are compared with classic_solver — here OUR Drake-free one."""
    from gcs_admm_b200.classic import solve_classic
    from gcs_admm_b200.generator import grid_problem, packed_to_dicts
    from gcs_admm_b200.lib import Solver
    off, A, b, s_pt, t_pt = grid_problem(6)
    As, bs = packed_to_dicts(off, A, b)
    ref = solve_classic(As, bs, 2, round_solution=False)
    assert ref["status"] == "optimal"
    s = Solver(pack_graph(As, bs), max_it=40010, eps_abs=0.0, eps_rel=0.0).enable_perf(inner_iters=1)
    s.step(40000)
    assert abs(_cost(s) - ref["cost"]) <= 1e-3 * ref["cost"]
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
