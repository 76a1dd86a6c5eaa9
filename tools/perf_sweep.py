"""Convergence of the inexact (`perf`) mode vs the exact (`parity`) mode: iterations and seconds to reach
max(pri, dual) < tol on a G x G grid, for several K (inner splitting iterations) and relaxation alpha."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import utils  # noqa
import numpy as np
from gcs_admm_b200.generator import grid_packed_graph
from gcs_admm_b200 import lib, perf

G = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
max_it = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
g = grid_packed_graph(G)
def cost(s):
    x_v, z_v, y_v, z_e = s.solution()
    return float(np.sum(np.linalg.norm(z_v[:, :2] - z_v[:, 2:], axis=1)) + 1e-4 * np.sum(z_e[:, 4]))
rows = []
def run(label, **kw):
    s = lib.Solver(g, max_it=max_it, abs_stop=1, abs_tol=tol, check_every=32)
    if kw:
        s.enable_perf(**kw)
    t0 = time.perf_counter(); st = s.run(max_it); dt = time.perf_counter() - t0
    rows.append(dict(mode=label, iterations=st["iterations"], reached=bool(st["converged"]), seconds=dt, pri=st["pri_res"], dual=st["dual_res"], rho=st["rho"], cost=cost(s)))
    print(json.dumps(rows[-1]), flush=True)
    s.close()
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["parity", "1", "2", "3", "5", "a1.0", "a1.8"]
if "parity" in modes:
    run("parity")
T = perf.perf_tables(g)
for K in (1, 2, 3, 5):
    if str(K) in modes:
        run(f"perf K={K} alpha=1.6", inner_iters=K, alpha=1.6, tables=T)
for a in (1.0, 1.8):
    if f"a{a}" in modes:
        run(f"perf K=3 alpha={a}", inner_iters=3, alpha=a, tables=T)
