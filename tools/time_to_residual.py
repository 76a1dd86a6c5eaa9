"""BASELINE metric "time to residual 1e-4": perf mode from a cold start until max(pri, dual) < tol, with a trace.
usage: time_to_residual.py G tol max_iters K [adapt_window_iters]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import utils  # noqa
import numpy as np
from gcs_admm_b200.generator import grid_packed_graph
from gcs_admm_b200 import lib, perf

G, tol, max_iters, K = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
window = int(sys.argv[5]) if len(sys.argv) > 5 else max_iters
g = grid_packed_graph(G)
T = perf.perf_tables(g)
# rho adapts while it < frac * max_it (reference rule :703): the window is a parameter of the reference's algorithm
s = lib.Solver(g, max_it=max_iters + 8, frac=window / (max_iters + 8), abs_stop=1, abs_tol=tol, check_every=64).enable_perf(inner_iters=K, tables=T)
t0 = time.perf_counter()
done, chunk, trace = 0, max(64, max_iters // 40), []
while done < max_iters:
    st = s.run(min(chunk, max_iters - done))
    done = st["iterations"]
    trace.append(dict(it=done, s=round(time.perf_counter() - t0, 3), pri=st["pri_res"], dual=st["dual_res"], rho=st["rho"]))
    print(json.dumps(trace[-1]), flush=True)
    if st["converged"] or st["diverged"]:
        break
x_v, z_v, y_v, z_e = s.solution()
cost = float(np.sum(np.linalg.norm(z_v[:, :2] - z_v[:, 2:], axis=1)) + 1e-4 * np.sum(z_e[:, 4]))
print(json.dumps(dict(workload=f"grid{G}x{G}", vertices=g.nV, edges=g.nE, mode=f"perf K={K}", tol=tol, reached=bool(st["converged"]), iterations=done,
                      seconds=time.perf_counter() - t0, pri=st["pri_res"], dual=st["dual_res"], rho=st["rho"], cost=cost,
                      straight_line=float(np.sqrt(2.0) * (G - 1)), rho_adaptation_window=window)))
s.close()
