"""BASELINE metric "time to residual 1e-4": perf mode from a cold start until max(pri, dual) < tol, with a trace.
usage: time_to_residual.py --grid G [--tol 1e-4] [--max-iters N] [--inner K] [--window W] [--adapt-every A] [--outer-alpha a] [--rho0 r] [--nu v] [--tau t] [--theta th] [--warm dijkstra]"""
import argparse, sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import utils  # noqa
import numpy as np
from gcs_admm_b200.generator import grid_packed_graph
from gcs_admm_b200 import lib, perf

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=100)
ap.add_argument("--tol", type=float, default=1e-4)
ap.add_argument("--max-iters", type=int, default=400000)
ap.add_argument("--inner", type=int, default=1)
ap.add_argument("--window", type=int, default=100, help="rho adapts while it < window (reference: 0.1 * MAX_IT = 100)")
ap.add_argument("--adapt-every", type=int, default=1)
ap.add_argument("--outer-alpha", type=float, default=1.0)
ap.add_argument("--rho0", type=float, default=1.0)
ap.add_argument("--nu", type=float, default=10.0)
ap.add_argument("--tau", type=float, default=2.0)
ap.add_argument("--frames", type=str, default="local", choices=["local", "global"])
ap.add_argument("--theta", type=float, default=1.0, help="penalty of the flow scalars = theta * rho")
ap.add_argument("--warm", type=str, default="none", choices=["none", "dijkstra", "euclid"], help="dual start from a cost-to-go field (gcs_admm_b200.warmstart)")
ap.add_argument("--stop-ref", action="store_true", help="stop on the residuals in global coordinates (the reference's definitions) instead of the local-frame ones")
ap.add_argument("--trace", type=int, default=20, help="trace points")
ap.add_argument("--budget", type=float, default=1e9, help="seconds")
a = ap.parse_args()
g = grid_packed_graph(a.grid)
T = perf.perf_tables(g, frames=a.frames, theta=a.theta)
# rho adapts while it < frac * max_it (reference rule :703): the window is a parameter of the reference's algorithm
s = lib.Solver(g, max_it=a.max_iters + 8, frac=a.window / (a.max_iters + 8), abs_stop=1, abs_tol=a.tol, check_every=256, rho0=a.rho0, nu=a.nu,
               tau_incr=a.tau, tau_decr=a.tau, outer_alpha=a.outer_alpha, adapt_every=a.adapt_every, stop_ref=int(a.stop_ref)).enable_perf(inner_iters=a.inner, tables=T)
t0 = time.perf_counter()
if a.warm != "none":
    from gcs_admm_b200 import warmstart
    mu0 = warmstart.dual_start(g, T["edge_delta"], a.rho0, field=a.warm)
    s.set_state(np.zeros((g.H, 5)), mu0, np.zeros((g.nE, 5)), a.rho0, 0)
    print(json.dumps(dict(warm=a.warm, host_seconds=round(time.perf_counter() - t0, 3))), flush=True)
done, chunk = 0, max(256, a.max_iters // a.trace)
while done < a.max_iters and time.perf_counter() - t0 < a.budget:
    st = s.run(min(chunk, a.max_iters - done))
    done = st["iterations"]
    _, zv_, _, ze_ = s.solution()
    cost_ = float(np.sum(np.linalg.norm(zv_[:, :2] - zv_[:, 2:], axis=1)) + 1e-4 * np.sum(ze_[:, 4]))
    print(json.dumps(dict(it=done, s=round(time.perf_counter() - t0, 3), pri=st["pri_res"], dual=st["dual_res"], inner=st["inner_res"], pri_ref=st["pri_res_ref"], dual_ref=st["dual_res_ref"], rho=st["rho"], cost=cost_)), flush=True)
    if st["converged"] or st["diverged"]:
        break
x_v, z_v, y_v, z_e = s.solution()
cost = float(np.sum(np.linalg.norm(z_v[:, :2] - z_v[:, 2:], axis=1)) + 1e-4 * np.sum(z_e[:, 4]))
print(json.dumps(dict(workload=f"grid{a.grid}x{a.grid}", vertices=g.nV, edges=g.nE, mode=f"perf K={a.inner}", tol=a.tol, reached=bool(st["converged"]), iterations=done,
                      seconds=time.perf_counter() - t0, pri=st["pri_res"], dual=st["dual_res"], inner=st["inner_res"], pri_ref=st["pri_res_ref"], dual_ref=st["dual_res_ref"], rho=st["rho"], cost=cost,
                      straight_line=float(np.sqrt(2.0) * (a.grid - 1)), params=vars(a))))
s.close()
