"""BASELINE config 4: a batch of independent start/goal queries on the benchmark4 region set.

    python tools/batch_queries.py --queries 4096 [--gpus N via torchrun: each rank takes queries rank::world]

Every query is the 40 benchmark4 regions plus its own (s, t) drawn by rejection sampling inside two distinct
random regions (np.random.default_rng(1)); queries are packed block-diagonally (gcs_admm_b200.graph.pack_batch)
and solved in one handle per GPU with per-problem residuals / rho / stop — no communication.
Prints one JSON line: problems/s, aggregate ADMM iterations/s, iteration statistics.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import utils  # noqa: E402,F401
from gcs_admm_b200.graph import pack_batch, pack_graph  # noqa: E402
from gcs_admm_b200.queries import make_queries  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--max-it", type=int, default=1000)
    ap.add_argument("--mode", default="parity", choices=["parity", "perf"])
    ap.add_argument("--inner", type=int, default=1, help="K of the perf mode")
    ap.add_argument("--eps-rel", type=float, default=1e-3, help="reference stop rule (1e-3); perf mode wants a tighter one")
    ap.add_argument("--eps-abs", type=float, default=1e-4)
    ap.add_argument("--all-pairs", action="store_true", help="draw (s, t) from any two regions (46 %% of the pairs then have no path); default: same component")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from gcs_admm_b200 import lib
    t0 = time.perf_counter()
    qs = make_queries(args.queries, feasible_only=not args.all_pairs)[rank::world]
    graphs = [pack_graph(A, b) for A, b in qs]
    big = pack_batch(graphs)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    s = lib.Solver(big, device=local, max_it=args.max_it, check_every=16, eps_rel=args.eps_rel, eps_abs=args.eps_abs)
    if args.mode == "perf":
        s.enable_perf(inner_iters=args.inner)
    st = s.run(args.max_it)
    x_v, z_v, y_v, z_e = s.solution()
    dt = time.perf_counter() - t0
    its = np.array([s.problem_status(p)["iterations"] for p in range(len(graphs))])
    conv = np.array([s.problem_status(p)["converged"] for p in range(len(graphs))])
    line = {"mode": args.mode + (f" K={args.inner}" if args.mode == "perf" else ""), "eps_rel": args.eps_rel, "eps_abs": args.eps_abs, "max_it": args.max_it,
            "rank": rank, "world": world, "queries": len(graphs), "vertices": int(big.nV), "edges": int(big.nE),
            "host_build_s": t_build, "solve_s": dt, "problems_per_s": len(graphs) / dt,
            "aggregate_problem_iterations_per_s": float(its.sum()) / dt, "converged": int(conv.sum()),
            "iterations_min_median_max": [int(its.min()), int(np.median(its)), int(its.max())],
            "inner_fail": st["inner_fail"], "inner_iters": st["inner_iters"]}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
