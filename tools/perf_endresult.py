"""End-result parity of the inexact (`perf`) x-update: reference stop rule (eps_abs 1e-4 / eps_rel 1e-3), then the
reference's rounding; compare vertex path and final cost with the reference's stored v3 run (tests/golden)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import utils  # noqa
import numpy as np
from conftest import load_golden
from gcs_admm_b200.solver import solve

for name in ("benchmark1", "benchmark2", "benchmark3", "benchmark4"):
    As, bs, n, d, keys = load_golden(name)
    gold_on = {k for k, y in zip(keys, d["v3_y_v_rounded"]) if y > 0.5}
    gold_len = sum(np.linalg.norm(x[:2] - x[2:]) for x, y in zip(d["v3_x_v_rounded"], d["v3_y_v_rounded"]) if y > 0.5)
    for K in (1, 2, 3):
        for max_it in (1000, 20000):
            res = solve(As, bs, n, seed=0, mode="perf", inner_iters=K, max_it=max_it)
            print(json.dumps(dict(problem=name, K=K, max_it=max_it, iterations=res["iterations"], converged=bool(res["converged"]),
                                  relax_cost=res["cost"], v3_cost=float(d["v3_cost"]), classic_cost=float(d["classic_cost"]),
                                  same_path=set(res["path"]) == gold_on, final_cost=res["final_cost"], gold_final=float(gold_len),
                                  final_rel=abs(res["final_cost"] - gold_len) / gold_len)), flush=True)
