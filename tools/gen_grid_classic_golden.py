"""Classic-relaxation optimum of the G x G grid problem (the scalable benchmark family) for a few sizes the host interior-point
comparator can still finish (gcs_admm_b200.classic: minutes per size), stored as a fixture for the GPU test that compares the ADMM
fixed point at those sizes (tests/test_gpu_perf.py::test_grid_fixed_point_equals_classic_optimum).
usage: gen_grid_classic_golden.py G [G ...]   -> tests/golden/grid_classic.json (merged with what is there)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import utils  # noqa
from gcs_admm_b200.classic import solve_classic
from gcs_admm_b200.generator import grid_problem, packed_to_dicts

path = os.path.join(ROOT, "tests", "golden", "grid_classic.json")
out = json.load(open(path)) if os.path.exists(path) else {}
for G in map(int, sys.argv[1:]):
    off, A, b, _, _ = grid_problem(G)
    As, bs = packed_to_dicts(off, A, b)
    t0 = time.time()
    rc = solve_classic(As, bs, 2, round_solution=False)
    out[str(G)] = {"vertices": len(As), "cost": rc["cost"], "status": rc["status"], "ip_iterations": rc["iterations"],
                   "residuals": rc["residuals"], "seconds": round(time.time() - t0, 1),
                   "generator": "grid_problem(G) defaults (overlap 0.1, chamfer U(0.25, 0.40), default_rng(0))"}
    print(G, out[str(G)], flush=True)
    json.dump(out, open(path, "w"), indent=1)
