"""Export the reference's problem data and stored runs into repo fixtures.

Run ONCE in the build container (needs /root/reference; the GPU box has no copy):

    python tools/export_golden.py

Reads  /root/reference/test_data/*.py          (problem files, imported unmodified)
       /root/reference/benchmark_data/*.pkl    (the author's stored runs = golden vectors)
Writes tests/golden/<name>.npz                 (problem + golden trajectory, numpy only)
       test_data/<name>.py                     (same problems re-emitted by our own writer)
"""
import importlib.util
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import utils  # noqa: E402,F401
from gcs_admm_b200.problem_io import write_test_file  # noqa: E402

REF = "/root/reference"
NAMES = ["test1", "test2", "test3", "test_autogen1", "test_autogen2",
         "benchmark1", "benchmark2", "benchmark3", "benchmark4"]


def load_ref_problem(name):
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", f"{REF}/test_data/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    os.makedirs(f"{ROOT}/tests/golden", exist_ok=True)
    os.makedirs(f"{ROOT}/test_data", exist_ok=True)
    for name in NAMES:
        mod = load_ref_problem(name)
        keys = list(mod.As.keys())
        out = dict(keys=np.array([str(k) for k in keys]), n=np.int64(mod.n),
                   s=np.asarray(mod.s, float), t=np.asarray(mod.t, float))
        for i, k in enumerate(keys):
            out[f"A_{i}"] = np.asarray(mod.As[k], float)
            out[f"b_{i}"] = np.asarray(mod.bs[k], float)
        for solver in ("admm_solver_v3", "classic_solver"):
            pkl = f"{REF}/benchmark_data/{solver}_{name}.pkl"
            if not os.path.exists(pkl):
                continue
            d = pickle.load(open(pkl, "rb"))
            tag = "v3" if solver.startswith("admm") else "classic"
            out[f"{tag}_cost"] = np.float64(d["cost"])
            out[f"{tag}_solve_time"] = np.float64(d["solve_time"])
            out[f"{tag}_y_v"] = np.array([float(d["y_v_sol"][k]) for k in keys])
            out[f"{tag}_x_v"] = np.array([np.asarray(d["x_v_sol"][k], float) for k in keys])
            out[f"{tag}_y_v_rounded"] = np.array([float(d["y_v_rounded"][k]) for k in keys])
            out[f"{tag}_x_v_rounded"] = np.array([np.asarray(d["x_v_rounded"][k], float) for k in keys])
            if d.get("ADMM"):
                out["v3_iterations"] = np.int64(d["iterations"])
                out["v3_rho_seq"] = np.asarray(d["rho_seq"], float)
                out["v3_pri_res_seq"] = np.asarray(d["pri_res_seq"], float)
                out["v3_dual_res_seq"] = np.asarray(d["dual_res_seq"], float)
        np.savez_compressed(f"{ROOT}/tests/golden/{name}.npz", **out)
        ints = {k: v for k, v in mod.As.items() if not isinstance(k, str)}
        write_test_file(f"{ROOT}/test_data/{name}.py", mod.As, mod.bs, s=mod.s, t=mod.t,
                        N=getattr(mod, "N", None), M=getattr(mod, "M", None),
                        header=f"{name}: 2-D GCS shortest-path problem ({len(ints)} regions).\n"
                               "Data exported from the reference problem set by tools/export_golden.py.\n")
        print(name, len(keys), "vertices ->", f"tests/golden/{name}.npz", f"test_data/{name}.py")


if __name__ == "__main__":
    main()
