"""compute-sanitizer target: a few iterations of both K1 kernels + the fused edge kernel on benchmark4 (eager launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import utils  # noqa
from conftest import load_golden
from gcs_admm_b200.graph import pack_graph
from gcs_admm_b200.lib import Solver
mode = sys.argv[1] if len(sys.argv) > 1 else "both"
g = pack_graph(*load_golden("benchmark4")[:2])
if mode in ("parity", "both"):
    s = Solver(g, use_graph=0)
    s.step(3)
    print("parity", s.status()["pri_res"])
    s.close()
if mode in ("perf", "both"):
    s = Solver(g, use_graph=0).enable_perf(inner_iters=2)
    s.step(5)
    print("perf", s.status()["pri_res"])
    s.close()
