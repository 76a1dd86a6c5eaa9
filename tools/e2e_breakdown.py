"""Wall-clock breakdown of the host-buffer path (create / table upload / run / download / destroy)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import utils  # noqa
from gcs_admm_b200.generator import grid_packed_graph
from gcs_admm_b200 import lib, perf
G = int(sys.argv[1]) if len(sys.argv) > 1 else 316
g = grid_packed_graph(G)
T = perf.perf_tables(g)
lib.load()
for rep in range(3):
    t = [time.perf_counter()]
    s = lib.Solver(g, max_it=1000, check_every=64, eps_abs=0.0, eps_rel=0.0); t.append(time.perf_counter())
    s.enable_perf(inner_iters=1, tables=T); t.append(time.perf_counter())
    s.run(346); t.append(time.perf_counter())
    s.solution(); t.append(time.perf_counter())
    s.history(); t.append(time.perf_counter())
    s.close(); t.append(time.perf_counter())
    print(json.dumps(dict(zip(["create", "enable_perf", "run346", "solution", "history", "close"], [round(b - a, 4) for a, b in zip(t[:-1], t[1:])]))))
