"""GPU diagnostic: glue the CUDA iterates to the C oracle and report where they part."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import utils
from conftest import load_golden
from gcs_admm_b200.graph import pack_graph
from gcs_admm_b200.lib import Solver
from c_oracle import COracle

name = sys.argv[1]; its = int(sys.argv[2])
As, bs, n, d, keys = load_golden(name)
g = pack_graph(As, bs)
s, o = Solver(g), COracle(g)
prev_fail = 0
for it in range(its):
    s.step(1); o.step(1)
    xc, mu, z, rho, k = s.state()
    xo, muo, zo = o.state()
    st = s.status()
    dx = np.abs(xc - xo).max(axis=1)
    if st["inner_fail"] != prev_fail or dx.max() > 1e-4:
        h = int(np.argmax(dx))
        print(f"it {k}: inner_fail {st['inner_fail']} (+{st['inner_fail']-prev_fail}) max|dxc| {dx.max():.3e} at half-edge {h} owner {g.he_owner[h]} type {g.vtype[g.he_owner[h]]}")
        prev_fail = st["inner_fail"]
    s.set_state(xo, muo, zo, rho=o.info()["rho"], it=k)
print("glued run done; inner iters/solve", s.status()["inner_iters"] / (its * g.nV), "oracle", o.info()["inner_iters"] / (its * g.nV))
s.close()
s = Solver(g)
t = time.time(); st = s.run(); print("free run", st, "wall", time.time() - t)
