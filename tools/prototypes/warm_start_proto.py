"""Prototype (CPU emulation of the perf kernel): iterations to max(pri, dual) < tol from a cold start vs a start whose DUALS are
initialised from a cost-to-go estimate (distance field), local frames.
usage: warm_start_proto.py G [rho] [mode: cold|euclid|dijkstra] [max_iters]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import utils  # noqa
from gcs_admm_b200.generator import grid_packed_graph
from gcs_admm_b200 import warmstart
from test_perf_mode import EmuPerfADMM, load_emu

G = int(sys.argv[1])
rho = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
mode = sys.argv[3] if len(sys.argv) > 3 else "cold"
max_iters = int(sys.argv[4]) if len(sys.argv) > 4 else 30000
tol = 1e-4
g = grid_packed_graph(G)
a = EmuPerfADMM(load_emu(), g, 1, frames="local")
a.rho = rho
if mode != "cold":
    a.mu[:] = warmstart.dual_start(g, a.delta, rho, field=mode)
t0 = time.time()
for it in range(1, max_iters + 1):
    a.step()
    if it % 500 == 0 or it in (1, 2, 5, 10, 20, 50, 100, 200) or max(a.pri[-1], a.dual[-1]) < tol:
        print(it, "pri %.3e dual %.3e cost %.5f  (%.1fs)" % (a.pri[-1], a.dual[-1], a.cost(), time.time() - t0), flush=True)
    if max(a.pri[-1], a.dual[-1]) < tol:
        break
print("G", G, "rho", rho, mode, "iterations", it, "cost", a.cost(), "straight", np.sqrt(2) * (G - 1))
