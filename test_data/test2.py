"""test2: 2-D GCS shortest-path problem (2 regions).
Data exported from the reference problem set by tools/export_golden.py."""
import os
import sys

import numpy as np

sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from utils import convert_pt_to_polytope, visualize_results

s = np.array([0.0, 1.0])
t = np.array([1.9, 1.0])
A_s, b_s = convert_pt_to_polytope(s, eps=1e-6)
A_t, b_t = convert_pt_to_polytope(t, eps=1e-6)

regions = [
    (np.array([[1.0, 1.0], [-1.0, 0.0], [0.0, -1.0]]), np.array([1.0, 0.0, 0.0])),
    (np.array([[0.0, -1.0], [1.0, 0.0], [-1.0, 1.0]]), np.array([0.0, 1.9, -0.9])),
]

As = {"s": A_s, "t": A_t}
bs = {"s": b_s, "t": b_t}
for _k, (_A, _b) in enumerate(regions):
    As[_k] = _A
    bs[_k] = _b

n = regions[0][0].shape[1]

# rounding hints (unused by the solvers, kept for format compatibility)
N = 1
M = 1
