"""Root-level ``utils`` module: the names the reference's problem files and
solver scripts import (reference ``utils.py``), implemented without pydrake.

``test_data/*.py`` problem files do ``from utils import convert_pt_to_polytope,
visualize_results`` after appending this directory to ``sys.path`` (reference
``test_data/benchmark1.py:11-13``); importing this module first makes them load
unmodified.
"""
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gcs_admm_b200  # noqa: E402,F401  (registers the package)
from gcs_admm_b200.graph import convert_pt_to_polytope, delta, build_graph  # noqa: E402,F401


def visualize_results(As, bs, x_v, y_v, x_v_rounded=None, y_v_rounded=None, legend=False, save_to_file=None):
    """Plot regions and the per-vertex segments (reference ``utils.py:101-194``).
    Plotting is outside the accelerated path; without matplotlib this is a no-op."""
    try:
        from gcs_admm_b200.plotting import draw
    except Exception as exc:  # matplotlib missing
        print(f"visualize_results: plotting unavailable ({exc})")
        return
    draw(As, bs, x_v, y_v, x_v_rounded, y_v_rounded, legend, save_to_file)


def save_data(data_file, As, bs, solve_time, cost, x_v_sol, y_v_sol, x_v_rounded, y_v_rounded,
              ADMM=True, iterations=None, rho_seq=None, pri_res_seq=None, dual_res_seq=None):
    """Pickle one run with the reference's schema (reference ``utils.py:197-233``)."""
    record = dict(As=As, bs=bs, solve_time=solve_time, cost=cost, x_v_sol=x_v_sol, y_v_sol=y_v_sol,
                  x_v_rounded=x_v_rounded, y_v_rounded=y_v_rounded, ADMM=ADMM)
    if ADMM:
        record.update(iterations=iterations, rho_seq=rho_seq, pri_res_seq=pri_res_seq,
                      dual_res_seq=dual_res_seq)
    folder = os.path.dirname(data_file)
    if folder:
        os.makedirs(folder, exist_ok=True)
    with open(data_file, 'wb') as fh:
        pickle.dump(record, fh)
