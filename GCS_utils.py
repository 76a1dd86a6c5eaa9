"""Root-level ``GCS_utils`` module with the reference's public names
(reference ``GCS_utils.py``: ``solve_convex_restriction``, ``rounding``, ``compute_cost``),
implemented without pydrake in ``gcs-admm_b200/rounding.py``."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gcs_admm_b200  # noqa: E402,F401
from gcs_admm_b200.rounding import compute_cost, rounding, solve_convex_restriction  # noqa: E402,F401
from utils import *  # noqa: E402,F401,F403  (the reference module re-exports utils)
