"""gcs-admm_b200 — B200-native full-vertex-split ADMM for shortest paths in graphs of convex sets.

The directory name carries a hyphen (it is the project name); import it as
``gcs_admm_b200`` through the loader module of that name at the repo root.

Layout:
  csrc/        CUDA kernels (sm_100a) + the C-ABI (``include/gcsadmm.h``)
  graph.py     Drake-free graph construction and the half-edge CSR layout
  lib.py       ctypes binding of ``libgcsadmm.so`` (fails loudly if absent)
  solver.py    ``solve(As, bs, n)``: host mirror of reference ``admm_solver_v3.py``
  rounding.py  randomized-DFS rounding + convex restriction (reference ``GCS_utils.py``)
  conic.py     small dense conic-QP interior-point solver (host side: rounding SOCPs)
  generator.py scalable synthetic 2-D problem generators (reference ``test_generator.py``)
  partition.py vertex partitioning + halo maps for multi-GPU
  dist.py      one-process-per-GPU driver (torch.distributed / NCCL halo exchange)
"""
__version__ = "0.1.0"
