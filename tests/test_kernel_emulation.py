"""Kernel K1 arithmetic, checked on the CPU: vertex_ipm.cuh / vertex_update.cuh compiled with
GCS_EMULATE (a warp = one host thread) against the C oracle's x-update on real trajectories.
The emulation library is test infrastructure only (never loaded by the product path)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from gcs_admm_b200.graph import pack_graph

CSRC = os.path.join(ROOT, "gcs-admm_b200", "csrc")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def emu():
    from conftest import build_emu
    lib = C.CDLL(build_emu())
    lib.gcsemu_vertex_update_all.restype = C.c_int
    lib.gcsemu_vertex_update_all.argtypes = [C.c_int, C.c_int, _ip, _dp, _dp, _ip, _ip, _bp, _bp, _dp, _dp, _dp, _dp,
                                             _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                             C.c_int, C.POINTER(C.c_long), C.c_void_p, C.c_double]
    lib.gcsemu_ws_stride.restype = C.c_int
    return lib


def emu_vertex_update(lib, g, mu, z, rho, tol=1e-9, max_iter=60, ws=None, theta=0.0):
    xc = np.zeros((g.H, 5))
    x_v, z_v, y_v = np.zeros((g.nV, 4)), np.zeros((g.nV, 4)), np.zeros(g.nV)
    iters = C.c_long()
    cent = np.ascontiguousarray(g.interior_points())
    fails = lib.gcsemu_vertex_update_all(g.nV, g.nE, g.poly_off, g.polyA.reshape(-1), g.polyb, g.he_off, g.he_edge,
                                         g.he_flags, g.vtype, cent.reshape(-1), xc.reshape(-1),
                                         np.ascontiguousarray(mu).reshape(-1), np.ascontiguousarray(z).reshape(-1),
                                         x_v.reshape(-1), z_v.reshape(-1), y_v, rho, 1.0, tol, max_iter,
                                         max(1, g.max_live_degree), max(1, g.max_rows), C.byref(iters),
                                         ws.ctypes.data_as(C.c_void_p) if ws is not None else None, theta)
    return xc, x_v, z_v, y_v, fails, iters.value


@pytest.mark.parametrize("name,its", [("benchmark1", 39), ("test2", 12), ("test3", 12), ("benchmark2", 30),
                                      ("test_autogen2", 6), ("benchmark4", 12), ("benchmark3", 6)])
def test_k1_matches_oracle_x_update(emu, name, its):
    from c_oracle import COracle
    As, bs, n, d, keys = load_golden(name)
    g = pack_graph(As, bs)
    o = COracle(g)
    worst = 0.0
    for it in range(its):
        xc0, mu, z = o.state()
        rho = o.info()["rho"]
        xc, x_v, z_v, y_v, fails, iters = emu_vertex_update(emu, g, mu, z, rho)
        assert fails == 0
        o.vertex_update()                       # oracle x-update from the same (z, mu, rho)
        xo, _, _ = o.state()
        _, zvo, yvo = o.solution()
        worst = max(worst, float(np.max(np.abs(xc - xo))))
        assert np.max(np.abs(xc - xo)) < 5e-5, (name, it)   # both solves stop at ~1e-9 gap
        assert np.max(np.abs(z_v - zvo)) < 1e-4 and np.max(np.abs(y_v - yvo)) < 1e-4   # z_v holds weakly determined second points
        o.step(1)
    print(name, "max |xc_kernel - xc_oracle| =", worst)
