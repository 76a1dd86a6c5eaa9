"""ctypes binding of the C oracle ``oracle/gcs_admm_oracle.c`` (test infrastructure).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libgcsoracle.so")
    src = os.path.join(_HERE, "gcs_admm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libgcsoracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.gcso_create.restype = C.c_void_p
        L.gcso_create.argtypes = [C.c_int, C.c_int, _ip, _dp, _dp, _ip, _ip, _ip, _ip, _ip, C.c_int, C.c_int, _dp]
        L.gcso_destroy.argtypes = [C.c_void_p]
        L.gcso_set_params.argtypes = [C.c_void_p] + [C.c_double] * 7 + [C.c_int, C.c_double, C.c_int]
        L.gcso_step.restype = C.c_int
        L.gcso_step.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.gcso_get_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                    C.POINTER(C.c_long), C.POINTER(C.c_int)]
        L.gcso_get_history.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.gcso_get_state.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.gcso_set_state.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_int]
        L.gcso_get_solution.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.gcso_vertex_update_all.restype = C.c_int
        L.gcso_vertex_update_all.argtypes = [C.c_void_p]
        L.gcso_cost.restype = C.c_double
        L.gcso_cost.argtypes = [C.c_void_p]
        L.gcso_num_threads.restype = C.c_int
        L.gcso_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def use_all_cores():
    """OpenMP threads = host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1); returns the count"""
    n = os.cpu_count() or 1
    lib().gcso_set_num_threads(n)
    return lib().gcso_num_threads()


class COracle:
    """ADMM of reference admm_solver_v3.py on a ``PackedGraph`` (CPU, fp64)."""

    def __init__(self, g, inner_tol=1e-9, inner_max_iter=60, max_it=1000, **params):
        self.g = g
        L = lib()
        cent = np.ascontiguousarray(g.interior_points())
        self._h = L.gcso_create(g.nV, g.nE, g.poly_off, g.polyA.reshape(-1), g.polyb, g.he_off, g.he_edge,
                                g.he_out, g.edge_he_tail, g.edge_he_head, g.src, g.dst, cent.reshape(-1))
        p = dict(rho0=1.0, tau_incr=2.0, tau_decr=2.0, nu=10.0, frac=0.1, eps_abs=1e-4, eps_rel=1e-3)
        p.update(params)
        L.gcso_set_params(self._h, p["rho0"], p["tau_incr"], p["tau_decr"], p["nu"], p["frac"], p["eps_abs"],
                          p["eps_rel"], max_it, inner_tol, inner_max_iter)
        self.max_it = max_it

    def __del__(self):
        if getattr(self, "_h", None):
            lib().gcso_destroy(self._h)
            self._h = None

    def step(self, k=1, check_stop=False):
        return lib().gcso_step(self._h, k, int(check_stop))

    def run(self, max_it=None):
        return self.step(max_it or self.max_it, True)

    def info(self):
        it, opt, rho, ii, nf = C.c_int(), C.c_int(), C.c_double(), C.c_long(), C.c_int()
        lib().gcso_get_info(self._h, C.byref(it), C.byref(opt), C.byref(rho), C.byref(ii), C.byref(nf))
        return dict(it=it.value, opt=bool(opt.value), rho=rho.value, inner_iters=ii.value, inner_fail=nf.value)

    def history(self):
        n = self.info()["it"] + 1
        rho, pri, dual = np.zeros(n), np.zeros(n), np.zeros(n)
        lib().gcso_get_history(self._h, rho, pri, dual)
        return rho, pri, dual

    def state(self):
        H, nE = self.g.H, self.g.nE
        xc, mu, z = np.zeros((H, 5)), np.zeros((H, 5)), np.zeros((nE, 5))
        lib().gcso_get_state(self._h, xc.reshape(-1), mu.reshape(-1), z.reshape(-1))
        return xc, mu, z

    def set_state(self, xc, mu, z, rho, it):
        lib().gcso_set_state(self._h, np.ascontiguousarray(xc, np.float64).reshape(-1),
                             np.ascontiguousarray(mu, np.float64).reshape(-1),
                             np.ascontiguousarray(z, np.float64).reshape(-1), float(rho), int(it))

    def vertex_update(self):
        return lib().gcso_vertex_update_all(self._h)

    def solution(self):
        nV = self.g.nV
        x, z, y = np.zeros((nV, 4)), np.zeros((nV, 4)), np.zeros(nV)
        lib().gcso_get_solution(self._h, x.reshape(-1), z.reshape(-1), y)
        return x, z, y

    def cost(self):
        return lib().gcso_cost(self._h)
