"""ctypes wrapper of the CPU oracle (oracle/liboracle.so, built from pdp_oracle.cpp by oracle/Makefile).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs -- never by the product package.  Mirrors the reference's 4-call protocol
(lqr_solver_parallel.hpp:33-49): update_problem_data -> backward | backward_without_factorization -> forward.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
LU, CHOLESKY = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "pdp_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, C.c_int, C.c_int, _ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_model.argtypes = [C.c_void_p] + [_dp] * 7
        L.oracle_update_problem_data.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_double]
        L.oracle_backward.argtypes = [C.c_void_p, _dp]
        L.oracle_backward_without_factorization.argtypes = [C.c_void_p, _dp]
        L.oracle_forward.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_status.argtypes = [C.c_void_p]
        L.oracle_num_segments.argtypes = [C.c_void_p]
        L.oracle_get_partition.argtypes = [C.c_void_p, _ip, _ip]
        L.oracle_get_gains.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.oracle_get_value.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_get_costates.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_get_summary.argtypes = [C.c_void_p, C.c_int] + [_dp] * 5
        L.oracle_get_interface.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_batch_pool_create.restype = C.c_void_p
        L.oracle_batch_pool_create.argtypes = [C.c_int, C.c_int, C.c_int, _ip, C.c_int]
        L.oracle_batch_pool_destroy.argtypes = [C.c_void_p]
        L.oracle_batch_solve.argtypes = [C.c_void_p] + [_dp] * 12 + [C.c_double, _dp, _dp, C.c_int, C.c_int]
        L.oracle_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous float64"
    return a.ctypes.data_as(_dp)


def _ncs_ptr(ncs):
    if ncs is None:
        return None, None
    a = np.ascontiguousarray(ncs, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


class OracleSolver:
    """One problem (batch index `b` of a `Problem`).  parallel=False -> LQRSolver semantics
    (lqr_solver.hpp), parallel=True -> LQRParallelSolver semantics (lqr_solver_parallel.hpp)."""

    def __init__(self, prob, b: int = 0, parallel: bool = False, num_segments: int = 1, load_balancing: bool = True,
                 condensed: int = CHOLESKY, nthreads: int = 1):
        self.p = prob
        self.b = b
        self._ncs, ncs_ptr = _ncs_ptr(prob.ncs)
        self.h = lib().oracle_create(prob.nx, prob.nu, prob.N, ncs_ptr, int(parallel), num_segments,
                                     int(load_balancing), condensed, nthreads)
        if not self.h:
            raise RuntimeError("oracle_create failed (bad dimensions)")
        self.parallel = parallel
        # keep the arrays alive: the oracle (like the reference) stores pointers, not copies
        self._keep = [np.ascontiguousarray(a[b]) for a in (prob.E, prob.c, prob.H, prob.h, prob.HN, prob.hN)]
        self._D = None if prob.D is None else np.ascontiguousarray(prob.D[b])
        lib().oracle_set_model(self.h, *[_p(a) for a in self._keep], _p(self._D))
        self.S = lib().oracle_num_segments(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_destroy(self.h)
            self.h = None

    def update_problem_data(self, ws, ys=None, zs=None, inv_rho=None, sigma: float = 1e-6):
        self._upd = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (ws, ys, zs, inv_rho)]
        lib().oracle_update_problem_data(self.h, *[_p(a) for a in self._upd], sigma)

    def backward(self, rho=None):
        r = None if rho is None else np.ascontiguousarray(rho, dtype=np.float64)
        return lib().oracle_backward(self.h, _p(r))

    def backward_without_factorization(self, rho=None):
        r = None if rho is None else np.ascontiguousarray(rho, dtype=np.float64)
        return lib().oracle_backward_without_factorization(self.h, _p(r))

    def forward(self, x0, ws):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        lib().oracle_forward(self.h, _p(x0), _p(ws))
        return ws

    def solve(self, ws_in=None, sigma: float = 1e-6, ys=None, zs=None, rho=None, inv_rho=None, x0=None):
        p = self.p
        ws_in = np.zeros(p.ws_len) if ws_in is None else ws_in
        self.update_problem_data(ws_in, ys, zs, inv_rho, sigma)
        self.backward(rho)
        out = np.array(ws_in, dtype=np.float64, copy=True)
        return self.forward(p.x0[self.b] if x0 is None else x0, out)

    def status(self):
        return lib().oracle_status(self.h)

    def partition(self):
        st = np.zeros(self.S, np.int32)
        ln = np.zeros(self.S, np.int32)
        lib().oracle_get_partition(self.h, st.ctypes.data_as(_ip), ln.ctypes.data_as(_ip))
        return st, ln

    def gains(self):
        p = self.p
        K = np.zeros((p.N, p.nu * p.nx))
        d = np.zeros((p.N, p.nu))
        Gt = np.zeros((p.N, p.nu * p.nx))
        lib().oracle_get_gains(self.h, _p(K), _p(d), _p(Gt))
        return K, d, Gt

    def value(self):
        p = self.p
        P = np.zeros((p.N + 1, p.nx * p.nx))
        pv = np.zeros((p.N + 1, p.nx))
        lib().oracle_get_value(self.h, _p(P), _p(pv))
        return P, pv

    def costates(self, ws):
        """lambda_1 .. lambda_N [N, nx] for the trajectory `ws` of the last solve (the reference's commented formula,
        lqr_kernel.hpp:205-211 / lqr_kernel_parallel.hpp:207-216)."""
        lam = np.zeros((self.p.N, self.p.nx))
        lib().oracle_get_costates(self.h, _p(np.ascontiguousarray(ws, dtype=np.float64)), _p(lam))
        return lam

    def summary(self, seg):
        n = self.p.nx
        P, F, Cm = np.zeros(n * n), np.zeros(n * n), np.zeros(n * n)
        pv, f = np.zeros(n), np.zeros(n)
        lib().oracle_get_summary(self.h, seg, _p(P), _p(pv), _p(F), _p(f), _p(Cm))
        return P, pv, F, f, Cm

    def interface(self):
        n = self.p.nx
        xh, uh = np.zeros((self.S, n)), np.zeros((self.S, n))
        lib().oracle_get_interface(self.h, _p(xh), _p(uh))
        return xh, uh


class OracleBatch:
    """Batched CPU baseline: OpenMP over problems, each a sequential-Riccati solve."""

    def __init__(self, prob):
        self.p = prob
        self._ncs, ncs_ptr = _ncs_ptr(prob.ncs)
        self.pool = lib().oracle_batch_pool_create(prob.nx, prob.nu, prob.N, ncs_ptr, prob.batch)

    def __del__(self):
        if getattr(self, "pool", None):
            lib().oracle_batch_pool_destroy(self.pool)
            self.pool = None

    def solve(self, ws_in=None, sigma=1e-6, ys=None, zs=None, rho=None, inv_rho=None, factorize=True, nthreads=0,
              ws_out=None):
        p = self.p
        ws_in = p.zeros_ws() if ws_in is None else ws_in
        ws_out = np.empty_like(ws_in) if ws_out is None else ws_out
        bad = lib().oracle_batch_solve(self.pool, _p(p.E), _p(p.c), _p(p.H), _p(p.h), _p(p.HN), _p(p.hN), _p(p.D),
                                       _p(ws_in), _p(ys), _p(zs), _p(rho), _p(inv_rho), sigma, _p(p.x0), _p(ws_out),
                                       int(factorize), nthreads)
        return ws_out, bad


def max_threads() -> int:
    return lib().oracle_max_threads()
