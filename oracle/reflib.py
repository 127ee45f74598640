"""ctypes wrapper of oracle/_ref/libpdpref.so: the REFERENCE'S OWN solver classes (lqr::LQRSolver,
lqr::LQRParallelSolver), compiled unmodified from /root/reference/include against oracle/eigen_shim (see
oracle/ref_driver.cpp, oracle/Makefile target `ref`).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libpdpref.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_LIB = None


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.c_int, C.c_int, _ip, C.c_int, C.c_int, C.c_int, C.c_int] + [_dp] * 7
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_update_problem_data.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_double]
        L.ref_backward.argtypes = [C.c_void_p, _dp, C.c_int]
        L.ref_forward.argtypes = [C.c_void_p, _dp, _dp]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(_dp)


class ReferenceSolver:
    """One problem (batch index b of a Problem) through the reference's LQRSolver / LQRParallelSolver."""

    def __init__(self, prob, b=0, parallel=False, num_segments=1, load_balancing=True, cholesky=True):
        self.p, self.b = prob, b
        ncs = None if prob.ncs is None else np.ascontiguousarray(prob.ncs, dtype=np.int32)
        self._keep = [np.ascontiguousarray(a[b], dtype=np.float64) for a in (prob.E, prob.c, prob.H, prob.h, prob.HN, prob.hN)]
        D = None if prob.D is None else np.ascontiguousarray(prob.D[b], dtype=np.float64)
        self.h = lib().ref_create(prob.nx, prob.nu, prob.N, None if ncs is None else ncs.ctypes.data_as(_ip),
                                  int(parallel), num_segments, int(load_balancing), int(cholesky),
                                  *[a.ctypes.data_as(_dp) for a in self._keep], _p(D))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_destroy(self.h)
            self.h = None

    def update_problem_data(self, ws, ys=None, zs=None, inv_rho=None, sigma=1e-6):
        lib().ref_update_problem_data(self.h, _p(ws), _p(ys), _p(zs), _p(inv_rho), sigma)

    def backward(self, rho=None, factorize=True):
        lib().ref_backward(self.h, _p(rho), int(factorize))

    def forward(self, x0):
        out = np.zeros(self.p.ws_len)
        lib().ref_forward(self.h, _p(x0), out.ctypes.data_as(_dp))
        return out

    def solve(self, ws_in=None, sigma=1e-6, ys=None, zs=None, rho=None, inv_rho=None):
        ws_in = np.zeros(self.p.ws_len) if ws_in is None else ws_in
        self.update_problem_data(ws_in, ys, zs, inv_rho, sigma)
        self.backward(rho)
        return self.forward(self.p.x0[self.b])
