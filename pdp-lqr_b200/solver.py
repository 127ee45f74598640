"""Host-side mirror of the reference solver interface on top of the C ABI.

`LQRCudaSolver` keeps the reference's call protocol and argument meaning
(/root/reference include/clqr/lqr/lqr_solver_parallel.hpp:19-62):
    LQRParallelSolver(model, num_segments, load_balancing=True, solver_type=CHOLESKY)
    update_problem_data(ws, ys, zs, inv_rho_vecs, sigma) -> backward(rho_vecs) | backward_without_factorization(rho_vecs)
    -> forward(x0, ws)
with the flat array layout of include/pdplqr.h instead of std::vector<Eigen::VectorXd>.  The C++17 header with
the reference's exact C++ signatures is include/pdplqr/lqr_cuda_solver.hpp; this Python class exists so that the
parity tests and bench.py read like the reference's example (examples/lqr_example.cpp:211-217).
PyTorch is used only for device memory / streams in the *_device variants.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

LU, CHOLESKY = capi.CONDENSED_LU, capi.CONDENSED_CHOLESKY


# filled by LQRCudaSolver.close() when PDPLQR_DEBUG_GUARDS=1 (tests/conftest.py asserts on it after every test)
GUARD_STATS = {"enabled": os.environ.get("PDPLQR_DEBUG_GUARDS", "0") not in ("", "0"), "handles": 0, "corrupted_bytes": 0,
               "unguarded": 0}


class PdplqrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pdplqr error {code}: {msg}")
        self.code = code


def _hp(a):
    """host pointer of a contiguous float64 numpy array (or None)."""
    if a is None:
        return None
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
        raise TypeError("expected a C-contiguous float64 numpy array")
    return a.ctypes.data


def _dptr(t):
    """device pointer of a torch CUDA tensor (or None)."""
    if t is None:
        return None
    import torch
    assert isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
    return t.data_ptr()


def wave_size(nx: int, nu: int, device: int = 0) -> int:
    """(problem, segment) groups one GPU keeps resident in the throughput-mode stage sweep (pdplqr_wave_size)."""
    n = capi.load().pdplqr_wave_size(nx, nu, device)
    if n <= 0:
        raise PdplqrError(n, "pdplqr_wave_size failed (no CUDA device, or (nx, nu) too large)")
    return n


class LQRCudaSolver:
    def __init__(self, nx, nu, N, batch=1, num_segments=1, load_balancing=True, solver_type=CHOLESKY, ncs=None,
                 device=0):
        self._lib = capi.load()
        self.nx, self.nu, self.N, self.batch, self.s = nx, nu, N, batch, nx + nu
        self.ws_len = N * self.s + nx
        h = C.c_void_p()
        ncs_arr = None if ncs is None else np.ascontiguousarray(ncs, dtype=np.int32)
        ncs_ptr = None if ncs_arr is None else ncs_arr.ctypes.data_as(C.POINTER(C.c_int))
        rc = self._lib.pdplqr_create(C.byref(h), nx, nu, N, ncs_ptr, batch, num_segments, int(load_balancing),  # 2 = equal split
                                     solver_type, device)
        if rc != capi.OK:
            why = self._lib.pdplqr_last_error(None).decode()   # text of the failed create (no handle survives it)
            raise PdplqrError(rc, {capi.ERR_INVALID: "invalid dimensions / arguments",
                                   capi.ERR_UNSUPPORTED: f"(nx={nx}, nu={nu}) exceeds the largest instantiated kernel size",
                                   capi.ERR_CUDA: "CUDA failure or no usable CUDA device (there is no CPU fallback)"}.get(rc, "create failed")
                              + (f": {why}" if why else ""))
        self._h = h
        self.num_segments = self._lib.pdplqr_num_segments(h)
        self.nc_total = 0 if ncs is None else int(np.sum(ncs))

    @classmethod
    def from_problem(cls, prob, **kw):
        s = cls(prob.nx, prob.nu, prob.N, batch=prob.batch, ncs=prob.ncs, **kw)
        s.set_model(prob)
        return s

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            if GUARD_STATS["enabled"]:   # PDPLQR_DEBUG_GUARDS=1: out-of-bounds device writes of this handle's lifetime
                n = self.debug_check_guards()
                GUARD_STATS["handles"] += 1
                GUARD_STATS["corrupted_bytes"] += max(n, 0)
                GUARD_STATS["unguarded"] += n < 0
            self._lib.pdplqr_destroy(self._h)
            self._h = None

    def debug_check_guards(self) -> int:
        """Guard bytes around this handle's device allocations that were overwritten (pdplqr_debug_check_guards):
        0 = clean, -1 = created without PDPLQR_DEBUG_GUARDS=1."""
        n = C.c_longlong()
        self._check(self._lib.pdplqr_debug_check_guards(self._h, C.byref(n)))
        return int(n.value)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != capi.OK:
            raise PdplqrError(rc, self._lib.pdplqr_last_error(self._h).decode())

    # ------------------------------------------------------------------ model
    def set_model(self, prob):
        """Upload E, c, H, h, HN, hN (and D): the reference re-reads `const LQRModel&` on every call
        (lqr_solver_parallel.hpp:52); the device copy must be refreshed explicitly after mutating the model."""
        self._keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (prob.E, prob.c, prob.H, prob.h, prob.HN, prob.hN)]
        D = None if prob.D is None else np.ascontiguousarray(prob.D, dtype=np.float64)
        self._check(self._lib.pdplqr_set_model(self._h, *[_hp(a) for a in self._keep], _hp(D)))

    def set_model_device(self, E, c, H, h, HN, hN, D=None):
        self._check(self._lib.pdplqr_set_model_device(self._h, *[_dptr(t) for t in (E, c, H, h, HN, hN)], _dptr(D)))

    def set_option(self, option: int, value: int):
        self._check(self._lib.pdplqr_set_option(self._h, option, value))

    def set_stream(self, cuda_stream_ptr: int):
        self._check(self._lib.pdplqr_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    # ------------------------------------------------------------------ the reference's 4-call protocol (host arrays)
    def update_problem_data(self, ws, ys=None, zs=None, inv_rho_vecs=None, sigma=1e-6):
        self._check(self._lib.pdplqr_update_problem_data(self._h, _hp(ws), _hp(ys), _hp(zs), _hp(inv_rho_vecs), sigma))

    def backward(self, rho_vecs=None):
        self._check(self._lib.pdplqr_backward(self._h, _hp(rho_vecs)))

    def backward_without_factorization(self, rho_vecs=None):
        self._check(self._lib.pdplqr_backward_without_factorization(self._h, _hp(rho_vecs)))

    def forward(self, x0, ws):
        self._check(self._lib.pdplqr_forward(self._h, _hp(np.ascontiguousarray(x0, dtype=np.float64)), _hp(ws)))
        return ws

    def solve(self, ws_in, x0, ws_out, sigma=1e-6, ys=None, zs=None, rho=None, inv_rho=None):
        self._check(self._lib.pdplqr_solve(self._h, _hp(ws_in), _hp(ys), _hp(zs), _hp(rho), _hp(inv_rho), sigma,
                                           _hp(x0), _hp(ws_out)))
        return ws_out

    # ------------------------------------------------------------------ device-resident variants (torch tensors)
    def update_problem_data_device(self, ws, ys=None, zs=None, inv_rho_vecs=None, sigma=1e-6):
        self._check(self._lib.pdplqr_update_problem_data_device(self._h, _dptr(ws), _dptr(ys), _dptr(zs),
                                                                 _dptr(inv_rho_vecs), sigma))

    def backward_device(self, rho_vecs=None):
        self._check(self._lib.pdplqr_backward_device(self._h, _dptr(rho_vecs)))

    def backward_without_factorization_device(self, rho_vecs=None):
        self._check(self._lib.pdplqr_backward_without_factorization_device(self._h, _dptr(rho_vecs)))

    def forward_device(self, x0, ws_out):
        self._check(self._lib.pdplqr_forward_device(self._h, _dptr(x0), _dptr(ws_out)))

    def solve_device(self, ws_in, x0, ws_out, sigma=1e-6, ys=None, zs=None, rho=None, inv_rho=None):
        """update_problem_data + backward + forward as one CUDA graph launch (pdplqr_solve_device); asynchronous."""
        self._check(self._lib.pdplqr_solve_device(self._h, _dptr(ws_in), _dptr(ys), _dptr(zs), _dptr(rho), _dptr(inv_rho),
                                                   sigma, _dptr(x0), _dptr(ws_out)))

    # ------------------------------------------------------------------ conic ADMM outer iteration (addition, a11)
    def admm_set_cones(self, cones, e_lb, e_ub):
        """cones: list of (stage, first_row, dim, type) sorted by stage, tiling every stage's rows."""
        arr = np.ascontiguousarray(np.array(cones, dtype=np.int32).reshape(-1, 4))
        cols = [np.ascontiguousarray(arr[:, i]) for i in range(4)]
        ip = C.POINTER(C.c_int)
        self._cone_keep = cols
        self._check(self._lib.pdplqr_admm_set_cones(self._h, len(arr), *[c.ctypes.data_as(ip) for c in cols],
                                                     _hp(np.ascontiguousarray(e_lb, dtype=np.float64)),
                                                     _hp(np.ascontiguousarray(e_ub, dtype=np.float64))))

    def admm_configure(self, use_graph=True, adaptive_rho=False, rho_tau=5.0, max_rho_updates=10):
        self._check(self._lib.pdplqr_admm_configure(self._h, int(use_graph), int(adaptive_rho), rho_tau, max_rho_updates))

    def admm_stats(self):
        """(graph launches so far, rho rescales of the last solve)."""
        a, b = C.c_int(), C.c_int()
        self._check(self._lib.pdplqr_admm_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def admm_solve(self, x0, ws, zs, ys, rho, sigma=1e-6, alpha=1.6, max_iter=50, eps_abs=1e-4, eps_rel=1e-4,
                   check_every=10):
        it = C.c_int()
        res = np.zeros(2)
        self._check(self._lib.pdplqr_admm_solve(self._h, _hp(np.ascontiguousarray(x0, dtype=np.float64)), _hp(ws),
                                                _hp(zs), _hp(ys), _hp(rho), sigma, alpha, max_iter, eps_abs, eps_rel,
                                                check_every, C.byref(it), _hp(res)))
        return it.value, res

    def admm_solve_device(self, x0, w, z, y, rho, inv_rho, sigma=1e-6, alpha=1.6, max_iter=50, eps_abs=1e-4,
                          eps_rel=1e-4, check_every=10):
        it = C.c_int()
        res = np.zeros(2)
        self._check(self._lib.pdplqr_admm_solve_device(self._h, _dptr(x0), _dptr(w), _dptr(z), _dptr(y), _dptr(rho),
                                                       _dptr(inv_rho), sigma, alpha, max_iter, eps_abs, eps_rel,
                                                       check_every, C.byref(it), _hp(res)))
        return it.value, res

    # ------------------------------------------------------------------ horizon sharding (one handle per time slice)
    def summary_doubles(self) -> int:
        return int(self._lib.pdplqr_summary_doubles(self._h))

    def root_summary_device(self, out):
        """[batch, summary_doubles] device tensor <- summary (P | F | C | p | f) of this handle's whole slice."""
        self._check(self._lib.pdplqr_get_root_summary_device(self._h, _dptr(out)))

    def set_root_boundary_device(self, xhat, lam=None):
        self._check(self._lib.pdplqr_set_root_boundary_device(self._h, _dptr(xhat), _dptr(lam)))

    def synchronize(self):
        self._check(self._lib.pdplqr_synchronize(self._h))

    # ------------------------------------------------------------------ accessors (additions)
    def partition(self):
        S = self.num_segments
        st, ln = np.zeros(S, np.int32), np.zeros(S, np.int32)
        ip = C.POINTER(C.c_int)
        self._check(self._lib.pdplqr_get_partition(self._h, st.ctypes.data_as(ip), ln.ctypes.data_as(ip)))
        return st, ln

    def gains(self):
        K = np.zeros((self.batch, self.N, self.nu * self.nx))
        d = np.zeros((self.batch, self.N, self.nu))
        Gt = np.zeros((self.batch, self.N, self.nu * self.nx))
        self._check(self._lib.pdplqr_get_gains(self._h, _hp(K), _hp(d), _hp(Gt)))
        return K, d, Gt

    def interface(self):
        xh = np.zeros((self.batch, self.num_segments, self.nx))
        uh = np.zeros((self.batch, self.num_segments, self.nx))
        self._check(self._lib.pdplqr_get_interface(self._h, _hp(xh), _hp(uh)))
        return xh, uh

    def summaries(self):
        n, S, B = self.nx, self.num_segments, self.batch
        P, F, Cm = (np.zeros((B, S, n * n)) for _ in range(3))
        p, f = np.zeros((B, S, n)), np.zeros((B, S, n))
        self._check(self._lib.pdplqr_get_summaries(self._h, _hp(P), _hp(p), _hp(F), _hp(f), _hp(Cm)))
        return P, p, F, f, Cm

    def costates(self, ws):
        """lambda_1 .. lambda_N [batch, N, nx] of the last solve (`ws` = the trajectory forward returned); the step the
        reference leaves commented out (lqr_kernel.hpp:205-211).  Segment-path handles only."""
        lam = np.zeros((self.batch, self.N, self.nx))
        self._check(self._lib.pdplqr_get_costates(self._h, _hp(np.ascontiguousarray(ws, dtype=np.float64)), _hp(lam)))
        return lam

    def costates_device(self, ws, lam):
        self._check(self._lib.pdplqr_get_costates_device(self._h, _dptr(ws), _dptr(lam)))

    def last_status(self):
        st = np.zeros(self.batch, np.int32)
        bad = self._lib.pdplqr_last_status(self._h, st.ctypes.data_as(C.POINTER(C.c_int)))
        return bad, st

    def launch_count(self) -> int:
        return int(self._lib.pdplqr_launch_count(self._h))

    def record_doubles(self):
        a, b = C.c_int(), C.c_int()
        self._check(self._lib.pdplqr_record_doubles(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


class Coupler:
    """Interface system of G time slices (ranks): summaries [batch, G, summary_doubles] + x0 -> entry state and exit
    costate of every slice.  Device tensors only (include/pdplqr.h, pdplqr_coupler_*)."""

    def __init__(self, nx, nu, num_shards, batch=1, device=0):
        self._lib = capi.load()
        h = C.c_void_p()
        rc = self._lib.pdplqr_coupler_create(C.byref(h), nx, nu, num_shards, batch, device)
        if rc != capi.OK:
            raise PdplqrError(rc, "coupler_create failed")
        self._h = h
        self.nx, self.G, self.batch = nx, num_shards, batch

    def set_stream(self, cuda_stream_ptr: int):
        self._lib.pdplqr_set_stream(self._h, C.c_void_p(cuda_stream_ptr))

    def solve_device(self, summaries, x0, xhat, lam):
        rc = self._lib.pdplqr_coupler_solve_device(self._h, _dptr(summaries), _dptr(x0), _dptr(xhat), _dptr(lam))
        if rc != capi.OK:
            raise PdplqrError(rc, self._lib.pdplqr_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            if GUARD_STATS["enabled"]:
                n = C.c_longlong()
                self._lib.pdplqr_debug_check_guards(self._h, C.byref(n))
                GUARD_STATS["handles"] += 1
                GUARD_STATS["corrupted_bytes"] += max(int(n.value), 0)
                GUARD_STATS["unguarded"] += n.value < 0
            self._lib.pdplqr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
