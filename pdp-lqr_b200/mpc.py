"""Receding-horizon (MPC) driver on top of the conic ADMM solve.

The reference ships no MPC loop: `examples/lqr_example.cpp` solves one horizon and the ADMM hooks
(`lqr_solver_parallel.hpp:33-37`) are what an outer loop would call.  This module adds that outer loop for the CUDA
solver: every control period the conic LQ problem is re-solved from the measured state, warm-started with the previous
solution shifted by one stage, and the first control is applied.

The model handed to the solver stays resident on the device (uploaded once, as the reference builds its workspaces
once); only `x0` and the warm start change between periods, so a period costs one `pdplqr_admm_solve`.
"""
from __future__ import annotations

import numpy as np


def shift_warm_start(ws, zs, ys, nx: int, nu: int, ncs):
    """Shift a solution by one stage (the usual MPC warm start).

    ws [batch, N*(nx+nu)+nx]: stage k <- stage k+1 for k < N-1; the last stage repeats its control (u_{N-1}) from the
        terminal state x_N; x_N is kept.
    zs, ys [batch, sum(ncs)]: the blocks of stage k <- stage k+1 where both stages have the same number of rows (the
        reference example has a shorter first block, `lqr_example.cpp:126-147`: nu rows at k = 0, nx + nu after); blocks
        whose sizes differ, the last running stage and the terminal stage keep their values.
    Returns new arrays; inputs are not modified.
    """
    ws = np.array(ws, dtype=np.float64, copy=True)
    zs = np.array(zs, dtype=np.float64, copy=True)
    ys = np.array(ys, dtype=np.float64, copy=True)
    s = nx + nu
    N = (ws.shape[1] - nx) // s
    if N >= 2:
        ws[:, : (N - 1) * s] = ws[:, s: N * s].copy()
    if N >= 1:   # last running stage: previous last control, state = previous terminal state
        ws[:, (N - 1) * s + nu: N * s] = ws[:, N * s: N * s + nx]
    ncs = np.asarray(ncs, dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(ncs)])
    for k in range(len(ncs) - 2):        # running stages 0 .. N-2 take the block of stage k+1 (never the terminal one)
        if ncs[k] == ncs[k + 1] and ncs[k] > 0:
            zs[:, off[k]: off[k + 1]] = zs[:, off[k + 1]: off[k + 2]].copy()
            ys[:, off[k]: off[k + 1]] = ys[:, off[k + 1]: off[k + 2]].copy()
    return ws, zs, ys


class RecedingHorizon:
    """Closed loop: plant x+ = A_0 x + B_0 u + c_0 (stage 0 of the model, batch-wise) driven by the first control of a
    conic LQ solve per period.

        rh = RecedingHorizon(solver, problem, rho=0.1)
        for t in range(T):
            u, info = rh.step()          # solves from rh.x, applies u_0, advances rh.x
    """

    def __init__(self, solver, problem, rho=0.1, sigma=1e-6, alpha=1.6, max_iter=200, eps_abs=1e-4, eps_rel=1e-4,
                 check_every=10, warm_start=True):
        self.sol, self.p = solver, problem
        self.nx, self.nu, self.N, self.batch = problem.nx, problem.nu, problem.N, problem.batch
        nct = problem.nc_total
        self.rho = np.full((self.batch, nct), float(rho)) if np.isscalar(rho) else np.ascontiguousarray(rho, dtype=np.float64)
        self.sigma, self.alpha = float(sigma), float(alpha)
        self.max_iter, self.eps_abs, self.eps_rel, self.check_every = int(max_iter), eps_abs, eps_rel, int(check_every)
        self.warm_start = bool(warm_start)
        self.x = np.array(problem.x0, dtype=np.float64, copy=True)
        self.ws = problem.zeros_ws()
        self.zs = np.zeros((self.batch, nct))
        self.ys = np.zeros((self.batch, nct))
        lb = np.where(np.isfinite(problem.e_lb), problem.e_lb, -1e20)
        ub = np.where(np.isfinite(problem.e_ub), problem.e_ub, 1e20)
        solver.admm_set_cones(problem.cones, lb, ub)
        s = self.nx + self.nu
        E0 = np.asarray(problem.E).reshape(self.batch, self.N, s, self.nx)[:, 0]     # column-major nx x s per stage
        self._E0 = np.transpose(E0, (0, 2, 1))                                         # [batch, nx, s]
        self._c0 = np.asarray(problem.c).reshape(self.batch, self.N, self.nx)[:, 0]
        self.history = []

    def plant(self, x, u, disturbance=None):
        w = np.concatenate([u, x], axis=1)
        xn = np.einsum("bij,bj->bi", self._E0, w) + self._c0
        return xn if disturbance is None else xn + disturbance

    def step(self, disturbance=None):
        if self.warm_start and self.history:
            self.ws, self.zs, self.ys = shift_warm_start(self.ws, self.zs, self.ys, self.nx, self.nu, self.p.ncs)
        else:
            self.ws[:] = 0.0
            self.zs[:] = 0.0
            self.ys[:] = 0.0
        iters, res = self.sol.admm_solve(self.x, self.ws, self.zs, self.ys, self.rho, sigma=self.sigma, alpha=self.alpha,
                                         max_iter=self.max_iter, eps_abs=self.eps_abs, eps_rel=self.eps_rel,
                                         check_every=self.check_every)
        u = self.ws[:, : self.nu].copy()
        info = {"iterations": iters, "r_prim": float(res[0]), "r_dual": float(res[1]), "x": self.x.copy()}
        self.history.append(info)
        self.x = self.plant(self.x, u, disturbance)
        return u, info
