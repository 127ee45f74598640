"""numpy restatement of the conic ADMM outer iteration of pdp-lqr_b200/csrc/admm_kernels.cuh, with the LQ solves done
by the CPU oracle (sequential LQRSolver semantics).  TEST INFRASTRUCTURE ONLY.  The outer iteration is NOT in the
reference (hooks only), so this file is what pins it: "parity unpinned" by the
reference, pinned against this restatement + solution properties (feasibility, KKT residuals): tests/test_oracle.py::
test_admm_restatement_converges_to_a_kkt_point_of_the_conic_qp / ..._second_order_cone_optimality check the converged
(w, y, z) of THIS file against the optimality conditions of the conic problem with the independent sparse KKT solve."""
import numpy as np

from . import oracle as O


def project(v, cones_k, lb, ub):
    v = v.copy()
    for (_, r0, d, typ) in cones_k:
        if typ == 0:
            v[r0:r0 + d] = np.minimum(np.maximum(v[r0:r0 + d], lb[r0:r0 + d]), ub[r0:r0 + d])
        elif typ == 1:
            t, x = v[r0], v[r0 + 1:r0 + d]
            nv = np.sqrt(np.sum(x * x))
            if nv <= t:
                pass
            elif nv <= -t:
                v[r0:r0 + d] = 0.0
            else:
                a = 0.5 * (t + nv)
                v[r0] = a
                v[r0 + 1:r0 + d] = x * (a / nv)
        else:
            nv = np.sqrt(np.sum(v[r0:r0 + d] ** 2))
            rad = ub[r0]
            if nv > rad:
                v[r0:r0 + d] *= rad / nv
    return v


def admm(prob, b, rho, sigma=1e-6, alpha=1.6, iters=20, ws=None, zs=None, ys=None):
    """Runs exactly `iters` iterations; returns (w, z, y, r_prim, r_dual) of the last one."""
    nx, nu, N, s = prob.nx, prob.nu, prob.N, prob.s
    coff, doff = prob.coff(), prob.doff()
    nct = prob.nc_total
    w = np.zeros(prob.ws_len) if ws is None else ws.copy()
    z = np.zeros(nct) if zs is None else zs.copy()
    y = np.zeros(nct) if ys is None else ys.copy()
    inv_rho = 1.0 / rho
    sol = O.OracleSolver(prob, b=b)
    cones_by_stage = [[c for c in prob.cones if c[0] == k] for k in range(N + 1)]
    r_prim = r_dual = 0.0
    for it in range(iters):
        sol.update_problem_data(w, y, z, inv_rho, sigma)
        if it == 0:
            sol.backward(rho)
        else:
            sol.backward_without_factorization(rho)
        wt = sol.forward(prob.x0[b], np.zeros(prob.ws_len))
        w_old = w
        w = alpha * wt + (1.0 - alpha) * w
        r_prim = r_dual = 0.0
        for k in range(N + 1):
            nc = int(prob.ncs[k])
            dim = s if k < N else nx
            # rows that are stationarity conditions: all of them, except those of x_0 (data, not a variable)
            rows = slice(0, nu) if k == 0 else slice(0, dim)
            wd = (wt - w_old)[k * s:k * s + dim]
            if nc == 0:
                r_dual = max(r_dual, float(np.max(np.abs(sigma * wd[rows]))) if wd[rows].size else 0.0)
                continue
            Dk = prob.D[b, doff[k]:doff[k + 1]].reshape(nc, dim, order="F")
            sl = slice(coff[k], coff[k + 1])
            zt = Dk @ wt[k * s:k * s + dim]
            zh = alpha * zt + (1.0 - alpha) * z[sl]
            znew = project(zh + y[sl] / rho[sl], cones_by_stage[k], prob.e_lb[b, sl], prob.e_ub[b, sl])
            y[sl] = y[sl] + rho[sl] * (zh - znew)
            # stationarity residual H w~ + h + D^T y + (dynamics multipliers) of the conic problem: the LQ solve makes the
            # augmented stationarity exact, so it equals  -(sigma (w~ - w_prev) + D^T rho ((1-alpha)(z~ - z_prev) + (z - z_prev)))
            rd = sigma * wd + Dk.T @ (rho[sl] * ((1.0 - alpha) * (zt - z[sl]) + (znew - z[sl])))
            r_dual = max(r_dual, float(np.max(np.abs(rd[rows]))))
            r_prim = max(r_prim, np.max(np.abs(zt - znew)))
            z[sl] = znew
    return w, z, y, r_prim, r_dual
