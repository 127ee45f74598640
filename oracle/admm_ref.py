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


def admm(prob, b, rho, sigma=1e-6, alpha=1.6, iters=20, ws=None, zs=None, ys=None, return_norms=False):
    """Runs exactly `iters` iterations (the first one factorises); returns (w, z, y, r_prim, r_dual) of the last one, with
    return_norms also (n_prim, n_dual) = (max |D w~|, |z|  ,  max |D^T y| over the stationarity rows): the scales of the
    convergence test and of the rho rule (admm_update_res_kernel)."""
    nx, nu, N, s = prob.nx, prob.nu, prob.N, prob.s
    coff, doff = prob.coff(), prob.doff()
    nct = prob.nc_total
    w = np.zeros(prob.ws_len) if ws is None else ws.copy()
    z = np.zeros(nct) if zs is None else zs.copy()
    y = np.zeros(nct) if ys is None else ys.copy()
    inv_rho = 1.0 / rho
    sol = O.OracleSolver(prob, b=b)
    cones_by_stage = [[c for c in prob.cones if c[0] == k] for k in range(N + 1)]
    r_prim = r_dual = n_prim = n_dual = 0.0
    for it in range(iters):
        sol.update_problem_data(w, y, z, inv_rho, sigma)
        if it == 0:
            sol.backward(rho)
        else:
            sol.backward_without_factorization(rho)
        wt = sol.forward(prob.x0[b], np.zeros(prob.ws_len))
        w_old = w
        w = alpha * wt + (1.0 - alpha) * w
        r_prim = r_dual = n_prim = n_dual = 0.0
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
            n_prim = max(n_prim, float(np.max(np.abs(zt))), float(np.max(np.abs(znew))))
            dty = (Dk.T @ y[sl])[rows]
            if dty.size:
                n_dual = max(n_dual, float(np.max(np.abs(dty))))
            z[sl] = znew
    if return_norms:
        return w, z, y, r_prim, r_dual, n_prim, n_dual
    return w, z, y, r_prim, r_dual


def admm_adaptive(prob, b, rho, sigma=1e-6, alpha=1.6, max_iter=1000, eps_abs=1e-4, eps_rel=1e-4, check_every=25,
                  rho_tau=5.0, max_rho_updates=10):
    """The device loop of pdplqr_admm_solve* with rho adaptation on (admm_ctl_kernel + the host part of the loop): every
    `check_every` iterations the convergence test  r <= eps_abs + eps_rel * n  and OSQP's rule  rho *= sqrt((r_prim / n_prim) /
    (r_dual / n_dual))  when that factor leaves [1 / tau, tau] (clamped to [1e-3, 1e3], rho to [1e-6, 1e6]); a rescale
    re-factorises.  Returns (w, z, y, iterations, rho_updates, (r_prim, r_dual), rho).  Every block of `check_every` iterations
    starts with a factorising solve here (same factors when rho did not change: differences at rounding level)."""
    rho = np.array(rho, dtype=np.float64, copy=True)
    w = z = y = None
    it, n_upd = 0, 0
    res = (0.0, 0.0)
    while it < max_iter:
        n = min(check_every - it % check_every, max_iter - it)
        w, z, y, rp, rd, npr, ndu = admm(prob, b, rho, sigma=sigma, alpha=alpha, iters=n, ws=w, zs=z, ys=y, return_norms=True)
        it += n
        res = (rp, rd)
        if rp <= eps_abs + eps_rel * npr and rd <= eps_abs + eps_rel * ndu:
            break
        if it < max_iter and n_upd < max_rho_updates:
            sc = np.sqrt((rp / max(npr, 1e-30)) / max(rd / max(ndu, 1e-30), 1e-30))
            if sc > rho_tau or sc < 1.0 / rho_tau:
                rho = np.clip(rho * min(max(sc, 1e-3), 1e3), 1e-6, 1e6)
                n_upd += 1
    return w, z, y, it, n_upd, res, rho
