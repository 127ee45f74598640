"""Independent numpy/scipy check: assemble the equality-constrained LQ KKT system and solve it with a sparse LU.
rho_dyn = 0 (default) is the exact system; rho_dyn = 1e-6 reproduces the semantics of the reference's QDLDLSolver, which
puts -rho_dyn I on the diagonal of the dynamics rows (kkt.hpp:196-203, qdldl_solver.hpp:40-42) -- the baseline the reference's
example prints next to the Riccati solvers (lqr_example.cpp:173-190), about 2e-5 away from the exact solution on config 1.
Variable order [w_0 .. w_{N-1}, x_N | lambda_1 .. lambda_N]; x_0 is fixed by its own (unperturbed) rows."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def augmented_cost(prob, b, ws_prev, sigma, ys=None, zs=None, rho=None, inv_rho=None):
    """H~_k = H + sigma I + D^T diag(rho) D ; h~_k = h - sigma w_prev - D^T (rho o (z - y/rho))
    (lqr_solver_parallel.hpp:129-137 + lqr_kernel.hpp:106-112)."""
    nx, nu, N, s = prob.nx, prob.nu, prob.N, prob.s
    coff, doff = prob.coff(), prob.doff()
    Hs, hs = [], []
    for k in range(N + 1):
        dim = s if k < N else nx
        Hk = (prob.H[b, k].reshape(dim, dim, order="F") if k < N else prob.HN[b].reshape(nx, nx, order="F")).copy()
        hk = (prob.h[b, k] if k < N else prob.hN[b]).copy()
        Hk += sigma * np.eye(dim)
        hk -= sigma * ws_prev[k * s:k * s + dim]
        nc = 0 if prob.ncs is None else int(prob.ncs[k])
        if nc > 0:
            Dk = prob.D[b, doff[k]:doff[k + 1]].reshape(nc, dim, order="F")
            r = rho[coff[k]:coff[k + 1]]
            g = zs[coff[k]:coff[k + 1]] - inv_rho[coff[k]:coff[k + 1]] * ys[coff[k]:coff[k + 1]]
            Hk += Dk.T @ (r[:, None] * Dk)
            hk -= Dk.T @ (r * g)
        Hs.append(Hk); hs.append(hk)
    return Hs, hs


def kkt_solve(prob, b=0, ws_prev=None, sigma=1e-6, ys=None, zs=None, rho=None, inv_rho=None, x0=None,
              return_costates=False, rho_dyn=0.0):
    nx, nu, N, s = prob.nx, prob.nu, prob.N, prob.s
    ws_prev = np.zeros(prob.ws_len) if ws_prev is None else ws_prev
    x0 = prob.x0[b] if x0 is None else x0
    Hs, hs = augmented_cost(prob, b, ws_prev, sigma, ys, zs, rho, inv_rho)
    nw = N * s + nx
    # constraints: x_{k+1} - E_k w_k = c_k (k=0..N-1), and x_0 = x0
    rows, cols, vals = [], [], []
    rhs_c = np.zeros(N * nx + nx)
    for k in range(N):
        Ek = prob.E[b, k].reshape(nx, s, order="F")
        r0 = k * nx
        for i in range(nx):
            for j in range(s):
                if Ek[i, j] != 0.0:
                    rows.append(r0 + i); cols.append(k * s + j); vals.append(-Ek[i, j])
            xcol = (k + 1) * s + nu + i if k + 1 < N else N * s + i
            rows.append(r0 + i); cols.append(xcol); vals.append(1.0)
        rhs_c[r0:r0 + nx] = prob.c[b, k]
    for i in range(nx):
        rows.append(N * nx + i); cols.append(nu + i); vals.append(1.0)
    rhs_c[N * nx:] = x0
    Cm = sp.csc_matrix((vals, (rows, cols)), shape=(N * nx + nx, nw))
    Hb = sp.block_diag([sp.csc_matrix(Hk) for Hk in Hs], format="csc")
    reg = None
    if rho_dyn != 0.0:   # QDLDLSolver semantics: -rho_dyn on the dynamics rows only (x_0 is eliminated exactly there)
        reg = sp.diags(np.concatenate([np.full(N * nx, -rho_dyn), np.zeros(nx)]), format="csc")
    K = sp.bmat([[Hb, Cm.T], [Cm, reg]], format="csc")
    rhs = np.concatenate([-np.concatenate(hs), rhs_c])
    lu = spla.splu(K)
    sol = lu.solve(rhs)
    for _ in range(2):   # iterative refinement: the 1e-9 parity checks must not be limited by the sparse LU itself
        sol = sol + lu.solve(rhs - K @ sol)
    if return_costates:   # multiplier mu_{k+1} of x_{k+1} - E_k w_k = c_k; the costate convention of pdplqr.h is -mu
        return sol[:nw], -sol[nw:nw + N * nx].reshape(N, nx)
    return sol[:nw]


def conic_kkt_violations(prob, b, w, z, y, rho, sigma=1e-6):
    """Optimality conditions of  min 1/2 w'Hw + h'w  s.t. dynamics, D_k w_k = z_k in K_k  at a candidate (w, z, y) -- the problem
    the conic outer iteration solves (not in the reference).  Returns the largest violation of each:
      link         |D w - z|                         dynamics     |x+ - E w - c|
      cone         distance-like violation of z in K (box: outside [lb, ub]; second-order cone: ||z_x|| - z_t; ball: ||z|| - radius)
      normal_cone  y in N_K(z): box: y > 0 only at ub, y < 0 only at lb (reported as |z - bound| of the offending rows);
                   second-order cone: ||y_x|| + y_t  (i.e. -y in K* = K) and |y'z|
      stationarity relative distance between w and the solution of the equality-constrained QP with (y, z, rho) folded in
                   (independent sparse KKT solve: at z = D w it is  H w + h + D'y + dynamics multipliers = 0)."""
    nx, nu, N, s = prob.nx, prob.nu, prob.N, prob.s
    coff, doff = prob.coff(), prob.doff()
    lb, ub = prob.e_lb[b], prob.e_ub[b]
    v = dict(link=0.0, dynamics=0.0, cone=0.0, normal_cone=0.0)
    for k in range(N + 1):
        dim = s if k < N else nx
        nc = int(prob.ncs[k])
        if nc:
            Dk = prob.D[b, doff[k]:doff[k + 1]].reshape(nc, dim, order="F")
            v["link"] = max(v["link"], float(np.max(np.abs(Dk @ w[k * s:k * s + dim] - z[coff[k]:coff[k + 1]]))))
        if k < N:
            Ek = prob.E[b, k].reshape(nx, s, order="F")
            xn = w[(k + 1) * s + nu:(k + 1) * s + nu + nx] if k + 1 < N else w[N * s:]
            v["dynamics"] = max(v["dynamics"], float(np.max(np.abs(Ek @ w[k * s:(k + 1) * s] + prob.c[b, k] - xn))))
    tol_act = 1e-7
    for (k, r0, d, typ) in prob.cones:
        sl = slice(coff[k] + r0, coff[k] + r0 + d)
        zz, yy = z[sl], y[sl]
        if typ == 0:
            v["cone"] = max(v["cone"], float(np.max(np.maximum(zz - ub[sl], 0.0))), float(np.max(np.maximum(lb[sl] - zz, 0.0))))
            up, lo = yy > tol_act, yy < -tol_act
            if np.any(up):
                v["normal_cone"] = max(v["normal_cone"], float(np.max(np.abs(zz[up] - ub[sl][up]))))
            if np.any(lo):
                v["normal_cone"] = max(v["normal_cone"], float(np.max(np.abs(zz[lo] - lb[sl][lo]))))
        elif typ == 1:
            v["cone"] = max(v["cone"], float(np.linalg.norm(zz[1:]) - zz[0]))
            v["normal_cone"] = max(v["normal_cone"], float(np.linalg.norm(yy[1:]) + yy[0]), abs(float(yy @ zz)))
        else:
            rad = ub[sl][0]
            nz = float(np.linalg.norm(zz))
            v["cone"] = max(v["cone"], nz - rad)
            # y = t z with t >= 0, and y = 0 strictly inside the ball
            if nz < rad - 1e-8:
                v["normal_cone"] = max(v["normal_cone"], float(np.max(np.abs(yy))))
            else:
                v["normal_cone"] = max(v["normal_cone"], float(np.linalg.norm(yy - (yy @ zz) / max(nz * nz, 1e-300) * zz)),
                                       max(0.0, -float(yy @ zz)))
    w_kkt = kkt_solve(prob, b=b, ws_prev=w, sigma=sigma, ys=y, zs=z, rho=rho, inv_rho=1.0 / rho)
    v["stationarity"] = float(np.max(np.abs(w_kkt - w)) / max(np.max(np.abs(w)), 1e-300))
    return v
