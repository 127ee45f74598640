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
