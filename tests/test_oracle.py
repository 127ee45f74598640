"""CPU tests pinning the oracle (oracle/pdp_oracle.cpp): the reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned by (i) committed golden fixtures, (ii) sequential == parallel for
every partition / condensed variant, (iii) an independent sparse KKT solve (rho_dyn = 0), (iv) KKT residuals."""
import os

import numpy as np
import pytest

import pdplqr_b200 as P
from conftest import rel_err
from kkt_ref import augmented_cost, conic_kkt_violations, kkt_solve

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_c1_golden_and_survey_probe(oracle):
    g = np.load(os.path.join(GOLD, "c1_quadrotor.npz"))
    p = P.problems.quadrotor_example()
    ws = oracle.OracleSolver(p).solve()
    assert rel_err(ws, g["ws_seq"]) < 1e-13
    assert rel_err(ws, g["ws_kkt"]) < 1e-12
    # first five controls (lqr_example.cpp:206-209 prints them); sign pattern -+-+ ; survey probe values
    for k in range(5):
        u = ws[k * 16:k * 16 + 4]
        assert abs(abs(u[0]) - abs(g["survey_probe"][k])) < 1e-10
        assert np.allclose(u, u[0] * np.array([1, -1, 1, -1]), atol=1e-12)
    xN = ws[-12:]
    assert abs(xN[2] - 0.9999999) < 1e-9 and np.max(np.abs(np.delete(xN, 2))) < 1e-13


def test_c1_qdldl_semantics_printout(oracle, capsys):
    """SURVEY.md section 8 (f3), optional part: the reference's example prints its QDLDL baseline next to the Riccati solvers
    (lqr_example.cpp:173-190).  QDLDLSolver regularises the dynamics rows with rho_dyn = 1e-6 (qdldl_solver.hpp:40-42,
    kkt.hpp:196-203), so its solution is NOT the exact one: reproduce that system and print the like-for-like comparison
    (the survey's probe: 1.9e-5 between the two on config 1; exact vs the Riccati solvers: 2e-15)."""
    p = P.problems.quadrotor_example()
    ws = oracle.OracleSolver(p).solve()
    exact = kkt_solve(p, rho_dyn=0.0)
    qdldl = kkt_solve(p, rho_dyn=1e-6)
    e_exact, e_qdldl = rel_err(ws, exact), rel_err(ws, qdldl)
    with capsys.disabled():
        print("\nconfig 1 (lqr_example.cpp): Riccati vs exact KKT (rho_dyn = 0): %.2e | vs QDLDLSolver semantics "
              "(rho_dyn = 1e-6): %.2e" % (e_exact, e_qdldl))
        for k in range(5):   # the five controls the example prints (lqr_example.cpp:206-209)
            print("  u_%d  riccati % .12f   qdldl-semantics % .12f" % (k, ws[k * 16], qdldl[k * 16]))
    assert e_exact < 1e-12
    assert 1e-6 < e_qdldl < 1e-4


@pytest.mark.parametrize("S", [2, 4, 8])
@pytest.mark.parametrize("ctype", [0, 1])
@pytest.mark.parametrize("lb", [True, False])
def test_parallel_equals_sequential(oracle, S, ctype, lb):
    p = P.problems.quadrotor_example()
    seq = oracle.OracleSolver(p).solve()
    o = oracle.OracleSolver(p, parallel=True, num_segments=S, load_balancing=lb, condensed=ctype, nthreads=4)
    par = o.solve()
    assert o.status() == 0
    assert rel_err(par, seq) < 1e-12


def test_partition_rule(oracle):
    # lqr_solver_parallel.hpp:70-80 ; SURVEY.md section 3.1 probe values
    p = P.problems.quadrotor_example()
    for S, want in ((2, [39, 61]), (4, [21, 21, 21, 37]), (8, [11] * 7 + [23])):
        o = oracle.OracleSolver(p, parallel=True, num_segments=S)
        assert list(o.partition()[1]) == want


def test_random_golden_and_kkt(oracle):
    g = np.load(os.path.join(GOLD, "random_6_3_40.npz"))
    q = P.problems.random_lq(6, 3, 40, batch=1, seed=11)
    ws = oracle.OracleSolver(q).solve(ws_in=g["wprev"], sigma=0.05)
    assert rel_err(ws, g["ws_seq"]) < 1e-13
    assert rel_err(ws, g["ws_kkt"]) < 1e-11
    for S in (2, 5):
        for ct in (0, 1):
            par = oracle.OracleSolver(q, parallel=True, num_segments=S, condensed=ct).solve(ws_in=g["wprev"], sigma=0.05)
            assert rel_err(par, ws) < 1e-11


def test_constrained_fold_in_matches_kkt(oracle):
    """a2/a3 with nc > 0: H += D^T rho D, h -= D^T (rho o (z - y/rho))  (lqr_kernel.hpp:106-112)."""
    q = P.problems.random_lq(4, 2, 12, batch=1, seed=3, nc=5)
    rng = np.random.default_rng(0)
    nct = q.nc_total
    ys, zs = rng.standard_normal(nct), rng.standard_normal(nct)
    rho = rng.uniform(0.05, 2.0, nct)
    inv_rho = 1.0 / rho
    wprev = rng.standard_normal(q.ws_len)
    ref = kkt_solve(q, 0, wprev, 1e-3, ys, zs, rho, inv_rho)
    for par, S in ((False, 1), (True, 3)):
        o = oracle.OracleSolver(q, parallel=par, num_segments=S, condensed=0)
        o.update_problem_data(wprev, ys, zs, inv_rho, 1e-3)
        o.backward(rho)
        ws = o.forward(q.x0[0], np.zeros(q.ws_len))
        assert rel_err(ws, ref) < 1e-10


def test_backward_without_factorization(oracle):
    """a6: affine-only re-solve with cached factors equals a full re-solve when only h/ws/z/y changed."""
    q = P.problems.random_lq(4, 2, 16, batch=1, seed=4, nc=3)
    rng = np.random.default_rng(1)
    nct = q.nc_total
    rho = rng.uniform(0.1, 1.0, nct)
    inv_rho = 1.0 / rho
    for par, S in ((False, 1), (True, 4)):
        o = oracle.OracleSolver(q, parallel=par, num_segments=S, condensed=1 if par else 0)
        w1, y1, z1 = rng.standard_normal(q.ws_len), rng.standard_normal(nct), rng.standard_normal(nct)
        o.update_problem_data(w1, y1, z1, inv_rho, 1e-2)
        o.backward(rho)
        o.forward(q.x0[0], np.zeros(q.ws_len))
        w2, y2, z2 = rng.standard_normal(q.ws_len), rng.standard_normal(nct), rng.standard_normal(nct)
        o.update_problem_data(w2, y2, z2, inv_rho, 1e-2)
        o.backward_without_factorization(rho)
        fast = o.forward(q.x0[0], np.zeros(q.ws_len))
        ref = kkt_solve(q, 0, w2, 1e-2, y2, z2, rho, inv_rho)
        assert rel_err(fast, ref) < 1e-10


def test_kkt_residuals_c2_small(oracle):
    """Dynamics + stationarity residual of the returned trajectory (SURVEY.md section 4 pin iv)."""
    p = P.problems.quadrotor_ltv(64)
    o = oracle.OracleSolver(p, parallel=True, num_segments=4)
    ws = o.solve()
    nx, nu, s, N = p.nx, p.nu, p.s, p.N
    x = np.zeros((N + 1, nx))
    for k in range(N):
        x[k] = ws[k * s + nu:(k + 1) * s]
    x[N] = ws[N * s:]
    dyn = 0.0
    for k in range(N):
        Ek = p.E[0, k].reshape(nx, s, order="F")
        dyn = max(dyn, np.max(np.abs(x[k + 1] - Ek @ ws[k * s:(k + 1) * s] - p.c[0, k])))
    assert dyn < 1e-12
    assert rel_err(ws, kkt_solve(p)) < 1e-10


def test_batch_pool_matches_single(oracle):
    p = P.problems.cartpole_batch(batch=16, N=32)
    ws, bad = oracle.OracleBatch(p).solve(nthreads=2)
    assert bad == 0
    for b in (0, 7, 15):
        assert rel_err(ws[b], oracle.OracleSolver(p, b=b).solve()) < 1e-13
        assert rel_err(ws[b], kkt_solve(p, b)) < 1e-9


def test_augmented_cost_helper():
    q = P.problems.random_lq(3, 2, 4, seed=2)
    Hs, hs = augmented_cost(q, 0, np.ones(q.ws_len), 0.5)
    assert Hs[0].shape == (5, 5) and Hs[-1].shape == (3, 3)
    assert np.allclose(hs[0], q.h[0, 0] - 0.5)


@pytest.mark.parametrize("S,condensed", [(2, "LU"), (4, "CHOLESKY"), (5, "LU")])
def test_oracle_costates_sequential_parallel_and_kkt(oracle, S, condensed):
    """Costates by the reference's commented formula (lqr_kernel.hpp:205-211, lqr_kernel_parallel.hpp:207-216):
    sequential == multipliers of the exact KKT system; segment-parallel (+ F^T uhat inside non-last segments) ==
    sequential."""
    from kkt_ref import kkt_solve
    p = P.problems.random_lq(6, 3, 25, batch=1, seed=91)
    rng = np.random.default_rng(4)
    wprev = rng.standard_normal(p.ws_len)
    seq = oracle.OracleSolver(p)
    ws = seq.solve(ws_in=wprev, sigma=0.02)
    lam = seq.costates(ws)
    w_kkt, lam_kkt = kkt_solve(p, 0, wprev, 0.02, return_costates=True)
    assert np.max(np.abs(ws - w_kkt)) < 1e-10
    assert np.max(np.abs(lam - lam_kkt)) < 1e-9 * max(1.0, np.max(np.abs(lam_kkt)))
    par = oracle.OracleSolver(p, parallel=True, num_segments=S, load_balancing=False,
                              condensed=getattr(oracle, condensed))
    wp = par.solve(ws_in=wprev, sigma=0.02)
    assert np.max(np.abs(wp - ws)) < 1e-10
    assert np.max(np.abs(par.costates(wp) - lam)) < 1e-9 * max(1.0, np.max(np.abs(lam)))


def test_admm_dual_residual_identity_equals_stationarity_definition(oracle):
    """The dual residual the ADMM kernel reports is computed from stage-local differences,
        sigma (w~ - w_prev) + D^T rho ((1 - alpha)(z~ - z_prev) + (z - z_prev)),
    which must equal the DEFINITION: the stationarity residual H w~ + h + D^T y + (dynamics multipliers) of the conic
    problem, evaluated here with the multipliers of an independent sparse KKT solve of the same LQ sub-problem."""
    from kkt_ref import kkt_solve
    from oracle import admm_ref
    p = P.problems.random_lq(5, 2, 9, batch=1, seed=12, nc=4)
    nx, nu, N, s = p.nx, p.nu, p.N, p.s
    rho = np.full(p.nc_total, 0.8)
    coff, doff = p.coff(), p.doff()
    for alpha in (1.0, 1.6):
        K = 6
        w0, z0, y0, _, _ = admm_ref.admm(p, 0, rho, sigma=1e-2, alpha=alpha, iters=K - 1)
        _, _, _, _, r_dual = admm_ref.admm(p, 0, rho, sigma=1e-2, alpha=alpha, iters=K)
        wt, lam = kkt_solve(p, 0, w0, 1e-2, y0, z0, rho, 1.0 / rho, return_costates=True)   # K-th LQ solve + costates
        cones = [[c for c in p.cones if c[0] == k] for k in range(N + 1)]
        worst = 0.0
        for k in range(N + 1):
            dim = s if k < N else nx
            nc = int(p.ncs[k])
            Dk = p.D[0, doff[k]:doff[k + 1]].reshape(nc, dim, order="F")
            sl = slice(coff[k], coff[k + 1])
            wk = wt[k * s:k * s + dim]
            zt = Dk @ wk
            zh = alpha * zt + (1.0 - alpha) * z0[sl]
            zn = admm_ref.project(zh + y0[sl] / rho[sl], cones[k], p.e_lb[0, sl], p.e_ub[0, sl])
            yn = y0[sl] + rho[sl] * (zh - zn)
            Hk = (p.H[0, k].reshape(s, s, order="F") if k < N else p.HN[0].reshape(nx, nx, order="F"))
            hk = p.h[0, k] if k < N else p.hN[0]
            v = Hk @ wk + hk + Dk.T @ yn
            if k < N:
                v = v + p.E[0, k].reshape(nx, s, order="F").T @ lam[k]       # E_k^T lambda_{k+1}
            if k >= 1:
                v[dim - nx:] -= lam[k - 1]                                      # - lambda_k on the rows of x_k
            rows = slice(0, nu) if k == 0 else slice(0, dim)
            worst = max(worst, float(np.max(np.abs(v[rows]))))
        assert abs(worst - r_dual) < 1e-9 * max(1.0, worst), (alpha, worst, r_dual)


def test_admm_restatement_converges_to_a_kkt_point_of_the_conic_qp(oracle):
    """Row a11 has no reference to pin against (the outer iteration is not in the reference), so oracle/admm_ref.py -- the
    restatement the CUDA loop is compared with -- is pinned against the optimality conditions of the problem it solves:
    box-constrained quadrotor (the example's own bounds), run to convergence, then checked WITHOUT the Riccati oracle:
    primal feasibility, stationarity through the independent sparse KKT solve, dual sign / complementarity per row, and
    independence of the fixed point from rho."""
    from oracle import admm_ref
    p = P.problems.quadrotor_example(N=20, constrained=True)
    nct, s, nx, N = p.nc_total, p.s, p.nx, p.N
    lb, ub = p.e_lb[0], p.e_ub[0]
    sols = {}
    for rho0 in (1.0, 10.0):
        rho = np.full(nct, rho0)
        w, z, y, r_prim, r_dual = admm_ref.admm(p, 0, rho, sigma=1e-6, alpha=1.6, iters=1500)
        assert r_prim < 1e-10 and r_dual < 1e-9
        sols[rho0] = (w, z, y, rho)
    w, z, y, rho = sols[1.0]
    # (i) primal feasibility: z = D w inside the box, dynamics exact
    coff, doff = p.coff(), p.doff()
    for k in range(N + 1):
        dim = s if k < N else nx
        nc = int(p.ncs[k])
        if nc:
            Dk = p.D[0, doff[k]:doff[k + 1]].reshape(nc, dim, order="F")
            assert np.max(np.abs(Dk @ w[k * s:k * s + dim] - z[coff[k]:coff[k + 1]])) < 1e-9
        if k < N:
            Ek = p.E[0, k].reshape(nx, s, order="F")
            xn = w[(k + 1) * s + p.nu:(k + 1) * s + p.nu + nx] if k + 1 < N else w[N * s:]
            assert np.max(np.abs(Ek @ w[k * s:(k + 1) * s] + p.c[0, k] - xn)) < 1e-10
    assert np.all(z <= ub + 1e-9) and np.all(z >= lb - 1e-9)
    # (ii) stationarity of the ORIGINAL problem, H w + h + D^T y + (dynamics multipliers) = 0: at z = D w the augmented data
    #      reduce to it, so the independent sparse KKT solve with (w, y, z) folded in must return w itself
    w_kkt = kkt_solve(p, ws_prev=w, sigma=1e-6, ys=y, zs=z, rho=rho, inv_rho=1.0 / rho)
    assert rel_err(w_kkt, w) < 1e-8
    # (iii) dual feasibility and complementarity of the box rows: y > 0 only at the upper bound, y < 0 only at the lower one
    act = np.abs(y) > 1e-7
    assert 0 < np.sum(act) < nct                                    # the bounds bite somewhere, not everywhere
    assert np.all(np.abs(z[y > 1e-7] - ub[y > 1e-7]) < 1e-8)
    assert np.all(np.abs(z[y < -1e-7] - lb[y < -1e-7]) < 1e-8)
    # (iv) the fixed point does not depend on the penalty
    w10, z10, y10, _ = sols[10.0]
    assert rel_err(w10, w) < 1e-7 and rel_err(y10, y) < 1e-6


def test_admm_restatement_second_order_cone_optimality(oracle):
    """The same pin for second-order cones (the C4 problem family at a short horizon, box + SOC rows): at convergence every
    cone's z lies in K, the multiplier in the normal cone of K at z (-y in K = K*, y^T z = 0), and (w, y, z) satisfy the
    stationarity of the original problem through the independent sparse KKT solve."""
    from oracle import admm_ref
    p = P.problems.random_conic_batch(batch=1, N=12, seed=5)
    nct = p.nc_total
    rho = np.full(nct, 10.0)
    w, z, y, r_prim, r_dual = admm_ref.admm(p, 0, rho, sigma=1e-6, alpha=1.6, iters=3000)
    assert r_prim < 1e-10 and r_dual < 1e-9
    coff = p.coff()
    n_soc = n_apex_or_boundary = 0
    for (k, r0, d, typ) in p.cones:
        sl = slice(coff[k] + r0, coff[k] + r0 + d)
        if typ == 0:
            assert np.all(z[sl] <= p.e_ub[0, sl] + 1e-9) and np.all(z[sl] >= p.e_lb[0, sl] - 1e-9)
            assert np.all(np.abs(z[sl][y[sl] > 1e-7] - p.e_ub[0, sl][y[sl] > 1e-7]) < 1e-8)
            assert np.all(np.abs(z[sl][y[sl] < -1e-7] - p.e_lb[0, sl][y[sl] < -1e-7]) < 1e-8)
        elif typ == 1:
            n_soc += 1
            zt, zx, yt, yx = z[sl][0], z[sl][1:], y[sl][0], y[sl][1:]
            assert zt >= np.linalg.norm(zx) - 1e-9                  # z in K
            assert -yt >= np.linalg.norm(yx) - 1e-8                 # -y in K* = K
            assert abs(float(y[sl] @ z[sl])) < 1e-8                 # complementarity
            n_apex_or_boundary += (zt - np.linalg.norm(zx)) < 1e-8
    assert n_soc >= 10 and n_apex_or_boundary >= 1                  # the cones are there, and at least one of them is active
    w_kkt = kkt_solve(p, ws_prev=w, sigma=1e-6, ys=y, zs=z, rho=rho, inv_rho=1.0 / rho)
    assert rel_err(w_kkt, w) < 1e-8


def test_admm_adaptive_restatement_rescales_and_reaches_the_same_optimum(oracle):
    """oracle/admm_ref.py::admm_adaptive (the restatement of the device loop with OSQP's rho rule, which the GPU test
    test_admm_rho_adaptation_matches_the_numpy_restatement compares iteration counts with): from a rho that is far too small it
    rescales, stops on its own test well before the iteration cap, and ends at the optimum a well-scaled fixed rho gives."""
    from oracle import admm_ref
    p = P.problems.quadrotor_example(N=20, constrained=True)
    p.x0[0, 2] = 0.0
    nct = p.nc_total
    w, z, y, it, n_upd, res, rho_end = admm_ref.admm_adaptive(p, 0, np.full(nct, 1e-3), sigma=1e-6, alpha=1.6, max_iter=4000,
                                                               eps_abs=1e-5, eps_rel=1e-5, check_every=25, rho_tau=5.0,
                                                               max_rho_updates=6)
    assert it % 25 == 0 and it < 4000 and 1 <= n_upd <= 6 and rho_end[0] > 1e-3
    w_opt, _, _, rp, rd = admm_ref.admm(p, 0, np.full(nct, 1.0), sigma=1e-6, alpha=1.6, iters=1500)
    assert rp < 1e-10 and rd < 1e-9
    assert rel_err(w, w_opt) < 1e-3
    # with the same small rho held fixed the primal residual is still far from the tolerance after as many iterations
    _, _, _, rp_fixed, _ = admm_ref.admm(p, 0, np.full(nct, 1e-3), sigma=1e-6, alpha=1.6, iters=it)
    assert rp_fixed > 100 * res[0]


@pytest.mark.parametrize("family", ["quadrotor-box", "conic-soc"])
def test_conic_kkt_checker_on_the_restatement(oracle, family):
    """tests/kkt_ref.py::conic_kkt_violations (the checker the GPU test test_admm_solution_satisfies_the_conic_kkt_conditions
    applies to the CUDA result) on the converged numpy restatement, and its sensitivity: a perturbed multiplier or iterate is flagged."""
    from oracle import admm_ref
    if family == "quadrotor-box":
        p, rho0, iters = P.problems.quadrotor_example(N=20, constrained=True), 1.0, 1500
    else:
        p, rho0, iters = P.problems.random_conic_batch(batch=1, N=12, seed=5), 10.0, 3000
    rho = np.full(p.nc_total, rho0)
    w, z, y, _, _ = admm_ref.admm(p, 0, rho, sigma=1e-6, alpha=1.6, iters=iters)
    v = conic_kkt_violations(p, 0, w, z, y, rho)
    assert v["link"] < 1e-9 and v["dynamics"] < 1e-10 and v["cone"] < 1e-9 and v["normal_cone"] < 1e-7 and v["stationarity"] < 1e-8, v
    act = np.argmax(np.abs(y))
    y_bad = y.copy(); y_bad[act] *= 1.01
    assert conic_kkt_violations(p, 0, w, z, y_bad, rho)["stationarity"] > 1e-6
    w_bad = w.copy(); w_bad[p.nu + 1] += 1e-3       # a state entry of stage 0 ... which is data: dynamics then fail at stage 0
    assert conic_kkt_violations(p, 0, w_bad, z, y, rho)["dynamics"] > 1e-6
