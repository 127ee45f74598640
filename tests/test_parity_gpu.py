"""GPU parity tests: the CUDA path, called through the C ABI (include/pdplqr.h via pdplqr_b200.LQRCudaSolver),
against the CPU oracle on identical seeded inputs.  Tolerance: 1e-9 relative (BASELINE.json north_star) --
states, controls and gains in FP64."""
import os

import numpy as np
import pytest

import pdplqr_b200 as P
from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gpu_solve(prob, S=1, lb=True, ws_in=None, sigma=1e-6, ctype=P.CHOLESKY):
    sol = P.LQRCudaSolver.from_problem(prob, num_segments=S, load_balancing=lb, solver_type=ctype)
    ws_in = prob.zeros_ws() if ws_in is None else ws_in
    sol.update_problem_data(ws_in, sigma=sigma)
    sol.backward()
    out = sol.forward(prob.x0, np.zeros_like(ws_in))
    return sol, out


def test_c1_example_matches_golden_and_oracle(oracle):
    """Config 1: examples/lqr_example.cpp as shipped (S = 4, load balancing, Cholesky)."""
    g = np.load(os.path.join(GOLD, "c1_quadrotor.npz"))
    p = P.problems.quadrotor_example()
    sol, ws = gpu_solve(p, S=4)
    assert list(sol.partition()[1]) == [21, 21, 21, 37]
    assert rel_err(ws[0], g["ws_seq"]) < TOL
    assert rel_err(ws[0], g["ws_kkt"]) < TOL
    for k in range(5):
        assert abs(abs(ws[0, k * 16]) - abs(g["survey_probe"][k])) < 1e-9
    xh, uh = sol.interface()
    assert rel_err(xh[0], g["xhat4"]) < TOL
    assert np.max(np.abs(uh[0, :3] - g["uhat4"][:3])) < TOL * 10.0   # |lam| ~ 1e-8 by cancellation of O(10) terms
    assert sol.last_status()[0] == 0


@pytest.mark.parametrize("S,lb", [(1, True), (2, True), (2, False), (4, False), (8, True), (8, False), (13, False)])
def test_c1_partitions_gains_and_interface(oracle, S, lb):
    p = P.problems.quadrotor_example()
    seq = oracle.OracleSolver(p).solve()
    sol, ws = gpu_solve(p, S=S, lb=lb)
    assert rel_err(ws[0], seq) < TOL
    o = oracle.OracleSolver(p, parallel=S > 1, num_segments=S, load_balancing=lb, condensed=oracle.LU)
    o.solve()
    if S > 1:
        assert list(sol.partition()[0]) == list(o.partition()[0])
    K, d, Gt = sol.gains()
    Ko, do, Gto = o.gains()
    assert rel_err(K[0], Ko) < TOL and rel_err(d[0], do) < TOL
    if S > 1:
        assert rel_err(Gt[0], Gto) < TOL
        xh, uh = sol.interface()
        xo, uo = o.interface()
        Ps, ps, Fs, fs, Cs = sol.summaries()
        # the interface costate lam = P x + p cancels to ~1e-8 here: its error scale is |P||x| + |p|, not |lam|
        lam_scale = max(1.0, float(np.max(np.abs(ps))))
        assert rel_err(xh[0], xo) < TOL and np.max(np.abs(uh[0, :S - 1] - uo[:S - 1])) < TOL * lam_scale
        for i in range(S):
            Po, po, Fo, fo, Co = o.summary(i)
            assert rel_err(Ps[0, i], Po) < TOL and rel_err(ps[0, i], po) < TOL
            if i < S - 1:
                assert rel_err(Fs[0, i], Fo) < TOL and rel_err(Cs[0, i], Co) < TOL
                assert np.max(np.abs(fs[0, i] - fo)) < TOL * max(1.0, np.max(np.abs(fo)))


@pytest.mark.parametrize("tree_lat", ["1", "0"])
@pytest.mark.parametrize("N,S", [(256, 16), (1024, 64), (1024, 128), (500, 100), (300, 33), (64, 2), (90, 3)])
def test_c2_quadrotor_ltv_multilevel_tree(oracle, N, S, tree_lat, monkeypatch):
    """Config 2 shape (nx=12, nu=4, LTV, single problem), many segments -> multi-level interface tree.
    tree_lat=1: latency-mode tree kernels (few combine groups); 0: the one-warp-per-combine throughput kernels."""
    monkeypatch.setenv("PDPLQR_TREE_LAT", tree_lat)
    p = P.problems.quadrotor_ltv(N)
    seq = oracle.OracleSolver(p).solve()
    sol, ws = gpu_solve(p, S=S, lb=False)
    assert rel_err(ws[0], seq) < TOL
    sol0, ws0 = gpu_solve(p, S=0)          # library-chosen segmentation
    assert rel_err(ws0[0], seq) < TOL
    assert sol.last_status()[0] == 0


@pytest.mark.parametrize("batch", [1, 33, 1000])
def test_c3_cartpole_batch_thread_path(oracle, batch):
    """Config 3 shape (nx=4, nu=1, N=128), thread-per-problem kernels, ragged batch sizes."""
    p = P.problems.cartpole_batch(batch=batch, N=128)
    ref, bad = oracle.OracleBatch(p).solve()
    assert bad == 0
    sol, ws = gpu_solve(p)
    assert ws.shape == ref.shape
    assert rel_err(ws, ref) < TOL
    for b in {0, batch // 2, batch - 1}:
        assert rel_err(ws[b], ref[b]) < TOL
    K, d, _ = sol.gains()
    o = oracle.OracleSolver(p, b=batch - 1)
    o.solve()
    Ko, do, _ = o.gains()
    assert rel_err(K[batch - 1], Ko) < TOL and rel_err(d[batch - 1], do) < TOL


def test_c3_batch_segmented_warp_path(oracle):
    """Same tiny systems through the warp-per-segment kernels (S > 1) with a batch dimension."""
    p = P.problems.cartpole_batch(batch=7, N=128)
    ref, _ = oracle.OracleBatch(p).solve()
    sol, ws = gpu_solve(p, S=4)
    assert rel_err(ws, ref) < TOL


@pytest.mark.parametrize("nx,nu", [(2, 1), (3, 2), (6, 3), (8, 8), (4, 1), (12, 4), (6, 2), (8, 4), (16, 4)])
@pytest.mark.parametrize("S", [1, 3])
def test_random_dense_problems_with_sigma_term(oracle, nx, nu, S):
    """Dense H with cross terms, nonzero h and c, sigma * w_prev active (update_problem_data fused in)."""
    p = P.problems.random_lq(nx, nu, 30, batch=3, seed=nx * 10 + nu)
    rng = np.random.default_rng(1)
    wprev = rng.standard_normal((p.batch, p.ws_len))
    sol, ws = gpu_solve(p, S=S, ws_in=wprev, sigma=0.05)
    for b in range(p.batch):
        ref = oracle.OracleSolver(p, b=b).solve(ws_in=wprev[b], sigma=0.05)
        assert rel_err(ws[b], ref) < TOL


def test_random_golden_fixture():
    g = np.load(os.path.join(GOLD, "random_6_3_40.npz"))
    q = P.problems.random_lq(6, 3, 40, batch=1, seed=11)
    for S in (1, 5):
        _, ws = gpu_solve(q, S=S, ws_in=g["wprev"][None].copy(), sigma=0.05)
        assert rel_err(ws[0], g["ws_seq"]) < TOL and rel_err(ws[0], g["ws_kkt"]) < TOL


@pytest.mark.parametrize("S", [1, 2])
def test_c4_dims_cta_path_unconstrained(oracle, S):
    """Config 4 dimensions (nx=30, nu=10): CTA-per-problem kernels (128 threads), unconstrained part."""
    p = P.problems.random_lq(30, 10, 24, batch=3, seed=5)
    sol, ws = gpu_solve(p, S=S)
    for b in range(p.batch):
        assert rel_err(ws[b], oracle.OracleSolver(p, b=b).solve()) < TOL


def test_repeated_solves_and_call_order():
    p = P.problems.quadrotor_example()
    sol = P.LQRCudaSolver.from_problem(p, num_segments=4)
    with pytest.raises(P.PdplqrError) as e:
        sol.backward()                      # update_problem_data must precede every backward
    assert e.value.code == P.capi.ERR_ORDER
    ws0 = p.zeros_ws()
    sol.update_problem_data(ws0, sigma=1e-6)
    sol.backward()
    a = sol.forward(p.x0, np.zeros_like(ws0)).copy()
    with pytest.raises(P.PdplqrError):
        sol.forward(p.x0, np.zeros_like(ws0))   # exactly one forward per backward
    sol.update_problem_data(ws0, sigma=1e-6)
    sol.backward()
    b = sol.forward(p.x0, np.zeros_like(ws0))
    assert np.array_equal(a, b)                 # deterministic
    out = sol.solve(ws0, p.x0, np.zeros_like(ws0))
    assert np.array_equal(a, out)
    assert sol.launch_count() > 0


def test_not_positive_definite_is_reported():
    p = P.problems.random_lq(4, 1, 10, batch=4, seed=9, dense_cost=False)
    p.H[2, :, 0] = -50.0          # R < 0 for problem 2  (H[.,.,0] is the (0,0) entry = R)
    sol, _ = gpu_solve(p)
    bad, st = sol.last_status()
    assert bad == 1 and st[2] != 0 and st[0] == 0


def test_device_pointer_variants_match_host_variants():
    import torch
    p = P.problems.cartpole_batch(batch=64, N=128)
    sol = P.LQRCudaSolver.from_problem(p)
    sol.set_stream(torch.cuda.current_stream().cuda_stream)
    ws_in = torch.zeros(p.batch, p.ws_len, dtype=torch.float64, device="cuda")
    x0 = torch.from_numpy(p.x0).cuda()
    out = torch.empty_like(ws_in)
    sol.update_problem_data_device(ws_in, sigma=1e-6)
    sol.backward_device()
    sol.forward_device(x0, out)
    torch.cuda.synchronize()
    _, ws = gpu_solve(p)
    assert np.array_equal(out.cpu().numpy(), ws)


# ------------------------------------------------------------------------------------------------------------
# constraint fold-in (a2/a3 with nc > 0) and backward_without_factorization (a6)
def _admm_vectors(p, seed):
    rng = np.random.default_rng(seed)
    nct = p.nc_total
    ys = rng.standard_normal((p.batch, nct))
    zs = rng.standard_normal((p.batch, nct))
    rho = rng.uniform(0.05, 2.0, (p.batch, nct))
    wprev = rng.standard_normal((p.batch, p.ws_len))
    return wprev, ys, zs, rho, np.ascontiguousarray(1.0 / rho)


@pytest.mark.parametrize("nx,nu,nc", [(4, 1, 3), (6, 3, 7), (12, 4, 16), (3, 2, 4), (30, 10, 44)])
@pytest.mark.parametrize("S", [1, 3])
def test_constraint_fold_in(oracle, nx, nu, nc, S):
    """H += D^T rho D, h -= D^T (rho o (z - y/rho))  (lqr_kernel.hpp:106-112) incl. the terminal stage."""
    from kkt_ref import kkt_solve
    p = P.problems.random_lq(nx, nu, 12, batch=2, seed=100 + nx, nc=nc)
    wprev, ys, zs, rho, inv_rho = _admm_vectors(p, 3)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=S)
    sol.update_problem_data(wprev, ys, zs, inv_rho, sigma=1e-3)
    sol.backward(rho)
    ws = sol.forward(p.x0, np.zeros_like(wprev))
    for b in range(p.batch):
        o = oracle.OracleSolver(p, b=b)
        o.update_problem_data(wprev[b], ys[b], zs[b], inv_rho[b], 1e-3)
        o.backward(rho[b])
        ref = o.forward(p.x0[b], np.zeros(p.ws_len))
        assert rel_err(ws[b], ref) < TOL
    assert rel_err(ws[0], kkt_solve(p, 0, wprev[0], 1e-3, ys[0], zs[0], rho[0], inv_rho[0])) < 1e-8


def test_c1_with_box_constraints_enabled(oracle):
    """The example with the constraints it disables by `nc = 0;` switched on (lqr_example.cpp:126-127,157-158):
    ncs = (nu, nx+nu, ..., nx) -- ragged constraint counts, identity D, rho = 0.01."""
    p = P.problems.quadrotor_example(constrained=True)
    rng = np.random.default_rng(8)
    nct = p.nc_total
    ys, zs = 0.1 * rng.standard_normal((1, nct)), 0.1 * rng.standard_normal((1, nct))
    rho = np.full((1, nct), 0.01)
    inv_rho = np.full((1, nct), 100.0)
    wprev = np.zeros((1, p.ws_len))
    o = oracle.OracleSolver(p)
    o.update_problem_data(wprev[0], ys[0], zs[0], inv_rho[0], 1e-6)
    o.backward(rho[0])
    ref = o.forward(p.x0[0], np.zeros(p.ws_len))
    for S in (1, 4):
        sol = P.LQRCudaSolver.from_problem(p, num_segments=S)
        ws = sol.solve(wprev, p.x0, np.zeros_like(wprev), sigma=1e-6, ys=ys, zs=zs, rho=rho, inv_rho=inv_rho)
        assert rel_err(ws[0], ref) < TOL


@pytest.mark.parametrize("nx,nu,nc,S", [(6, 3, 5, 1), (6, 3, 5, 4), (12, 4, 8, 3), (12, 4, 0, 5), (4, 1, 0, 1), (30, 10, 12, 2),
                                        (30, 10, 12, 1), (30, 10, 0, 3), (16, 4, 6, 1)])
def test_backward_without_factorization(oracle, nx, nu, nc, S):
    """Affine-only re-solve with cached factors (lqr_kernel.hpp:149-178, lqr_solver_parallel.hpp:148-154, :190-211):
    same protocol on the oracle and on the GPU; also against a fresh factorising solve of the second iterate."""
    p = P.problems.random_lq(nx, nu, 20, batch=2, seed=7 + nx + S, nc=nc)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=S)
    if nc == 0 and nx + nu > 8:
        sol.set_option(P.capi.OPT_AFFINE_CACHE, 1)
    its = []
    for seed in (1, 2, 3):
        if nc > 0:
            its.append(_admm_vectors(p, seed))
        else:
            its.append((np.random.default_rng(seed).standard_normal((p.batch, p.ws_len)), None, None, None, None))
    rho = its[0][3]
    outs = []
    for i, (w, y, z, _, _) in enumerate(its):
        inv = None if rho is None else np.ascontiguousarray(1.0 / rho)
        sol.update_problem_data(w, y, z, inv, sigma=1e-2)
        if i == 0:
            sol.backward(rho)
        else:
            sol.backward_without_factorization(rho)
        outs.append(sol.forward(p.x0, np.zeros_like(w)).copy())
    for b in range(p.batch):
        o = oracle.OracleSolver(p, b=b, parallel=S > 1, num_segments=S, condensed=oracle.LU)
        for i, (w, y, z, _, _) in enumerate(its):
            inv = None if rho is None else 1.0 / rho[b]
            o.update_problem_data(w[b], None if y is None else y[b], None if z is None else z[b], inv, 1e-2)
            if i == 0:
                o.backward(None if rho is None else rho[b])
            else:
                o.backward_without_factorization(None if rho is None else rho[b])
            ref = o.forward(p.x0[b], np.zeros(p.ws_len))
            assert rel_err(outs[i][b], ref) < TOL, (i, b)
    # a fresh factorising solve of the last iterate gives the same trajectory
    w, y, z, _, _ = its[-1]
    sol2 = P.LQRCudaSolver.from_problem(p, num_segments=S)
    inv = None if rho is None else np.ascontiguousarray(1.0 / rho)
    ws2 = sol2.solve(w, p.x0, np.zeros_like(w), sigma=1e-2, ys=y, zs=z, rho=rho, inv_rho=inv)
    assert rel_err(outs[-1], ws2) < TOL


def test_nofact_before_backward_is_an_order_error():
    p = P.problems.random_lq(6, 3, 10, seed=1)
    sol = P.LQRCudaSolver.from_problem(p)
    sol.update_problem_data(p.zeros_ws(), sigma=1e-6)
    with pytest.raises(P.PdplqrError) as e:
        sol.backward_without_factorization()
    assert e.value.code == P.capi.ERR_ORDER


@pytest.mark.parametrize("G,S_local", [(2, 1), (4, 4), (3, 0), (8, 16)])
def test_horizon_shards_on_one_gpu(oracle, G, S_local):
    """Horizon sharding (config 5 flow) emulated on ONE GPU: G handles own G time slices (interior shards +
    the terminal one), their slice summaries feed the coupler, every slice rolls out from its own boundary values.
    Same calls as sharding.HorizonShardedSolver, with the all_gather replaced by a torch.stack."""
    import torch
    from pdplqr_b200 import sharding
    from pdplqr_b200.solver import Coupler
    N = 512
    p = P.problems.quadrotor_ltv(N)
    ref = oracle.OracleSolver(p).solve()
    dev = torch.device("cuda", 0)
    x0 = torch.from_numpy(p.x0).to(dev)
    sols, outs, sums = [], [], []
    for r, (start, count) in enumerate(sharding.horizon_slices(N, G)):
        last = r == G - 1
        loc = sharding.slice_problem(p, start, count, last)
        s = P.LQRCudaSolver(p.nx, p.nu, count, num_segments=S_local, load_balancing=False)
        if not last:
            s.set_option(P.capi.OPT_INTERIOR_SHARD, 1)
        s.set_model(loc)
        s.update_problem_data_device(None, sigma=1e-6)
        s.backward_device()
        sm = torch.empty(1, s.summary_doubles(), dtype=torch.float64, device=dev)
        s.root_summary_device(sm)
        s.synchronize()
        sols.append(s); sums.append(sm[0])
    coup = Coupler(p.nx, p.nu, G)
    xhat = torch.empty(1, G, p.nx, dtype=torch.float64, device=dev)
    lam = torch.empty(1, G, p.nx, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    coup.solve_device(torch.stack(sums).unsqueeze(0).contiguous(), x0, xhat, lam)
    coup._lib.pdplqr_synchronize(coup._h)
    full = np.zeros(p.ws_len)
    costates = np.zeros((N, p.nx))
    for r, (start, count) in enumerate(sharding.horizon_slices(N, G)):
        out = torch.zeros(1, count * p.s + p.nx, dtype=torch.float64, device=dev)
        sols[r].set_root_boundary_device(xhat[:, r].contiguous(), lam[:, r].contiguous())
        sols[r].forward_device(x0, out)
        lam_loc = torch.zeros(1, count, p.nx, dtype=torch.float64, device=dev)
        sols[r].costates_device(out, lam_loc)          # local lambda_k = global lambda_{start + k}, k = 1..count
        sols[r].synchronize()
        o = out.cpu().numpy()[0]
        n = count * p.s + (p.nx if r == G - 1 else 0)
        full[start * p.s:start * p.s + n] = o[:n]
        costates[start:start + count] = lam_loc.cpu().numpy()[0]
    assert rel_err(full, ref) < TOL
    o_seq = oracle.OracleSolver(p)
    o_seq.solve()
    Pk, pk = o_seq.value()
    for k in (1, N // G, N // G + 1, N // 2, N - 1, N):   # lambda_k = P_k x_k + p_k (lqr_kernel.hpp:205-211)
        xk = ref[k * p.s + p.nu:(k + 1) * p.s] if k < N else ref[N * p.s:]
        lk = Pk[k].reshape(p.nx, p.nx, order="F") @ xk + pk[k]
        assert np.max(np.abs(costates[k - 1] - lk)) < 1e-8 * max(1.0, np.max(np.abs(lk)))
    xh_ref, _ = sharding.couple_numpy(torch.stack(sums).cpu().numpy(), p.x0[0])
    assert rel_err(xhat.cpu().numpy()[0], xh_ref) < TOL


# ------------------------------------------------------------------------------------------------------------
# conic ADMM outer iteration (row a11; NOT in the reference -> pinned against oracle/admm_ref.py)
@pytest.mark.parametrize("S", [1, 3])
def test_admm_iterates_match_numpy_restatement(oracle, S):
    from oracle import admm_ref
    p = P.problems.random_lq(6, 3, 14, batch=2, seed=31, nc=6)
    # mix cone types: box on rows 0..2, second-order cone on rows 3..5 (terminal stage: all box)
    p.cones = []
    for k in range(p.N + 1):
        nc = int(p.ncs[k])
        p.cones += [(k, 0, 3, 0), (k, 3, nc - 3, 1)] if k < p.N else [(k, 0, nc, 0)]
    rho = np.full((p.batch, p.nc_total), 0.7)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=S)
    sol.admm_set_cones(p.cones, p.e_lb, p.e_ub)
    ws, zs, ys = p.zeros_ws(), np.zeros((p.batch, p.nc_total)), np.zeros((p.batch, p.nc_total))
    iters, res = sol.admm_solve(p.x0, ws, zs, ys, rho, sigma=1e-4, alpha=1.6, max_iter=15, eps_abs=0.0, eps_rel=0.0,
                                check_every=5)
    assert iters == 15
    rp = rd = 0.0
    for b in range(p.batch):
        w, z, y, r_prim, r_dual = admm_ref.admm(p, b, rho[b], sigma=1e-4, alpha=1.6, iters=15)
        assert rel_err(ws[b], w) < TOL and rel_err(zs[b], z) < TOL
        assert np.max(np.abs(ys[b] - y)) < TOL * max(1.0, np.max(np.abs(y)))
        rp, rd = max(rp, r_prim), max(rd, r_dual)
    assert abs(res[0] - rp) < 1e-9 * max(1.0, rp) and abs(res[1] - rd) < 1e-9 * max(1.0, rd)


def test_admm_quadrotor_mpc_converges_to_feasible_point(oracle):
    """Config-1 problem with its box constraints switched on (lqr_example.cpp:126-158 minus `nc = 0;`), rho = 0.01
    as in the example... scaled up for faster convergence; checks feasibility and agreement with the numpy restatement."""
    from oracle import admm_ref
    p = P.problems.quadrotor_example(N=20, constrained=True)
    p.x0[0, 2] = 0.0
    lb = np.where(np.isfinite(p.e_lb), p.e_lb, -1e20)
    ub = np.where(np.isfinite(p.e_ub), p.e_ub, 1e20)
    rho = np.full((1, p.nc_total), 0.1)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=2)
    sol.admm_set_cones(p.cones, lb, ub)
    ws, zs, ys = p.zeros_ws(), np.zeros((1, p.nc_total)), np.zeros((1, p.nc_total))
    iters, res = sol.admm_solve(p.x0, ws, zs, ys, rho, sigma=1e-6, alpha=1.6, max_iter=300, eps_abs=0.0, eps_rel=0.0,
                                check_every=50)
    assert iters == 300 and res[0] < 5e-2
    u = ws[0, :20 * 16].reshape(20, 16)[:, :4]
    assert np.max(np.abs(u - np.clip(u, -0.9916, 2.4084))) < 5e-2        # ADMM iterate: feasible up to the residual
    assert np.all(zs >= lb - 1e-12) and np.all(zs <= ub + 1e-12)         # z is exactly feasible by construction
    p.e_lb, p.e_ub = lb, ub
    w, z, y, r_prim, _ = admm_ref.admm(p, 0, rho[0], sigma=1e-6, alpha=1.6, iters=300)
    assert rel_err(ws[0], w) < 1e-8 and abs(res[0] - r_prim) < 1e-8


def test_pipelined_host_solve_matches_unpipelined(oracle):
    """pdplqr_solve() on a large thread-path batch overlaps H2D / kernels / D2H over batch chunks; results must be
    bit-identical to the 3-call protocol (ragged last chunk included)."""
    p = P.problems.cartpole_batch(batch=5000, N=32)
    rng = np.random.default_rng(3)
    w = 0.01 * rng.standard_normal((p.batch, p.ws_len))
    sol = P.LQRCudaSolver.from_problem(p)
    a = sol.solve(w, p.x0, np.zeros_like(w), sigma=1e-3).copy()
    sol.update_problem_data(w, sigma=1e-3)
    sol.backward()
    b = sol.forward(p.x0, np.zeros_like(w))
    assert np.array_equal(a, b)
    ref, _ = oracle.OracleBatch(p.select(slice(4990, 5000))).solve(ws_in=np.ascontiguousarray(w[4990:]), sigma=1e-3)
    assert rel_err(a[4990:], ref) < TOL
    assert sol.last_status()[0] == 0


@pytest.mark.parametrize("sparse", ["1", "0"])
def test_selection_matrix_constraints_sparse_and_dense_paths(oracle, sparse, monkeypatch):
    """Constraint matrices whose rows have at most one non-zero (box / cone rows on single variables: the reference
    example's own constraint set, lqr_example.cpp:133-147, and config 4) take a structure-exploiting path (no dense D
    in the kernels); PDPLQR_SPARSE_D=0 forces the dense path.  Both must match the oracle (dense, as the reference)."""
    from oracle import admm_ref
    monkeypatch.setenv("PDPLQR_SPARSE_D", sparse)
    p = P.problems.random_conic_batch(batch=3, N=10, nx=6, nu=3, seed=5, soc=False)
    # scale some rows so that values other than 1 are exercised
    p.D[:, ::7] *= 1.5
    rng = np.random.default_rng(4)
    nct = p.nc_total
    wprev, ys, zs = rng.standard_normal((3, p.ws_len)), rng.standard_normal((3, nct)), rng.standard_normal((3, nct))
    rho = rng.uniform(0.1, 2.0, (3, nct))
    inv = np.ascontiguousarray(1.0 / rho)
    for S in (1, 3):
        sol = P.LQRCudaSolver.from_problem(p, num_segments=S)
        outs = []
        for it in range(2):
            sol.update_problem_data(wprev * (it + 1), ys, zs, inv, sigma=1e-3)
            (sol.backward if it == 0 else sol.backward_without_factorization)(rho)
            outs.append(sol.forward(p.x0, np.zeros_like(wprev)).copy())
        for b in range(3):
            o = oracle.OracleSolver(p, b=b)
            for it in range(2):
                o.update_problem_data(wprev[b] * (it + 1), ys[b], zs[b], inv[b], 1e-3)
                (o.backward if it == 0 else o.backward_without_factorization)(rho[b])
                assert rel_err(outs[it][b], o.forward(p.x0[b], np.zeros(p.ws_len))) < TOL
    # and through the ADMM loop (projection kernel uses the same compact form)
    lb = np.where(np.isfinite(p.e_lb), p.e_lb, -1e20); ub = np.where(np.isfinite(p.e_ub), p.e_ub, 1e20)
    p.e_lb, p.e_ub = lb, ub
    sol = P.LQRCudaSolver.from_problem(p, num_segments=1)
    sol.admm_set_cones(p.cones, lb, ub)
    ws, z, y = p.zeros_ws(), np.zeros((3, nct)), np.zeros((3, nct))
    r2 = np.full((3, nct), 0.5)
    sol.admm_solve(p.x0, ws, z, y, r2, sigma=1e-4, alpha=1.6, max_iter=8, eps_abs=0.0, eps_rel=0.0, check_every=8)
    w_ref, z_ref, _, _, _ = admm_ref.admm(p, 1, r2[1], sigma=1e-4, alpha=1.6, iters=8)
    assert rel_err(ws[1], w_ref) < TOL and rel_err(z[1], z_ref) < TOL


@pytest.mark.parametrize("nx,nu,N,S,batch", [(4, 1, 1, 1, 5), (4, 1, 2, 1, 40), (4, 1, 3, 3, 2), (12, 4, 1, 1, 1), (12, 4, 2, 2, 1),
                                             (12, 4, 5, 5, 2), (6, 3, 7, 0, 3), (3, 2, 129, 1, 70)])
def test_edge_horizons_and_segment_lengths(oracle, nx, nu, N, S, batch):
    """Minimum horizon (LQRModel requires N >= 1, lqr_model.hpp:75-77), one-stage segments, thread-path ring shorter
    than its depth, trajectory chunks that do not divide N."""
    p = P.problems.random_lq(nx, nu, N, batch=batch, seed=N * 7 + nx)
    rng = np.random.default_rng(N)
    w = rng.standard_normal((batch, p.ws_len))
    sol, ws = gpu_solve(p, S=S, lb=False, ws_in=w, sigma=0.02)
    for b in range(batch):
        assert rel_err(ws[b], oracle.OracleSolver(p, b=b).solve(ws_in=w[b], sigma=0.02)) < TOL


def test_ragged_constraint_counts_with_empty_stages(oracle):
    """ncs varies per stage and is zero on some stages (LQRModel::ncs, lqr_model.hpp:71; the reference example itself
    uses nu rows at k=0, nx+nu inside, nx at k=N)."""
    nx, nu, N = 6, 3, 9
    p = P.problems.random_lq(nx, nu, N, batch=2, seed=13)
    ncs = np.array([3, 0, 5, 0, 0, 9, 1, 0, 4, 2], np.int32)
    rng = np.random.default_rng(2)
    dims = np.array([nx + nu] * N + [nx])
    dtot = int(np.sum(ncs * dims))
    p.ncs = ncs
    p.D = rng.standard_normal((2, dtot))
    nct = int(ncs.sum())
    ys, zs = rng.standard_normal((2, nct)), rng.standard_normal((2, nct))
    rho = rng.uniform(0.1, 1.0, (2, nct))
    inv = np.ascontiguousarray(1.0 / rho)
    w = rng.standard_normal((2, p.ws_len))
    for S in (1, 3):
        sol = P.LQRCudaSolver.from_problem(p, num_segments=S)
        out = sol.solve(w, p.x0, np.zeros_like(w), sigma=1e-2, ys=ys, zs=zs, rho=rho, inv_rho=inv)
        for b in range(2):
            o = oracle.OracleSolver(p, b=b)
            o.update_problem_data(w[b], ys[b], zs[b], inv[b], 1e-2)
            o.backward(rho[b])
            assert rel_err(out[b], o.forward(p.x0[b], np.zeros(p.ws_len))) < TOL


def test_invalid_arguments_are_rejected():
    with pytest.raises(P.PdplqrError) as e:
        P.LQRCudaSolver(12, 4, 0)                      # N >= 1  (lqr_model.hpp:75-77)
    assert e.value.code == P.capi.ERR_INVALID
    P.LQRCudaSolver(5, 5, 10).close()                  # not an instantiated pair: padded to one (lqr_model.hpp:66-89)
    with pytest.raises(P.PdplqrError) as e:
        P.LQRCudaSolver(64, 5, 10)                     # beyond the largest instantiated kernel size
    assert e.value.code == P.capi.ERR_UNSUPPORTED
    with pytest.raises(P.PdplqrError):
        P.LQRCudaSolver(12, 4, 10, solver_type=7)      # unknown condensed solver type (lqr_solver_parallel.hpp:98-99)
    p = P.problems.random_lq(6, 3, 8, seed=1, nc=4)
    sol = P.LQRCudaSolver.from_problem(p)
    with pytest.raises(P.PdplqrError):
        sol.update_problem_data(p.zeros_ws(), sigma=1e-6)   # ys / zs / inv_rho missing for a constrained model


def test_equal_split_partition_mode(oracle):
    """load_balancing = 2 (addition): num_segments equal parts instead of the reference rule's long last segment."""
    p = P.problems.quadrotor_ltv(203)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=10, load_balancing=2)
    st, ln = sol.partition()
    assert list(ln) == [21, 21, 21] + [20] * 7 and int(st[-1] + ln[-1]) == 203
    ws = sol.solve(p.zeros_ws(), p.x0, p.zeros_ws())
    assert rel_err(ws[0], oracle.OracleSolver(p).solve()) < TOL


@pytest.mark.parametrize("seg_t", ["64", "128"])
@pytest.mark.parametrize("nx,nu,nc", [(12, 4, 0), (12, 4, 8), (6, 3, 0)])
def test_stage_kernel_group_sizes(oracle, nx, nu, nc, seg_t, monkeypatch):
    """PDPLQR_SEG_T: two and four warps per (problem, segment) instead of the default single warp (throughput mode:
    more groups than the latency-mode threshold), with and without the constraint fold-in."""
    monkeypatch.setenv("PDPLQR_SEG_T", seg_t)
    p = P.problems.random_lq(nx, nu, 24, batch=110, seed=5 + nc, nc=nc)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=3)
    assert sol.num_segments * p.batch > 2 * 148
    rng = np.random.default_rng(11)
    wprev = rng.standard_normal((p.batch, p.ws_len))
    picks = (0, 57, 109)
    if nc:
        _, ys, zs, rho, inv_rho = _admm_vectors(p, 4)
        sol.update_problem_data(wprev, ys, zs, inv_rho, sigma=0.01)
        sol.backward(rho)
    else:
        sol.update_problem_data(wprev, sigma=0.01)
        sol.backward()
    ws = sol.forward(p.x0, np.zeros_like(wprev))
    for b in picks:
        o = oracle.OracleSolver(p, b=b)
        if nc:
            o.update_problem_data(wprev[b], ys[b], zs[b], inv_rho[b], 0.01)
            o.backward(rho[b])
            ref = o.forward(p.x0[b], np.zeros(p.ws_len))
        else:
            ref = o.solve(ws_in=wprev[b], sigma=0.01)
        assert rel_err(ws[b], ref) < TOL


# ------------------------------------------------------------------------------------------------------------
# costate recovery (commented out in the reference, lqr_kernel.hpp:205-211)
@pytest.mark.parametrize("nx,nu,N,S", [(12, 4, 40, 1), (12, 4, 40, 4), (6, 3, 30, 3), (16, 4, 24, 2), (3, 2, 17, 5),
                                       (4, 1, 33, 1), (3, 2, 20, 1), (5, 2, 21, 1), (7, 3, 19, 3)])
def test_costates_match_kkt_multipliers_and_value_function(oracle, nx, nu, N, S):
    """lambda_k vs (i) the multipliers of an independent sparse KKT solve and (ii) the reference's own (commented)
    formula lambda_k = P_k x_k + p_k with P_k = Lxx Lxx^T, p_k from the sequential oracle."""
    from kkt_ref import kkt_solve
    p = P.problems.random_lq(nx, nu, N, batch=2, seed=40 + nx)
    rng = np.random.default_rng(2)
    wprev = rng.standard_normal((p.batch, p.ws_len))
    sol, ws = gpu_solve(p, S=S, ws_in=wprev, sigma=0.05)
    lam = sol.costates(ws)
    for b in range(p.batch):
        w_kkt, lam_kkt = kkt_solve(p, b, wprev[b], 0.05, return_costates=True)
        assert rel_err(ws[b], w_kkt) < TOL
        assert rel_err(lam[b], lam_kkt) < TOL
        o = oracle.OracleSolver(p, b=b)
        ref = o.solve(ws_in=wprev[b], sigma=0.05)
        lam_ref = o.costates(ref)                      # lambda_k = Lxx (Lxx^T x_k) + p_k, all k
        assert np.max(np.abs(lam[b] - lam_ref)) < TOL * max(1.0, np.max(np.abs(lam_ref)))


@pytest.mark.parametrize("S", [1, 3])
def test_costates_with_constraint_fold_in(oracle, S):
    """Constrained stages (dense D, ragged counts incl. the terminal stage): the costates satisfy the KKT system of the
    ADMM-augmented problem."""
    from kkt_ref import kkt_solve
    p = P.problems.random_lq(6, 3, 14, batch=2, seed=77, nc=5)
    wprev, ys, zs, rho, inv_rho = _admm_vectors(p, 9)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=S)
    sol.update_problem_data(wprev, ys, zs, inv_rho, sigma=1e-3)
    sol.backward(rho)
    ws = sol.forward(p.x0, np.zeros_like(wprev))
    lam = sol.costates(ws)
    for b in range(p.batch):
        w_kkt, lam_kkt = kkt_solve(p, b, wprev[b], 1e-3, ys[b], zs[b], rho[b], inv_rho[b], return_costates=True)
        assert rel_err(ws[b], w_kkt) < TOL and rel_err(lam[b], lam_kkt) < TOL


def test_costates_call_order_and_thread_path(oracle):
    p = P.problems.random_lq(12, 4, 10, batch=1, seed=3)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=2)
    with pytest.raises(P.PdplqrError):
        sol.costates(p.zeros_ws())                      # nothing solved yet
    q = P.problems.cartpole_batch(batch=40, N=16)       # thread-per-problem path (ragged last tile: 40 = 32 + 8)
    sol2, ws2 = gpu_solve(q)
    lam = sol2.costates(ws2)
    for b in (0, 31, 32, 39):
        o = oracle.OracleSolver(q, b=b)
        ref = o.solve()
        assert rel_err(ws2[b], ref) < TOL
        lam_ref = o.costates(ref)
        assert np.max(np.abs(lam[b] - lam_ref)) < TOL * max(1.0, np.max(np.abs(lam_ref)))
