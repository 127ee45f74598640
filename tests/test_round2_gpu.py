"""Round-2 GPU parity tests (through the C ABI): parity at BENCHMARK size and shape, arbitrary (nx, nu) by padding,
the real multi-process NCCL horizon-sharded path, constrained horizon shards, ADMM at the C4 dimensions.
Tolerance: 1e-9 relative (BASELINE.json north_star)."""
import os
import sys

import numpy as np
import pytest

import pdplqr_b200 as P
from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ---------------------------------------------------------------------------------- benchmark size and shape
@pytest.mark.parametrize("logN", [17, 20])
def test_long_horizon_at_benchmark_segmentation(oracle, logN):
    """C5 shape: nx12/nu4, N = 2^17 and the benchmarked 2^20, the bench's wave-aligned equal segmentation
    (load_balancing = 2; whole waves of ~200-stage segments and a 13-level interface tree at 2^20) against the SEQUENTIAL oracle over the
    whole horizon -- checks that 1e-9 survives the tree depth and the interface conditioning."""
    bench = _bench()
    N = 1 << logN
    p = P.problems.quadrotor_ltv(N)
    S = bench.c5_segments(1) if logN == 20 else bench.wave_aligned(N // 64, bench.stage_wave())
    rng = np.random.default_rng(17)
    wprev = 0.01 * rng.standard_normal((1, p.ws_len))
    sol = P.LQRCudaSolver.from_problem(p, num_segments=S, load_balancing=2)
    assert sol.num_segments == S
    ws = sol.solve(wprev, p.x0, np.empty_like(wprev), sigma=1e-6)
    assert sol.last_status()[0] == 0
    ref = oracle.OracleSolver(p, parallel=False).solve(ws_in=wprev[0].copy(), sigma=1e-6)
    assert rel_err(ws[0], ref) < TOL
    # per-stage check as well (a max-norm over 16 M numbers would hide a locally wrong stage with small entries)
    e = np.abs(ws[0, :N * 16] - ref[:N * 16]).reshape(N, 16).max(axis=1)
    scale = np.maximum(np.abs(ref[:N * 16]).reshape(N, 16).max(axis=1), 1e-3)
    assert float(np.max(e / scale)) < 1e-7
    del sol


def test_c3_at_benchmark_batch(oracle):
    """C3 at the benchmarked 65,536 x (nx4, nu1, N128): sampled problems incl. the first and the last tile of 32, and
    the pipelined host-buffer call the bench's e2e number goes through."""
    B = 65536
    q = P.problems.cartpole_batch(batch=B, N=128, seed=1234)
    rng = np.random.default_rng(17)
    wprev = 0.01 * rng.standard_normal((B, q.ws_len))
    sol = P.LQRCudaSolver.from_problem(q)
    ws = sol.solve(wprev, q.x0, np.empty_like(wprev), sigma=1e-6)
    assert sol.last_status()[0] == 0
    pick = np.unique(np.concatenate([np.arange(40), np.arange(B - 40, B), rng.integers(0, B, 120)]))
    sub = q.select(pick)
    ref, bad = oracle.OracleBatch(sub).solve(ws_in=np.ascontiguousarray(wprev[pick]), sigma=1e-6)
    assert bad == 0
    for i, b in enumerate(pick):
        assert rel_err(ws[b], ref[i]) < TOL, b


# ---------------------------------------------------------------------------------- arbitrary (nx, nu)
@pytest.mark.parametrize("nx,nu,N,S,batch", [(5, 5, 30, 1, 3), (5, 5, 30, 4, 2), (7, 2, 25, 3, 2), (13, 3, 40, 5, 1),
                                             (20, 6, 24, 2, 2), (3, 1, 20, 1, 70), (1, 1, 9, 1, 5), (9, 9, 12, 2, 1)])
def test_arbitrary_dimensions_by_padding(oracle, nx, nu, N, S, batch):
    """(nx, nu) pairs that are not instantiated run on the cheapest instantiated pair that contains them (zero /
    identity padding): states, controls, gains, interface values and costates equal the oracle's on the
    caller's dimensions (the reference takes any n, m at run time, lqr_model.hpp:66-89)."""
    p = P.problems.random_lq(nx, nu, N, batch=batch, seed=7 * nx + nu)
    rng = np.random.default_rng(3)
    wprev = rng.standard_normal((batch, p.ws_len))
    sol = P.LQRCudaSolver.from_problem(p, num_segments=S, load_balancing=False)
    sol.update_problem_data(wprev, sigma=0.02)
    sol.backward()
    ws = sol.forward(p.x0, np.zeros_like(wprev))
    K, d, Gt = sol.gains()
    lam = sol.costates(ws)
    for b in range(min(batch, 3)):
        o = oracle.OracleSolver(p, b=b, parallel=S > 1, num_segments=S, load_balancing=False, condensed=oracle.LU)
        ref = o.solve(ws_in=wprev[b], sigma=0.02)
        assert rel_err(ws[b], ref) < TOL
        Ko, do, Gto = o.gains()
        assert rel_err(K[b], Ko) < TOL and rel_err(d[b], do) < TOL
        if S > 1:
            assert rel_err(Gt[b], Gto) < TOL
            xh, uh = sol.interface()
            xo, uo = o.interface()
            assert rel_err(xh[b], xo) < TOL
        seq = oracle.OracleSolver(p, b=b)
        sref = seq.solve(ws_in=wprev[b], sigma=0.02)
        lam_ref = seq.costates(sref)
        assert np.max(np.abs(lam[b] - lam_ref)) < TOL * max(1.0, np.max(np.abs(lam_ref)))


@pytest.mark.parametrize("nx,nu,nc,S", [(5, 3, 4, 1), (7, 2, 6, 3), (13, 3, 9, 2)])
def test_arbitrary_dimensions_with_constraints_and_no_refactor(oracle, nx, nu, nc, S):
    """Padded sizes through the constraint fold-in (dense D with zero columns for the padded variables) and the
    backward_without_factorization protocol."""
    p = P.problems.random_lq(nx, nu, 14, batch=2, seed=50 + nx, nc=nc)
    rng = np.random.default_rng(4)
    nct = p.nc_total
    rho = rng.uniform(0.5, 2.0, (p.batch, nct))
    inv_rho = np.ascontiguousarray(1.0 / rho)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=S, load_balancing=False)
    os_ = [oracle.OracleSolver(p, b=b) for b in range(p.batch)]
    for it in range(3):
        wprev = rng.standard_normal((p.batch, p.ws_len))
        ys, zs = rng.standard_normal((p.batch, nct)), rng.standard_normal((p.batch, nct))
        sol.update_problem_data(wprev, ys, zs, inv_rho, sigma=1e-3)
        if it == 0:
            sol.backward(rho)
        else:
            sol.backward_without_factorization(rho)
        ws = sol.forward(p.x0, np.zeros_like(wprev))
        for b in range(p.batch):
            os_[b].update_problem_data(wprev[b], ys[b], zs[b], inv_rho[b], 1e-3)
            if it == 0:
                os_[b].backward(rho[b])
            else:
                os_[b].backward_without_factorization(rho[b])
            assert rel_err(ws[b], os_[b].forward(p.x0[b], np.zeros(p.ws_len))) < TOL


def test_device_pointer_api_with_padded_dimensions(oracle):
    import torch
    p = P.problems.random_lq(5, 2, 33, batch=4, seed=9)
    dev = torch.device("cuda", 0)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=3)
    wprev = torch.from_numpy(np.random.default_rng(1).standard_normal((4, p.ws_len))).to(dev)
    out = torch.empty_like(wprev)
    x0 = torch.from_numpy(p.x0).to(dev)
    sol.update_problem_data_device(wprev, sigma=0.1)
    sol.backward_device()
    sol.forward_device(x0, out)
    sol.synchronize()
    got, wp = out.cpu().numpy(), wprev.cpu().numpy()
    for b in range(4):
        assert rel_err(got[b], oracle.OracleSolver(p, b=b).solve(ws_in=wp[b], sigma=0.1)) < TOL


# ---------------------------------------------------------------------------------- horizon shards
def _sharded_emulation(p, G, S_local, wprev, sigma, ys=None, zs=None, rho=None, inv_rho=None):
    """G handles own G time slices on ONE GPU (interior shards + the terminal one); same calls as
    sharding.HorizonShardedSolver with the all_gather replaced by a torch.stack."""
    import torch
    from pdplqr_b200 import sharding
    from pdplqr_b200.solver import Coupler
    dev = torch.device("cuda", 0)
    N = p.N
    x0 = torch.from_numpy(p.x0).to(dev)
    sols, sums, locs = [], [], []
    slices = sharding.horizon_slices(N, G)
    for r, (start, count) in enumerate(slices):
        last = r == G - 1
        loc = sharding.slice_problem(p, start, count, last)
        s = P.LQRCudaSolver(p.nx, p.nu, count, num_segments=S_local, load_balancing=False, ncs=loc.ncs)
        if not last:
            s.set_option(P.capi.OPT_INTERIOR_SHARD, 1)
        s.set_model(loc)
        lo = start * p.s
        wp = np.ascontiguousarray(wprev[:, lo:lo + count * p.s + p.nx])
        kw = {}
        if loc.ncs is not None:
            c0, c1 = loc.con_slice
            kw = dict(ys=np.ascontiguousarray(ys[:, c0:c1]), zs=np.ascontiguousarray(zs[:, c0:c1]),
                      inv_rho_vecs=np.ascontiguousarray(inv_rho[:, c0:c1]))
        s.update_problem_data(wp, sigma=sigma, **kw)
        s.backward(None if loc.ncs is None else np.ascontiguousarray(rho[:, loc.con_slice[0]:loc.con_slice[1]]))
        sm = torch.empty(1, s.summary_doubles(), dtype=torch.float64, device=dev)
        s.root_summary_device(sm)
        s.synchronize()
        sols.append(s); sums.append(sm[0]); locs.append(loc)
    coup = Coupler(p.nx, p.nu, G)
    xhat = torch.empty(1, G, p.nx, dtype=torch.float64, device=dev)
    lam = torch.empty(1, G, p.nx, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    coup.solve_device(torch.stack(sums).unsqueeze(0).contiguous(), x0, xhat, lam)
    coup._lib.pdplqr_synchronize(coup._h)
    full = np.zeros(p.ws_len)
    for r, (start, count) in enumerate(slices):
        out = torch.zeros(1, count * p.s + p.nx, dtype=torch.float64, device=dev)
        sols[r].set_root_boundary_device(xhat[0, r], lam[0, r])
        sols[r].forward_device(x0, out)
        sols[r].synchronize()
        o = out.cpu().numpy()[0]
        n = count * p.s + (p.nx if r == G - 1 else 0)
        full[start * p.s:start * p.s + n] = o[:n]
    return full, sols


@pytest.mark.parametrize("nx,nu,nc,G,S_local", [(12, 4, 8, 3, 2), (6, 3, 5, 4, 1), (7, 2, 4, 2, 3)])
def test_horizon_shards_with_constraints(oracle, nx, nu, nc, G, S_local):
    """Constrained problems split into time slices (the constraint rows travel with their stages; an interior slice has
    no terminal rows) -- incl. a padded (nx, nu)."""
    p = P.problems.random_lq(nx, nu, 36, batch=1, seed=21 + nx, nc=nc)
    rng = np.random.default_rng(6)
    nct = p.nc_total
    wprev = rng.standard_normal((1, p.ws_len))
    ys, zs = rng.standard_normal((1, nct)), rng.standard_normal((1, nct))
    rho = rng.uniform(0.5, 2.0, (1, nct))
    inv_rho = np.ascontiguousarray(1.0 / rho)
    full, _ = _sharded_emulation(p, G, S_local, wprev, 1e-3, ys, zs, rho, inv_rho)
    o = oracle.OracleSolver(p)
    o.update_problem_data(wprev[0], ys[0], zs[0], inv_rho[0], 1e-3)
    o.backward(rho[0])
    assert rel_err(full, o.forward(p.x0[0], np.zeros(p.ws_len))) < TOL


def test_root_boundary_is_consumed_by_forward(oracle):
    """A root boundary set for one solve must not leak into the next: a terminal-slice handle reused as an ordinary
    solver rolls out from the caller's x0 again; an interior shard refuses to roll out without a fresh boundary."""
    import torch
    p = P.problems.quadrotor_ltv(64)
    dev = torch.device("cuda", 0)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=4)
    ref = oracle.OracleSolver(p).solve()
    x0 = torch.from_numpy(p.x0).to(dev)
    out = torch.zeros(1, p.ws_len, dtype=torch.float64, device=dev)
    other = torch.ones(1, p.nx, dtype=torch.float64, device=dev)
    sol.update_problem_data_device(None, sigma=1e-6); sol.backward_device()
    sol.set_root_boundary_device(other, None)
    sol.forward_device(x0, out); sol.synchronize()
    assert rel_err(out.cpu().numpy()[0], ref) > 1e-3          # rolled out from the boundary state, not from x0
    sol.update_problem_data_device(None, sigma=1e-6); sol.backward_device()
    sol.forward_device(x0, out); sol.synchronize()
    assert rel_err(out.cpu().numpy()[0], ref) < TOL           # boundary consumed: the caller's x0 again
    shard = P.LQRCudaSolver(p.nx, p.nu, 64, num_segments=2)
    shard.set_option(P.capi.OPT_INTERIOR_SHARD, 1)
    shard.set_model(p)
    shard.update_problem_data_device(None, sigma=1e-6); shard.backward_device()
    with pytest.raises(P.PdplqrError) as e:
        shard.forward_device(x0, out)
    assert e.value.code == P.capi.ERR_ORDER


# ---------------------------------------------------------------------------------- real multi-process NCCL path
def _nccl_worker(rank, world, port, q, N, nseg):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    try:
        import torch
        import torch.distributed as dist
        import pdplqr_b200 as P2
        from pdplqr_b200 import sharding as sh
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        start, count = sh.horizon_slices(N, world)[rank]
        local = P2.problems.quadrotor_ltv(N, start=start, count=count)       # every rank builds ONLY its slice
        side = torch.cuda.Stream()                                           # a non-default stream: order must still hold
        with torch.cuda.stream(side):
            hs = sh.HorizonShardedSolver(None, rank, world, num_segments=nseg, device=rank, local=local)
            wfull = 0.01 * np.random.default_rng(17).standard_normal((1, N * 16 + 12))
            ws = torch.from_numpy(np.ascontiguousarray(wfull[:, start * 16:(start + count) * 16 + 12])).to(dev)
            out = torch.zeros_like(ws)
            for _ in range(3):                                               # repeated solves back to back, no host syncs
                hs.solve_device(ws, 1e-6, out)
        torch.cuda.synchronize()
        q.put((rank, start, count, out.cpu().numpy()[0]))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # report instead of hanging the parent
        import traceback
        q.put((rank, -1, -1, traceback.format_exc() + repr(e)))


@pytest.mark.parametrize("world", [2, 4])
def test_horizon_sharded_solver_nccl_multiprocess(oracle, world):
    """sharding.HorizonShardedSolver.solve_device as bench.py runs it: one process per GPU, NCCL all_gather of the slice
    summaries, every rank generating only its own time slice -- against the sequential oracle over the whole horizon."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    N, nseg = 1 << 14, 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, 29700 + world, q, N, nseg)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=120)
    p = P.problems.quadrotor_ltv(N)
    wfull = 0.01 * np.random.default_rng(17).standard_normal((1, N * 16 + 12))
    ref = oracle.OracleSolver(p).solve(ws_in=wfull[0].copy(), sigma=1e-6)
    full = np.zeros(p.ws_len)
    for rank, start, count, o in res:
        assert start >= 0, o
        n = count * 16 + (12 if rank == world - 1 else 0)
        full[start * 16:start * 16 + n] = o[:n]
    assert rel_err(full, ref) < TOL


# ---------------------------------------------------------------------------------- ADMM at the C4 dimensions
def test_admm_box_and_soc_at_c4_dimensions(oracle):
    """The conic outer iteration with box + second-order-cone rows at the C4 dimensions (nx30, nu10, nc = 44 per
    interior stage) against oracle/admm_ref.py (NOT in the reference: parity unpinned by construction)."""
    from oracle import admm_ref
    p = P.problems.random_conic_batch(batch=3, N=12, seed=5)
    assert (p.nx, p.nu, int(p.ncs[1])) == (30, 10, 44) and any(c[3] == 1 for c in p.cones)
    lb = np.where(np.isfinite(p.e_lb), p.e_lb, -1e20)
    ub = np.where(np.isfinite(p.e_ub), p.e_ub, 1e20)
    rho = np.full((p.batch, p.nc_total), 0.1)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=1)
    sol.admm_set_cones(p.cones, lb, ub)
    ws, zs, ys = p.zeros_ws(), np.zeros((p.batch, p.nc_total)), np.zeros((p.batch, p.nc_total))
    iters, res = sol.admm_solve(p.x0, ws, zs, ys, rho, sigma=1e-6, alpha=1.6, max_iter=25, eps_abs=0.0, eps_rel=0.0,
                                check_every=25)
    assert iters == 25
    p.e_lb, p.e_ub = lb, ub
    for b in range(p.batch):
        w, z, y, r_prim, r_dual = admm_ref.admm(p, b, rho[b], sigma=1e-6, alpha=1.6, iters=25)
        assert rel_err(ws[b], w) < TOL and rel_err(zs[b], z) < TOL
        assert np.max(np.abs(ys[b] - y)) < TOL * max(1.0, np.max(np.abs(y)))
    # the SOC rows really are inside their cones after the projection
    coff = p.coff()
    for (k, r0, d, typ) in p.cones:
        if typ == 1:
            v = zs[0, coff[k] + r0:coff[k] + r0 + d]
            assert np.linalg.norm(v[1:]) <= v[0] + 1e-12


# ---------------------------------------------------------------------------------- ADMM as one CUDA graph (row f1)
def _cone_problem(nx, nu, N, batch, seed, nc):
    p = P.problems.random_lq(nx, nu, N, batch=batch, seed=seed, nc=nc)
    p.cones = []
    for k in range(p.N + 1):
        n = int(p.ncs[k])
        p.cones += [(k, 0, 2, 0), (k, 2, n - 2, 1)] if k < p.N and n > 3 else [(k, 0, n, 0)]
    return p


@pytest.mark.parametrize("nx,nu,S", [(6, 3, 1), (6, 3, 3), (12, 4, 2), (5, 2, 1), (7, 3, 2)])
def test_admm_graph_equals_host_loop_and_numpy(oracle, nx, nu, S):
    """One conic solve = ONE CUDA graph launch (factorising iteration + WHILE node over affine-only iterations, convergence
    test on the device); same iterates as the host-issued loop and as oracle/admm_ref.py, incl. padded (nx, nu)."""
    from oracle import admm_ref
    p = _cone_problem(nx, nu, 12, 2, 31 + nx, 6)
    rho = np.full((p.batch, p.nc_total), 0.7)
    res = {}
    for mode in ("graph", "host"):
        sol = P.LQRCudaSolver.from_problem(p, num_segments=S)
        sol.admm_set_cones(p.cones, p.e_lb, p.e_ub)
        sol.admm_configure(use_graph=(mode == "graph"))
        ws, zs, ys = p.zeros_ws(), np.zeros((p.batch, p.nc_total)), np.zeros((p.batch, p.nc_total))
        l0 = sol.launch_count()
        iters, r = sol.admm_solve(p.x0, ws, zs, ys, rho, sigma=1e-4, alpha=1.6, max_iter=17, eps_abs=0.0, eps_rel=0.0,
                                  check_every=5)
        res[mode] = (iters, r.copy(), ws, zs, ys, sol.launch_count() - l0, sol.admm_stats())
    (ig, rg, wg, zg, yg, lg, sg), (ih, rh, wh, zh, yh, lh, sh) = res["graph"], res["host"]
    assert ig == ih == 17 and sg[0] == 1 and sh[0] == 0          # exactly one graph launch for the whole solve
    assert np.array_equal(wg, wh) and np.array_equal(zg, zh) and np.array_equal(yg, yh) and np.array_equal(rg, rh)
    assert lg == lh                                                # same kernels ran, counted from the captured graph
    rp = rd = 0.0
    for b in range(p.batch):
        w, z, y, r_prim, r_dual = admm_ref.admm(p, b, rho[b], sigma=1e-4, alpha=1.6, iters=17)
        assert rel_err(wg[b], w) < TOL and rel_err(zg[b], z) < TOL
        rp, rd = max(rp, r_prim), max(rd, r_dual)
    assert abs(rg[0] - rp) < 1e-9 * max(1.0, rp) and abs(rg[1] - rd) < 1e-9 * max(1.0, rd)


def test_admm_device_convergence_test_and_rho_adaptation(oracle):
    """Early exit decided on the device (the graph's WHILE condition), and the OSQP rho rule: a badly scaled rho is
    rescaled (re-factorisation, one more graph launch per rescale) and the solve then converges in fewer iterations."""
    p = P.problems.quadrotor_example(N=20, constrained=True)
    p.x0[0, 2] = 0.0
    lb = np.where(np.isfinite(p.e_lb), p.e_lb, -1e20)
    ub = np.where(np.isfinite(p.e_ub), p.e_ub, 1e20)
    out = {}
    for adaptive in (False, True):
        sol = P.LQRCudaSolver.from_problem(p, num_segments=2)
        sol.admm_set_cones(p.cones, lb, ub)
        sol.admm_configure(use_graph=True, adaptive_rho=adaptive, rho_tau=5.0, max_rho_updates=6)
        rho = np.full((1, p.nc_total), 1e-3)          # far too small: the primal residual dominates
        ws, zs, ys = p.zeros_ws(), np.zeros((1, p.nc_total)), np.zeros((1, p.nc_total))
        iters, r = sol.admm_solve(p.x0, ws, zs, ys, rho, sigma=1e-6, alpha=1.6, max_iter=4000, eps_abs=1e-5, eps_rel=1e-5,
                                  check_every=25)
        out[adaptive] = (iters, r, sol.admm_stats(), ws.copy())
    it0, r0, st0, w0 = out[False]
    it1, r1, st1, w1 = out[True]
    assert it0 % 25 == 0 and it1 % 25 == 0                       # exits happen on check iterations
    assert st0 == (1, 0) and st1[1] >= 1 and st1[0] == 1 + st1[1]   # one launch, plus one per rescale
    assert it1 < it0 and it1 < 4000                              # adaptation pays, and the device test stopped the loop
    assert r1[0] <= 1e-5 + 1e-5 * 10 and rel_err(w1, w0) < 1e-2 or it0 == 4000


# ---------------------------------------------------------------------------------- graph-launched solve, status slots
@pytest.mark.parametrize("nx,nu,N,S,nc", [(12, 4, 256, 0, 0), (6, 3, 40, 4, 5), (5, 2, 30, 3, 0), (4, 1, 64, 1, 0)])
def test_solve_device_graph_equals_protocol_calls(oracle, nx, nu, N, S, nc):
    """pdplqr_solve_device (one CUDA graph launch: update_problem_data + backward + forward) on a side stream gives the
    same bits as the three protocol calls, on repeated launches and after a pointer change (re-capture)."""
    import torch
    p = P.problems.quadrotor_ltv(N) if (nx, nu) == (12, 4) else P.problems.random_lq(nx, nu, N, batch=3, seed=nx, nc=nc)
    dev = torch.device("cuda", 0)
    side = torch.cuda.Stream()
    rng = np.random.default_rng(2)
    T = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    wprev = T(rng.standard_normal((p.batch, p.ws_len)))
    kw = {}
    if nc:
        nct = p.nc_total
        rho = rng.uniform(0.5, 2.0, (p.batch, nct))
        kw = dict(ys=T(rng.standard_normal((p.batch, nct))), zs=T(rng.standard_normal((p.batch, nct))), rho=T(rho),
                  inv_rho=T(1.0 / rho))
    x0 = T(p.x0)
    outs = []
    for graph in (False, True):
        sol = P.LQRCudaSolver.from_problem(p, num_segments=S)
        sol.set_stream(side.cuda_stream)
        out = torch.zeros_like(wprev)
        with torch.cuda.stream(side):
            for rep in range(3):
                if graph:
                    sol.solve_device(wprev, x0, out, sigma=1e-3, **kw)
                else:
                    sol.update_problem_data_device(wprev, kw.get("ys"), kw.get("zs"), kw.get("inv_rho"), sigma=1e-3)
                    sol.backward_device(kw.get("rho"))
                    sol.forward_device(x0, out)
            if graph:   # other output pointer -> the graph is captured again
                out2 = torch.zeros_like(wprev)
                sol.solve_device(wprev, x0, out2, sigma=1e-3, **kw)
                sol.synchronize()
                assert torch.equal(out, out2)
        sol.synchronize()
        assert sol.last_status()[0] == 0
        outs.append(out.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    for b in range(p.batch):
        o = oracle.OracleSolver(p, b=b)
        if nc:
            o.update_problem_data(wprev.cpu().numpy()[b], kw["ys"].cpu().numpy()[b], kw["zs"].cpu().numpy()[b],
                                  kw["inv_rho"].cpu().numpy()[b], 1e-3)
            o.backward(kw["rho"].cpu().numpy()[b])
            ref = o.forward(p.x0[b], np.zeros(p.ws_len))
        else:
            ref = o.solve(ws_in=wprev.cpu().numpy()[b], sigma=1e-3)
        assert rel_err(outs[1][b], ref) < TOL


def test_not_positive_definite_reported_per_problem_across_segments():
    """Status slots are per (problem, segment) now (no clearing launch in the solve chain): a non-positive pivot in any
    segment of a problem is reported for that problem only, and a later clean solve reports none."""
    p = P.problems.random_lq(6, 3, 30, batch=4, seed=5)
    bad = p.H.copy()
    Hk = bad[2, 17].reshape(9, 9, order="F")
    Hk[:3, :3] = -5.0 * np.eye(3)                     # indefinite R at stage 17 of problem 2
    bad[2, 17] = Hk.reshape(-1, order="F")
    q = P.problems.Problem(p.nx, p.nu, p.N, p.batch, p.E, p.c, bad, p.h, p.HN, p.hN, p.x0)
    sol = P.LQRCudaSolver.from_problem(q, num_segments=5, load_balancing=False)
    sol.solve(q.zeros_ws(), q.x0, q.zeros_ws())
    n, st = sol.last_status()
    assert n == 1 and st[2] == 18 and st[0] == st[1] == st[3] == 0
    sol.set_model(p)
    sol.solve(p.zeros_ws(), p.x0, p.zeros_ws())
    assert sol.last_status()[0] == 0


# ---------------------------------------------------------------------------------- single-process sharded solve (C ABI)
@pytest.mark.parametrize("G,nc", [(1, 0), (2, 0), (2, 6), (4, 0)])
def test_sharded_c_abi_matches_oracle(oracle, G, nc):
    """pdplqr_sharded_* : one process drives G devices (time slices, NCCL all-gather, redundant interface solve) -- host
    arrays in, host arrays out; with and without constraint rows."""
    import ctypes as C
    import torch
    if torch.cuda.device_count() < G:
        pytest.skip(f"needs {G} GPUs")
    lib = P.capi.load()
    p = P.problems.quadrotor_ltv(3000) if nc == 0 else P.problems.random_lq(6, 3, 90, batch=1, seed=8, nc=nc)
    rng = np.random.default_rng(4)
    wprev = 0.1 * rng.standard_normal((1, p.ws_len))
    hs = C.c_void_p()
    ncs = None if p.ncs is None else np.ascontiguousarray(p.ncs, dtype=np.int32)
    rc = lib.pdplqr_sharded_create(C.byref(hs), p.nx, p.nu, p.N, None if ncs is None else ncs.ctypes.data_as(C.POINTER(C.c_int)),
                                   G, None, 0 if nc == 0 else 3, P.CHOLESKY)
    assert rc == 0, lib.pdplqr_last_error(None)
    dp = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64).ctypes.data
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (p.E, p.c, p.H, p.h, p.HN, p.hN)]
    D = None if p.D is None else np.ascontiguousarray(p.D, dtype=np.float64)
    assert lib.pdplqr_sharded_set_model(hs, *[a.ctypes.data for a in keep], None if D is None else D.ctypes.data) == 0, \
        lib.pdplqr_sharded_last_error(hs)
    out = np.zeros_like(wprev)
    kw = [None] * 4
    if nc:
        nct = p.nc_total
        ys, zs = rng.standard_normal((1, nct)), rng.standard_normal((1, nct))
        rho = rng.uniform(0.5, 2.0, (1, nct))
        inv = np.ascontiguousarray(1.0 / rho)
        kw = [ys, zs, rho, inv]
    for _ in range(2):   # repeated solves on the same handle
        rc = lib.pdplqr_sharded_solve(hs, wprev.ctypes.data, *[dp(a) for a in kw], 1e-4, np.ascontiguousarray(p.x0).ctypes.data,
                                      out.ctypes.data)
        assert rc == 0, lib.pdplqr_sharded_last_error(hs)
    assert lib.pdplqr_sharded_num_devices(hs) == G
    lib.pdplqr_sharded_destroy(hs)
    o = oracle.OracleSolver(p)
    if nc:
        o.update_problem_data(wprev[0], kw[0][0], kw[1][0], kw[3][0], 1e-4)
        o.backward(kw[2][0])
        ref = o.forward(p.x0[0], np.zeros(p.ws_len))
    else:
        ref = o.solve(ws_in=wprev[0], sigma=1e-4)
    assert rel_err(out[0], ref) < TOL


@pytest.mark.gpu
def test_wave_size_is_whole_sms_and_default_segmentation_uses_it():
    """pdplqr_wave_size: (problem, segment) groups resident at once in the throughput-mode stage sweep = SMs x CTAs per SM of
    the kernel that would run; num_segments = 0 on a long horizon picks whole waves."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for nx, nu in [(12, 4), (30, 10), (6, 3)]:
        w = P.wave_size(nx, nu)
        assert w > 0 and w % sms == 0, (nx, nu, w)
    p = P.problems.quadrotor_ltv(1 << 16)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=0, load_balancing=2)
    assert sol.num_segments % P.wave_size(12, 4) == 0 or sol.num_segments < P.wave_size(12, 4)


# ---------------------------------------------------------------------------------- memcheck substitute (guard bands)
def test_debug_guard_bands_detect_and_stay_clean(oracle, monkeypatch):
    """PDPLQR_DEBUG_GUARDS=1: allocations are NaN-filled and sit between guard bands.  A full protocol run (segments, tree,
    constraints, no-refactor, costates) still matches the oracle (nothing reads memory the library did not write), leaves
    every band intact, and the detector reports a deliberate 3-byte overrun (self-test).  The whole GPU suite is also run
    once in this mode (profiles/r2_pytest_gpu_guards.log)."""
    import ctypes as C
    p = P.problems.random_lq(6, 3, 40, batch=3, seed=5, nc=5)
    plain = P.LQRCudaSolver.from_problem(p, num_segments=4)
    assert plain.debug_check_guards() == -1 or P.solver.GUARD_STATS["enabled"]
    monkeypatch.setenv("PDPLQR_DEBUG_GUARDS", "1")
    sol = P.LQRCudaSolver.from_problem(p, num_segments=4)
    sol.set_option(P.capi.OPT_AFFINE_CACHE, 1)
    rng = np.random.default_rng(3)
    nct = p.nc_total
    rho = np.full((p.batch, nct), 0.5)
    for it in range(2):
        ws_prev = rng.standard_normal((p.batch, p.ws_len))
        ys, zs = rng.standard_normal((p.batch, nct)), rng.standard_normal((p.batch, nct))
        sol.update_problem_data(ws_prev, ys, zs, 1.0 / rho, sigma=1e-3)
        if it == 0:
            sol.backward(rho)
        else:
            sol.backward_without_factorization(rho)
        out = sol.forward(p.x0, np.zeros((p.batch, p.ws_len)))
        lam = sol.costates(out)
        assert np.all(np.isfinite(out)) and np.all(np.isfinite(lam))
        for b in range(p.batch):
            ref = oracle.OracleSolver(p, b=b).solve(ws_in=ws_prev[b], sigma=1e-3, ys=ys[b], zs=zs[b], rho=rho[b],
                                                    inv_rho=1.0 / rho[b])
            assert rel_err(out[b], ref) < 1e-9
    assert sol.debug_check_guards() == 0
    n = C.c_longlong(-12345)
    assert sol._lib.pdplqr_debug_check_guards(sol._h, C.byref(n)) == 0 and n.value == 3
    if P.solver.GUARD_STATS["enabled"]:   # the suite-wide fixture would (rightly) flag the deliberate overrun
        sol._lib.pdplqr_destroy(sol._h)
        sol._h = None


@pytest.mark.parametrize("family", ["quadrotor-box", "conic-soc"])
def test_admm_rho_adaptation_matches_the_numpy_restatement(oracle, family):
    """The rho-adaptation policy pinned numerically, not only behaviourally: the device loop with
    OSQP's rule on (convergence test and rescale decision on the device, re-factorisation per rescale) takes the same
    number of iterations and rescales as oracle/admm_ref.py::admm_adaptive and ends at the same iterate -- on the example's
    box-constrained quadrotor (segment path, S = 2) and on the C4 family (nx30 / nu10, box + second-order cones, selection-matrix
    constraint rows) at a short horizon."""
    from oracle import admm_ref
    if family == "quadrotor-box":
        p = P.problems.quadrotor_example(N=20, constrained=True)
        p.x0[0, 2] = 0.0
        S, rho0, kw = 2, 1e-3, dict(max_iter=4000, eps_abs=1e-5, eps_rel=1e-5, check_every=25)
    else:
        p = P.problems.random_conic_batch(batch=1, N=12, seed=5)
        S, rho0, kw = 1, 0.1, dict(max_iter=3000, eps_abs=1e-6, eps_rel=1e-6, check_every=20)
    lb = np.where(np.isfinite(p.e_lb), p.e_lb, -1e20)
    ub = np.where(np.isfinite(p.e_ub), p.e_ub, 1e20)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=S)
    sol.admm_set_cones(p.cones, lb, ub)
    sol.admm_configure(use_graph=True, adaptive_rho=True, rho_tau=5.0, max_rho_updates=6)
    rho = np.full((1, p.nc_total), rho0)
    ws, zs, ys = p.zeros_ws(), np.zeros((1, p.nc_total)), np.zeros((1, p.nc_total))
    iters, r = sol.admm_solve(p.x0, ws, zs, ys, rho, sigma=1e-6, alpha=1.6, **kw)
    _, n_upd = sol.admm_stats()
    w, z, y, it_ref, n_ref, res_ref, _ = admm_ref.admm_adaptive(p, 0, rho[0], sigma=1e-6, alpha=1.6, rho_tau=5.0,
                                                                max_rho_updates=6, **kw)
    assert n_ref >= 1 and it_ref < kw["max_iter"]
    assert (iters, n_upd) == (it_ref, n_ref), ((iters, n_upd, list(r)), (it_ref, n_ref, res_ref))
    assert rel_err(ws[0], w) < 1e-8 and rel_err(zs[0], z) < 1e-8 and rel_err(ys[0], y) < 1e-7
    assert abs(r[0] - res_ref[0]) < 1e-6 * max(1.0, res_ref[0]) and abs(r[1] - res_ref[1]) < 1e-6 * max(1.0, res_ref[1])


@pytest.mark.parametrize("family", ["quadrotor-box", "conic-soc"])
def test_admm_solution_satisfies_the_conic_kkt_conditions(oracle, family):
    """Row a11 is 'parity unpinned' by the reference (the outer iteration is not in it): besides the comparison with the numpy
    restatement, the CUDA result itself is checked against the optimality conditions of the conic problem, with no Riccati oracle
    and no ADMM restatement in the loop (tests/kkt_ref.py::conic_kkt_violations: constraint link, dynamics, cone membership,
    multiplier in the normal cone, stationarity through the independent sparse KKT solve)."""
    from kkt_ref import conic_kkt_violations
    if family == "quadrotor-box":
        p, S, rho0, iters = P.problems.quadrotor_example(N=20, constrained=True), 2, 1.0, 1500
    else:
        p, S, rho0, iters = P.problems.random_conic_batch(batch=2, N=12, seed=5), 1, 10.0, 3000
    lb = np.where(np.isfinite(p.e_lb), p.e_lb, -1e20)
    ub = np.where(np.isfinite(p.e_ub), p.e_ub, 1e20)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=S)
    sol.admm_set_cones(p.cones, lb, ub)
    rho = np.full((p.batch, p.nc_total), rho0)
    ws, zs, ys = p.zeros_ws(), np.zeros((p.batch, p.nc_total)), np.zeros((p.batch, p.nc_total))
    it, r = sol.admm_solve(p.x0, ws, zs, ys, rho, sigma=1e-6, alpha=1.6, max_iter=iters, eps_abs=0.0, eps_rel=0.0,
                           check_every=iters)
    assert it == iters and r[0] < 1e-9 and r[1] < 1e-8
    for b in range(p.batch):
        v = conic_kkt_violations(p, b, ws[b], zs[b], ys[b], rho[b])
        assert v["link"] < 1e-9 and v["dynamics"] < 1e-10 and v["cone"] < 1e-9 and v["normal_cone"] < 1e-7 \
            and v["stationarity"] < 1e-8, (b, v)
