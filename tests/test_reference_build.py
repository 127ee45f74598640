"""Pins the oracle port against the REFERENCE ITSELF: oracle/_ref/libpdpref.so is the reference's own headers
(lqr_solver.hpp, lqr_solver_parallel.hpp, lqr_kernel*.hpp, condensed_system.hpp), compiled unmodified from
/root/reference against a minimal Eigen-API shim (oracle/eigen_shim; Eigen3 is absent from this image).  The .so is
built in the build container only (oracle/Makefile target `ref`) and travels to the GPU box; tests skip when it is
missing.  GPU-vs-reference comparisons live in the gpu-marked test at the bottom."""
import os

import numpy as np
import pytest

import pdplqr_b200 as P
from conftest import rel_err
from oracle import reflib

pytestmark = pytest.mark.skipif(not reflib.available(), reason="oracle/_ref/libpdpref.so not built (needs /root/reference)")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_reproduces_config1_golden(oracle):
    """examples/lqr_example.cpp as shipped: LQRSolver and LQRParallelSolver(4, true, CHOLESKY)."""
    g = np.load(os.path.join(GOLD, "c1_quadrotor.npz"))
    p = P.problems.quadrotor_example()
    seq = reflib.ReferenceSolver(p).solve()
    par = reflib.ReferenceSolver(p, parallel=True, num_segments=4, load_balancing=True, cholesky=True).solve()
    assert rel_err(seq, g["ws_seq"]) < 1e-12
    assert rel_err(par, g["ws_seq"]) < 1e-12
    assert rel_err(seq, g["ws_kkt"]) < 1e-12
    assert rel_err(seq, oracle.OracleSolver(p).solve()) < 1e-13


@pytest.mark.parametrize("S,lb,chol", [(2, True, True), (4, False, False), (8, True, False), (8, False, True)])
def test_oracle_port_equals_reference_parallel(oracle, S, lb, chol):
    p = P.problems.quadrotor_ltv(96)
    rng = np.random.default_rng(S)
    wprev = 0.1 * rng.standard_normal(p.ws_len)
    ref = reflib.ReferenceSolver(p, parallel=True, num_segments=S, load_balancing=lb, cholesky=chol).solve(wprev, 1e-3)
    o = oracle.OracleSolver(p, parallel=True, num_segments=S, load_balancing=lb, condensed=1 if chol else 0)
    assert rel_err(o.solve(ws_in=wprev, sigma=1e-3), ref) < 1e-11


@pytest.mark.parametrize("parallel,S", [(False, 1), (True, 3)])
def test_oracle_port_equals_reference_constrained_and_nofact(oracle, parallel, S):
    """constraint fold-in (lqr_kernel.hpp:106-112) and backward_without_factorization (:149-178) on both sides."""
    q = P.problems.random_lq(6, 3, 18, batch=1, seed=21, nc=5)
    rng = np.random.default_rng(2)
    nct = q.nc_total
    rho = rng.uniform(0.1, 2.0, nct)
    inv_rho = 1.0 / rho
    r = reflib.ReferenceSolver(q, parallel=parallel, num_segments=S, cholesky=False)
    o = oracle.OracleSolver(q, parallel=parallel, num_segments=S, condensed=0)
    for it in range(3):
        w, y, z = rng.standard_normal(q.ws_len), rng.standard_normal(nct), rng.standard_normal(nct)
        r.update_problem_data(w, y, z, inv_rho, 1e-2)
        o.update_problem_data(w, y, z, inv_rho, 1e-2)
        r.backward(rho, factorize=(it == 0))
        if it == 0:
            o.backward(rho)
        else:
            o.backward_without_factorization(rho)
        a = r.forward(q.x0[0])
        b = o.forward(q.x0[0], np.zeros(q.ws_len))
        assert rel_err(b, a) < 1e-11, it


def test_reference_matches_random_golden():
    g = np.load(os.path.join(GOLD, "random_6_3_40.npz"))
    q = P.problems.random_lq(6, 3, 40, batch=1, seed=11)
    ws = reflib.ReferenceSolver(q).solve(g["wprev"], 0.05)
    assert rel_err(ws, g["ws_seq"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("S", [1, 4])
def test_gpu_equals_reference_binary(S):
    """The CUDA path against the compiled reference directly (not via the port)."""
    p = P.problems.quadrotor_example()
    ref = reflib.ReferenceSolver(p, parallel=S > 1, num_segments=S).solve()
    sol = P.LQRCudaSolver.from_problem(p, num_segments=S)
    ws = sol.solve(p.zeros_ws(), p.x0, p.zeros_ws())
    assert rel_err(ws[0], ref) < 1e-9
    q = P.problems.random_lq(12, 4, 30, batch=1, seed=77, nc=6)
    rng = np.random.default_rng(0)
    nct = q.nc_total
    w, y, z = rng.standard_normal((1, q.ws_len)), rng.standard_normal((1, nct)), rng.standard_normal((1, nct))
    rho = rng.uniform(0.1, 1.0, (1, nct))
    inv = np.ascontiguousarray(1.0 / rho)
    r = reflib.ReferenceSolver(q, parallel=S > 1, num_segments=S, cholesky=False)
    ref2 = r.solve(w[0], 1e-3, y[0], z[0], rho[0], inv[0])
    sol2 = P.LQRCudaSolver.from_problem(q, num_segments=S)
    ws2 = sol2.solve(w, q.x0, np.zeros_like(w), sigma=1e-3, ys=y, zs=z, rho=rho, inv_rho=inv)
    assert rel_err(ws2[0], ref2) < 1e-9


REF_EXAMPLE = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "lqr_example_cuda")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_EXAMPLE), reason="oracle/_ref/lqr_example_cuda not built (needs /root/reference)")
def test_reference_example_with_cuda_solver():
    """The reference's OWN examples/lqr_example.cpp with the 2-line integration change of INTEGRATION.md section 1
    (oracle/Makefile target ref_example): its model code fills the reference's Node / LQRModel, its LQRSolver block
    runs on the host, and the LQRParallelSolver block runs on LQRCudaSolver.  The two printed solutions must agree."""
    import re
    import subprocess
    out = subprocess.run([REF_EXAMPLE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()

    def rows(tag):
        inputs = [np.array(l.split(":", 1)[1].split(), dtype=float) for l in lines if re.match(r"Input \d+ \(%s\)" % tag, l)]
        k = next(i for i, l in enumerate(lines) if l.startswith("Final state (%s)" % tag))
        return np.concatenate(inputs + [np.array(lines[k + 1].split(), dtype=float)])

    cpu, gpu = rows("LQRSolver"), rows("LQRCudaSolver")   # (the patch also renames the printed tag)
    assert cpu.size == gpu.size == 5 * 4 + 12
    assert abs(cpu[0] - (-2.898056669662)) < 1e-9          # SURVEY.md's independent probe value of u_0[0]
    assert rel_err(gpu, cpu) < 1e-9
