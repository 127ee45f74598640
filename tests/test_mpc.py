"""Receding-horizon driver (pdplqr_b200.mpc): warm-start shift on the CPU, closed loop
on the GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pdplqr_b200 as P  # noqa: E402
from pdplqr_b200.mpc import RecedingHorizon, shift_warm_start  # noqa: E402


def test_shift_warm_start_moves_every_stage_one_period_ahead():
    nx, nu, N, batch = 3, 2, 5, 2
    s = nx + nu
    rng = np.random.default_rng(0)
    ws = rng.standard_normal((batch, N * s + nx))
    ncs = np.array([nu, s, s, s, s, nx])            # the example's pattern: short first block, terminal block of nx rows
    nct = int(ncs.sum())
    zs, ys = rng.standard_normal((batch, nct)), rng.standard_normal((batch, nct))
    ws0, zs0, ys0 = ws.copy(), zs.copy(), ys.copy()
    w2, z2, y2 = shift_warm_start(ws, zs, ys, nx, nu, ncs)
    assert np.array_equal(ws, ws0) and np.array_equal(zs, zs0) and np.array_equal(ys, ys0)   # inputs untouched
    W0 = ws0[:, : N * s].reshape(batch, N, s)
    W2 = w2[:, : N * s].reshape(batch, N, s)
    assert np.array_equal(W2[:, : N - 1], W0[:, 1:])                      # stage k <- stage k+1
    assert np.array_equal(W2[:, N - 1, :nu], W0[:, N - 1, :nu])           # last control repeated
    assert np.array_equal(W2[:, N - 1, nu:], ws0[:, N * s:])              # from the old terminal state
    assert np.array_equal(w2[:, N * s:], ws0[:, N * s:])
    off = np.concatenate([[0], np.cumsum(ncs)])
    assert np.array_equal(z2[:, off[0]: off[1]], zs0[:, off[0]: off[1]])  # block sizes differ (nu vs s): kept
    for k in (1, 2, 3):
        assert np.array_equal(z2[:, off[k]: off[k + 1]], zs0[:, off[k + 1]: off[k + 2]])
        assert np.array_equal(y2[:, off[k]: off[k + 1]], ys0[:, off[k + 1]: off[k + 2]])
    assert np.array_equal(z2[:, off[4]:], zs0[:, off[4]:])                # last running stage and terminal stage kept


def test_shift_warm_start_degenerate_horizons():
    for N in (1, 2):
        nx, nu = 2, 1
        ws = np.arange(1.0, N * 3 + 2 + 1).reshape(1, -1)
        w2, z2, y2 = shift_warm_start(ws, np.zeros((1, 0)), np.zeros((1, 0)), nx, nu, np.zeros(N + 1, int))
        assert w2.shape == ws.shape and z2.shape == (1, 0)
        assert np.array_equal(w2[0, (N - 1) * 3 + nu: N * 3], ws[0, N * 3:])


@pytest.mark.gpu
def test_receding_horizon_quadrotor_tracks_and_warm_start_pays():
    """Box-constrained quadrotor (the example with its constraints switched on): closed loop from rest towards
    z = 1 m.  The applied controls respect the input box up to the ADMM tolerance, the height converges, and the
    shifted warm start needs fewer ADMM iterations than cold starts after the first period."""
    def run(warm):
        p = P.problems.quadrotor_example(N=20, constrained=True)
        p.x0[0, 2] = 0.0
        sol = P.LQRCudaSolver.from_problem(p, num_segments=2)
        rh = RecedingHorizon(sol, p, rho=0.1, max_iter=400, eps_abs=1e-4, eps_rel=1e-4, check_every=10, warm_start=warm)
        us = []
        for _ in range(25):
            u, info = rh.step()
            us.append(u[0].copy())
        return rh, np.array(us)
    rh_w, u_w = run(True)
    rh_c, u_c = run(False)
    assert np.max(np.abs(u_w - np.clip(u_w, -0.9916, 2.4084))) < 2e-2
    assert abs(rh_w.x[0, 2] - 1.0) < abs(0.0 - 1.0) * 0.5                  # moved at least half way in 25 periods
    assert np.max(np.abs(u_w - u_c)) < 5e-2                                # same closed loop up to the tolerance
    it_w = sum(h["iterations"] for h in rh_w.history[1:])
    it_c = sum(h["iterations"] for h in rh_c.history[1:])
    assert it_w < it_c
