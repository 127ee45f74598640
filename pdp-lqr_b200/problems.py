"""Synthetic LQ problems for the five BASELINE.json configs, in the flat FP64 column-major layout of the
C ABI (include/pdplqr.h).  All generators are new code (the reference ships exactly one problem, the
quadrotor example at /root/reference examples/lqr_example.cpp:53-171, restated in `quadrotor_example`).

Flat layout of a batch of `batch` problems (matches lqr::Node, lqr_model.hpp:8-64, column-major):
    E  [batch, N, nx*s]   E_k = [B_k A_k]            c  [batch, N, nx]
    H  [batch, N, s*s]    H_k = [R S; S^T Q]         h  [batch, N, s]      (h_k = [r; q])
    HN [batch, nx*nx]     hN [batch, nx]             x0 [batch, nx]
    D  [batch, d_total]   concat_k (nc_k x dim_k), col-major (None when there are no constraints)
    e_lb / e_ub [batch, nc_total]                    cone descriptors: see `Problem.cones`
    ws [batch, N*s + nx]  w_k = [u_k; x_k], w_N = x_N
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Problem:
    nx: int
    nu: int
    N: int
    batch: int
    E: np.ndarray
    c: np.ndarray
    H: np.ndarray
    h: np.ndarray
    HN: np.ndarray
    hN: np.ndarray
    x0: np.ndarray
    ncs: np.ndarray | None = None  # [N+1] int32, shared by the batch
    D: np.ndarray | None = None
    e_lb: np.ndarray | None = None
    e_ub: np.ndarray | None = None
    # cones tiling every stage's constraint rows: list of (stage k, first row, dim, type) sorted by stage;
    # type 0 = box [e_lb, e_ub], 1 = second-order cone (first row t >= ||rest||), 2 = ball (radius e_ub[first row])
    cones: list = field(default_factory=list)
    name: str = ""

    @property
    def s(self) -> int:
        return self.nx + self.nu

    @property
    def ws_len(self) -> int:
        return self.N * self.s + self.nx

    @property
    def nc_total(self) -> int:
        return 0 if self.ncs is None else int(np.sum(self.ncs))

    def coff(self) -> np.ndarray:
        ncs = np.zeros(self.N + 1, np.int64) if self.ncs is None else self.ncs.astype(np.int64)
        return np.concatenate([[0], np.cumsum(ncs)])

    def doff(self) -> np.ndarray:
        ncs = np.zeros(self.N + 1, np.int64) if self.ncs is None else self.ncs.astype(np.int64)
        dims = np.full(self.N + 1, self.s, np.int64)
        dims[-1] = self.nx
        return np.concatenate([[0], np.cumsum(ncs * dims)])

    def zeros_ws(self) -> np.ndarray:
        return np.zeros((self.batch, self.ws_len))

    def select(self, idx) -> "Problem":
        """Sub-batch view (used to bound CPU-baseline samples)."""
        idx = np.atleast_1d(np.arange(self.batch)[idx])
        opt = lambda a: None if a is None else np.ascontiguousarray(a[idx])
        return Problem(self.nx, self.nu, self.N, len(idx), opt(self.E), opt(self.c), opt(self.H), opt(self.h),
                       opt(self.HN), opt(self.hN), opt(self.x0), self.ncs, opt(self.D), opt(self.e_lb),
                       opt(self.e_ub), self.cones, self.name)


def _cm(M: np.ndarray) -> np.ndarray:
    """Column-major flatten."""
    return np.asarray(M, dtype=np.float64).flatten(order="F")


# ------------------------------------------------------------------------------------------------
# C1: the shipped quadrotor example (examples/lqr_example.cpp:53-171), nx=12, nu=4, N=100, nc=0.
_QUAD_A = np.array([
    [1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0., 0., 0.],
    [0., 1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0., 0.],
    [0., 0., 1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0.],
    [0.0488, 0., 0., 1., 0., 0., 0.0016, 0., 0., 0.0992, 0., 0.],
    [0., -0.0488, 0., 0., 1., 0., 0., -0.0016, 0., 0., 0.0992, 0.],
    [0., 0., 0., 0., 0., 1., 0., 0., 0., 0., 0., 0.0992],
    [0., 0., 0., 0., 0., 0., 1., 0., 0., 0., 0., 0.],
    [0., 0., 0., 0., 0., 0., 0., 1., 0., 0., 0., 0.],
    [0., 0., 0., 0., 0., 0., 0., 0., 1., 0., 0., 0.],
    [0.9734, 0., 0., 0., 0., 0., 0.0488, 0., 0., 0.9846, 0., 0.],
    [0., -0.9734, 0., 0., 0., 0., 0., -0.0488, 0., 0., 0.9846, 0.],
    [0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.9846]])
_QUAD_B = np.array([
    [0., -0.0726, 0., 0.0726],
    [-0.0726, 0., 0.0726, 0.],
    [-0.0152, 0.0152, -0.0152, 0.0152],
    [-0., -0.0006, -0., 0.0006],
    [0.0006, 0., -0.0006, 0.0000],
    [0.0106, 0.0106, 0.0106, 0.0106],
    [0., -1.4512, 0., 1.4512],
    [-1.4512, 0., 1.4512, 0.],
    [-0.3049, 0.3049, -0.3049, 0.3049],
    [-0., -0.0236, 0., 0.0236],
    [0.0236, 0., -0.0236, 0.],
    [0.2107, 0.2107, 0.2107, 0.2107]])
_QUAD_QDIAG = np.array([0., 0., 10., 10., 10., 10., 0., 0., 0., 5., 5., 5.])
_QUAD_RDIAG = np.array([0.1, 0.1, 0.1, 0.1])
_QUAD_XREF = np.array([0., 0., 1., 0., 0., 0., 0., 0., 0., 0., 0., 0.])
QUAD_U_MIN = np.array([-0.9916] * 4)
QUAD_U_MAX = np.array([2.4084] * 4)
QUAD_X_MIN = np.array([-0.52359878, -0.52359878, -np.inf, -np.inf, -np.inf, -1.] + [-np.inf] * 6)
QUAD_X_MAX = np.array([0.52359878, 0.52359878, np.inf, np.inf, np.inf, np.inf, np.inf, np.inf, 2.5] + [np.inf] * 3)


def quadrotor_example(N: int = 100, constrained: bool = False) -> Problem:
    """Config 1.  `constrained=True` enables the box constraints the example disables with `nc = 0;`
    (lqr_example.cpp:126-127,157-158): k=0: nu rows on u; 0<k<N: nu+nx rows (identity D); k=N: nx rows."""
    nx, nu = 12, 4
    s = nx + nu
    Q = np.diag(_QUAD_QDIAG)
    R = np.diag(_QUAD_RDIAG)
    q = -_QUAD_XREF @ Q
    Ek = np.hstack([_QUAD_B, _QUAD_A])
    Hk = np.zeros((s, s))
    Hk[:nu, :nu] = R
    Hk[nu:, nu:] = Q
    hk = np.concatenate([np.zeros(nu), q])
    E = np.tile(_cm(Ek), (1, N, 1))
    c = np.zeros((1, N, nx))
    H = np.tile(_cm(Hk), (1, N, 1))
    h = np.tile(hk, (1, N, 1))
    p = Problem(nx, nu, N, 1, E, c, H, h, _cm(Q)[None], q[None].copy(), np.zeros((1, nx)), name="C1-quadrotor")
    if constrained:
        ncs = np.full(N + 1, s, np.int32)
        ncs[0] = nu
        ncs[N] = nx
        Ds, lbs, ubs = [], [], []
        for k in range(N + 1):
            if k == 0:
                Dk = np.zeros((nu, s)); Dk[:, :nu] = np.eye(nu)
                lb, ub = QUAD_U_MIN, QUAD_U_MAX
            elif k < N:
                Dk = np.eye(s)
                lb, ub = np.concatenate([QUAD_U_MIN, QUAD_X_MIN]), np.concatenate([QUAD_U_MAX, QUAD_X_MAX])
            else:
                Dk = np.eye(nx)
                lb, ub = QUAD_X_MIN, QUAD_X_MAX
            Ds.append(_cm(Dk)); lbs.append(lb); ubs.append(ub)
        p.ncs = ncs
        p.D = np.concatenate(Ds)[None]
        p.e_lb = np.concatenate(lbs)[None]
        p.e_ub = np.concatenate(ubs)[None]
        p.cones = [(k, 0, int(ncs[k]), 0) for k in range(N + 1)]
    return p


LTV_CHUNK = 1 << 16   # stages per independently seeded chunk of quadrotor_ltv


def quadrotor_ltv(N: int, seed: int = 20251018, x0_seed: int = 7, start: int = 0, count: int | None = None) -> Problem:
    """Configs 2 / 5: quadrotor replicated to N stages, stored per stage (LTV layout) with a seeded
    perturbation so that stages are distinct:  A_k = A + 1e-3 U(-1,1) o |A|, c_k = 1e-3 N(0,1).
    The perturbations are drawn per chunk of LTV_CHUNK stages (chunk j: seed + j), so that a rank of a horizon-sharded
    run can generate ITS time slice [start, start + count) of the N-stage problem without building the rest
    (returned as an N = count problem with the attributes `start` and `x0_global`; terminal cost only on the last)."""
    nx, nu = 12, 4
    s = nx + nu
    count = N - start if count is None else count
    assert 0 <= start and count >= 1 and start + count <= N
    Q = np.diag(_QUAD_QDIAG)
    R = np.diag(_QUAD_RDIAG)
    q = -_QUAD_XREF @ Q
    Hk = np.zeros((s, s)); Hk[:nu, :nu] = R; Hk[nu:, nu:] = Q
    hk = np.concatenate([np.zeros(nu), q])
    # E_k column-major: columns 0..nu-1 = B, nu.. = A
    Ecm = np.empty((count, s, nx))           # [k, col, row]  == column-major (nx x s)
    Ecm[:, :nu, :] = _QUAD_B.T[None]
    c = np.empty((count, nx))
    absA = np.abs(_QUAD_A)
    for j in range(start // LTV_CHUNK, (start + count - 1) // LTV_CHUNK + 1):
        k0, k1 = j * LTV_CHUNK, min((j + 1) * LTV_CHUNK, N)
        rng = np.random.default_rng(seed + j)
        Ak = _QUAD_A[None] + 1e-3 * rng.uniform(-1, 1, (k1 - k0, nx, nx)) * absA[None]
        ck = 1e-3 * rng.standard_normal((k1 - k0, nx))
        lo, hi = max(k0, start), min(k1, start + count)
        Ecm[lo - start:hi - start, nu:, :] = np.transpose(Ak[lo - k0:hi - k0], (0, 2, 1))
        c[lo - start:hi - start] = ck[lo - k0:hi - k0]
    E = Ecm.reshape(1, count, nx * s)
    H = np.broadcast_to(_cm(Hk), (1, count, s * s)).copy()
    h = np.broadcast_to(hk, (1, count, s)).copy()
    x0 = 0.1 * np.random.default_rng(x0_seed).standard_normal((1, nx))
    whole = start == 0 and count == N
    is_last = start + count == N
    p = Problem(nx, nu, count, 1, E, c[None], H, h, _cm(Q)[None] if is_last else np.zeros((1, nx * nx)),
                q[None].copy() if is_last else np.zeros((1, nx)), x0 if whole else np.zeros((1, nx)),
                name=f"quadrotor-LTV-N{N}" + ("" if whole else f"[{start}:{start + count}]"))
    p.start, p.x0_global = start, x0
    return p


def cartpole_batch(batch: int = 65536, N: int = 128, seed: int = 1234) -> Problem:
    """Config 3: `batch` independent linearised cart-poles (nx=4, nu=1), Euler dt=0.02, per-problem
    parameters m_c in U[0.5,2], m_p in U[0.05,0.5], l in U[0.3,1]; Q = diag(1,1,10,1)(1+0.1U), R = 0.1,
    Q_N = 10 Q, x0 in U[-0.5,0.5]^4.  Stored LTV per problem (every stage has its own copy, plus a small
    seeded per-stage perturbation of A and c so stages are distinct)."""
    nx, nu, dt, g = 4, 1, 0.02, 9.81
    s = nx + nu
    rng = np.random.default_rng(seed)
    mc = rng.uniform(0.5, 2.0, batch)
    mp = rng.uniform(0.05, 0.5, batch)
    ln = rng.uniform(0.3, 1.0, batch)
    # state (pos, vel, theta, omega), upright linearisation
    Ac = np.zeros((batch, nx, nx))
    Ac[:, 0, 1] = 1.0
    Ac[:, 1, 2] = -mp * g / mc
    Ac[:, 2, 3] = 1.0
    Ac[:, 3, 2] = (mc + mp) * g / (mc * ln)
    Bc = np.zeros((batch, nx, nu))
    Bc[:, 1, 0] = 1.0 / mc
    Bc[:, 3, 0] = -1.0 / (mc * ln)
    A = np.eye(nx)[None] + dt * Ac
    B = dt * Bc
    Qd = np.array([1., 1., 10., 1.])[None] * (1.0 + 0.1 * rng.uniform(0, 1, (batch, nx)))
    Ecm = np.empty((batch, N, s, nx))
    Ecm[:, :, :nu, :] = np.transpose(B, (0, 2, 1))[:, None]
    Ak = A[:, None] * (1.0 + 1e-3 * rng.uniform(-1, 1, (batch, N, nx, nx)))
    Ecm[:, :, nu:, :] = np.transpose(Ak, (0, 1, 3, 2))
    E = Ecm.reshape(batch, N, nx * s)
    c = 1e-3 * rng.standard_normal((batch, N, nx))
    Hm = np.zeros((batch, s, s))
    Hm[:, 0, 0] = 0.1
    idx = np.arange(nx)
    Hm[:, nu + idx, nu + idx] = Qd
    H = np.broadcast_to(Hm.reshape(batch, 1, s * s), (batch, N, s * s)).copy()
    h = np.zeros((batch, N, s))
    HNm = np.zeros((batch, nx, nx))
    HNm[:, idx, idx] = 10.0 * Qd
    x0 = rng.uniform(-0.5, 0.5, (batch, nx))
    return Problem(nx, nu, N, batch, E, c, H, h, HNm.reshape(batch, nx * nx), np.zeros((batch, nx)), x0,
                   name=f"C3-cartpole-b{batch}-N{N}")


def random_conic_batch(batch: int = 4096, N: int = 256, nx: int = 30, nu: int = 10, seed: int = 99,
                       soc: bool = True) -> Problem:
    """Config 4: OSQP-benchmark-style random system, A = I + 0.1 N(0,1) rescaled to spectral radius 1,
    B = N(0,1), Q = diag(U[0,10]) with 30% zeros, R = 0.1 I; box on all of [u;x] (identity D) plus one
    SOC of dim 4 on rows (t; u_1..u_3) -> nc = s + 4 for 0<k<N; k=0: nu box rows (+SOC); k=N: nx box rows.
    The system matrices are shared by the batch up to a per-problem scaling; x0 differs per problem."""
    s = nx + nu
    rng = np.random.default_rng(seed)
    A = np.eye(nx) + 0.1 * rng.standard_normal((nx, nx))
    A /= np.max(np.abs(np.linalg.eigvals(A)))
    B = rng.standard_normal((nx, nu))
    qd = rng.uniform(0, 10, nx) * (rng.uniform(0, 1, nx) > 0.3)
    Hk = np.zeros((s, s)); Hk[:nu, :nu] = 0.1 * np.eye(nu); Hk[nu:, nu:] = np.diag(qd)
    scale = 1.0 + 0.05 * rng.uniform(-1, 1, (batch, 1, 1))
    Ecm = np.empty((batch, s, nx))
    Ecm[:, :nu, :] = B.T[None] * scale
    Ecm[:, nu:, :] = A.T[None]
    E = np.broadcast_to(Ecm.reshape(batch, 1, nx * s), (batch, N, nx * s)).copy()
    c = np.zeros((batch, N, nx))
    H = np.broadcast_to(_cm(Hk), (batch, N, s * s)).copy()
    h = np.zeros((batch, N, s))
    HN = np.broadcast_to(_cm(np.diag(qd)), (batch, nx * nx)).copy()
    hN = np.zeros((batch, nx))
    x0 = rng.uniform(-1, 1, (batch, nx))
    nsoc = 4 if soc else 0
    ncs = np.full(N + 1, s + nsoc, np.int32)
    ncs[0] = nu + nsoc
    ncs[N] = nx
    umax, xmax = 1.0 + rng.uniform(0, 1, nu), 2.0 + rng.uniform(0, 3, nx)
    Ds, lbs, ubs, cones = [], [], [], []
    for k in range(N + 1):
        if k == N:
            Dk = np.eye(nx); lb, ub = -xmax, xmax
            cones.append((k, 0, nx, 0))
        else:
            nb = nu if k == 0 else s
            Dk = np.zeros((nb + nsoc, s))
            Dk[:nb, :nb] = np.eye(nb)
            lb = np.concatenate([-umax, -xmax])[:nb]; ub = np.concatenate([umax, xmax])[:nb]
            cones.append((k, 0, nb, 0))
            if nsoc:
                # second-order cone on the first four controls:  || (u_1, u_2, u_3) || <= u_0 + 1  is not conic in w,
                # so the cone is put on the rows (u_0; u_1, u_2, u_3):  || (u_1,u_2,u_3) || <= u_0
                for r in range(4):
                    Dk[nb + r, r] = 1.0
                lb = np.concatenate([lb, np.full(4, -np.inf)]); ub = np.concatenate([ub, np.full(4, np.inf)])
                cones.append((k, nb, 4, 1))
        Ds.append(_cm(Dk)); lbs.append(lb); ubs.append(ub)
    D1 = np.concatenate(Ds)
    p = Problem(nx, nu, N, batch, E, c, H, h, HN, hN, x0, ncs,
                np.broadcast_to(D1, (batch, D1.size)).copy(),
                np.broadcast_to(np.concatenate(lbs), (batch, int(ncs.sum()))).copy(),
                np.broadcast_to(np.concatenate(ubs), (batch, int(ncs.sum()))).copy(),
                cones, name=f"C4-conic-b{batch}-N{N}")
    return p


def random_lq(nx: int, nu: int, N: int, batch: int = 1, seed: int = 0, nc: int = 0, dense_cost: bool = True) -> Problem:
    """Generic well-posed random LTV LQ problem (dense H with cross terms S, nonzero h and c) for parity tests,
    optionally with `nc` random constraint rows per stage (terminal: min(nc, nx))."""
    s = nx + nu
    rng = np.random.default_rng(seed)
    A = np.eye(nx)[None, None] + 0.2 * rng.standard_normal((batch, N, nx, nx)) / np.sqrt(nx)
    B = rng.standard_normal((batch, N, nx, nu)) / np.sqrt(nx)
    Ecm = np.empty((batch, N, s, nx))
    Ecm[:, :, :nu, :] = np.transpose(B, (0, 1, 3, 2))
    Ecm[:, :, nu:, :] = np.transpose(A, (0, 1, 3, 2))
    E = Ecm.reshape(batch, N, nx * s)
    c = 0.1 * rng.standard_normal((batch, N, nx))
    if dense_cost:
        W = rng.standard_normal((batch, N, s, s)) / np.sqrt(s)
        Hm = W @ np.transpose(W, (0, 1, 3, 2)) + 0.1 * np.eye(s)[None, None]
    else:
        Hm = np.zeros((batch, N, s, s))
        Hm[..., np.arange(s), np.arange(s)] = rng.uniform(0.1, 2.0, (batch, N, s))
    H = Hm.reshape(batch, N, s * s)  # symmetric: row/col-major agree
    h = rng.standard_normal((batch, N, s))
    Wn = rng.standard_normal((batch, nx, nx)) / np.sqrt(nx)
    HN = (Wn @ np.transpose(Wn, (0, 2, 1)) + 0.1 * np.eye(nx)[None]).reshape(batch, nx * nx)
    hN = rng.standard_normal((batch, nx))
    x0 = rng.standard_normal((batch, nx))
    p = Problem(nx, nu, N, batch, E, c, H, h, HN, hN, x0, name=f"random-nx{nx}-nu{nu}-N{N}-b{batch}")
    if nc > 0:
        ncs = np.full(N + 1, nc, np.int32)
        ncs[N] = min(nc, nx)
        doff = np.concatenate([[0], np.cumsum(ncs.astype(np.int64) * np.array([s] * N + [nx]))])
        p.ncs = ncs
        p.D = rng.standard_normal((batch, int(doff[-1]))) / np.sqrt(s)
        nct = int(ncs.sum())
        p.e_lb = -rng.uniform(0.5, 1.5, (batch, nct))
        p.e_ub = rng.uniform(0.5, 1.5, (batch, nct))
        p.cones = [(k, 0, int(ncs[k]), 0) for k in range(N + 1)]
    return p
