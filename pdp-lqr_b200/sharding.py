"""Multi-GPU partitioning of the PDP-LQR hot path, one process per GPU (torch.distributed for the plumbing).

Two shardings, each only where the path shards naturally:
  * batch   -- independent problems: contiguous batch slices per rank, NO data-path collective.
  * horizon -- one very long problem: contiguous time slices per rank.  Each rank reduces its slice to ONE segment
               summary (P | F | C | p | f, 3 nx^2 + 2 nx doubles = 3,648 B at nx = 12), one all_gather of that
               summary per solve (NCCL over NVLink on GPUs, gloo in the CPU tests), every rank then solves the tiny
               G-slice interface system redundantly (cheaper than a second collective) and rolls out its own slice.
The reference's analogue is threads <-> segments with the serial condensed solve on the master thread
(lqr_solver_parallel.hpp:144-145,215); composition of summaries is associative, which is what makes
the hierarchy segments -> GPU -> box legitimate.
"""
from __future__ import annotations

import numpy as np


def batch_slices(batch: int, world: int):
    """Contiguous, balanced batch slices [(start, count)] for `world` ranks."""
    base, rem = divmod(batch, world)
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((s, n))
        s += n
    return out


def horizon_slices(N: int, world: int):
    """Contiguous time slices [(first_stage, n_stages)]; the last rank also owns the terminal stage N."""
    if world > N:
        raise ValueError("more ranks than stages")
    return batch_slices(N, world)


def slice_problem(prob, start: int, count: int, is_last: bool):
    """Time slice [start, start+count) of a single long-horizon Problem as its own Problem (constraints not sliced)."""
    from .problems import Problem
    assert prob.ncs is None, "horizon slicing of constrained problems is not implemented"
    sl = slice(start, start + count)
    nx = prob.nx
    HN = prob.HN if is_last else np.zeros_like(prob.HN)
    hN = prob.hN if is_last else np.zeros_like(prob.hN)
    return Problem(prob.nx, prob.nu, count, prob.batch, np.ascontiguousarray(prob.E[:, sl]),
                   np.ascontiguousarray(prob.c[:, sl]), np.ascontiguousarray(prob.H[:, sl]),
                   np.ascontiguousarray(prob.h[:, sl]), HN, hN, np.zeros((prob.batch, nx)),
                   name=f"{prob.name}[{start}:{start + count}]")


def couple_numpy(summaries, x0):
    """Host reference of the G-slice interface solve (same algebra as the device coupler; used by the CPU/gloo tests
    of the sharding plumbing).  summaries: [G, 3 nx^2 + 2 nx] rows (P | F | C | p | f), column-major blocks.
    Returns xhat [G, nx] (entry states) and lam [G, nx] (exit costates; the last one is zero).
    Serial recursion of condensed_system.hpp:82-138 (LU form)."""
    G = summaries.shape[0]
    nx = x0.shape[0]
    n2 = nx * nx
    P = [summaries[i, 0:n2].reshape(nx, nx, order="F").copy() for i in range(G)]
    F = [summaries[i, n2:2 * n2].reshape(nx, nx, order="F") for i in range(G)]
    Cm = [summaries[i, 2 * n2:3 * n2].reshape(nx, nx, order="F") for i in range(G)]
    p = [summaries[i, 3 * n2:3 * n2 + nx].copy() for i in range(G)]
    f = [summaries[i, 3 * n2 + nx:3 * n2 + 2 * nx] for i in range(G)]
    W = [None] * G
    for i in range(G - 2, -1, -1):
        W[i] = np.linalg.inv(np.eye(nx) + Cm[i] @ P[i + 1])
        p[i] = p[i] + F[i].T @ (W[i].T @ (p[i + 1] + P[i + 1] @ f[i]))
        P[i] = P[i] + F[i].T @ P[i + 1] @ W[i] @ F[i]
    xhat = np.zeros((G, nx))
    lam = np.zeros((G, nx))
    xhat[0] = x0
    for i in range(G - 1):
        xhat[i + 1] = W[i] @ (F[i] @ xhat[i] + f[i] - Cm[i] @ p[i + 1])
        lam[i] = P[i + 1] @ xhat[i + 1] + p[i + 1]
    return xhat, lam


def all_gather_rows(local_row, world: int, group=None):
    """all_gather of one fixed-size row per rank -> [world, len] tensor (NCCL for CUDA tensors, gloo for CPU)."""
    import torch
    import torch.distributed as dist
    out = torch.empty((world,) + tuple(local_row.shape), dtype=local_row.dtype, device=local_row.device)
    if world == 1:
        out[0] = local_row
        return out
    dist.all_gather([out[r] for r in range(world)], local_row.contiguous(), group=group)
    return out


class HorizonShardedSolver:
    """One long-horizon problem split into per-rank time slices (BASELINE.json config 5).

    solve_device(ws_prev_local, sigma, ws_out_local): update_problem_data + backward on the local slice, one
    all_gather of the slice summary, redundant interface solve, forward on the local slice."""

    def __init__(self, prob, rank: int, world: int, num_segments: int = 0, device: int = 0):
        import torch
        from .solver import Coupler, LQRCudaSolver
        from . import capi
        self.rank, self.world = rank, world
        start, count = horizon_slices(prob.N, world)[rank]
        self.start, self.count = start, count
        self.is_last = rank == world - 1
        self.local = slice_problem(prob, start, count, self.is_last)
        self.dev = torch.device("cuda", device)
        self.sol = LQRCudaSolver(prob.nx, prob.nu, count, batch=1, num_segments=num_segments, load_balancing=2,
                                 device=device)
        if not self.is_last:
            self.sol.set_option(capi.OPT_INTERIOR_SHARD, 1)
        self.sol.set_model(self.local)
        self.coupler = Coupler(prob.nx, prob.nu, world, batch=1, device=device)
        self.nx = prob.nx
        self.srec = self.sol.summary_doubles()
        self.my_sum = torch.empty(1, self.srec, dtype=torch.float64, device=self.dev)
        self.xhat = torch.empty(1, world, prob.nx, dtype=torch.float64, device=self.dev)
        self.lam = torch.empty(1, world, prob.nx, dtype=torch.float64, device=self.dev)
        self.x0 = torch.from_numpy(np.ascontiguousarray(prob.x0)).to(self.dev)

    def set_stream(self, ptr: int):
        self.sol.set_stream(ptr)
        self.coupler.set_stream(ptr)

    def solve_device(self, ws_prev, sigma, ws_out):
        self.sol.update_problem_data_device(ws_prev, sigma=sigma)
        self.sol.backward_device()
        self.sol.root_summary_device(self.my_sum)
        allsum = all_gather_rows(self.my_sum[0], self.world)            # [world, srec]  (the only collective)
        self.coupler.solve_device(allsum.unsqueeze(0).contiguous(), self.x0, self.xhat, self.lam)
        self.sol.set_root_boundary_device(self.xhat[:, self.rank].contiguous(), self.lam[:, self.rank].contiguous())
        self.sol.forward_device(self.x0, ws_out)
