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
    """Time slice [start, start+count) of a long-horizon Problem as its own Problem.  Constraint rows travel with their
    stages; the slice's "terminal" stage count is the next slice's first stage, so an interior slice has no terminal rows
    (and no terminal cost: its handle runs with PDPLQR_OPT_INTERIOR_SHARD)."""
    from .problems import Problem
    sl = slice(start, start + count)
    nx = prob.nx
    HN = prob.HN if is_last else np.zeros_like(prob.HN)
    hN = prob.hN if is_last else np.zeros_like(prob.hN)
    out = Problem(prob.nx, prob.nu, count, prob.batch, np.ascontiguousarray(prob.E[:, sl]),
                  np.ascontiguousarray(prob.c[:, sl]), np.ascontiguousarray(prob.H[:, sl]),
                  np.ascontiguousarray(prob.h[:, sl]), HN, hN, np.zeros((prob.batch, nx)),
                  name=f"{prob.name}[{start}:{start + count}]")
    if prob.ncs is not None:
        coff, doff = prob.coff(), prob.doff()
        ncs = np.zeros(count + 1, np.int32)
        ncs[:count] = prob.ncs[start:start + count]
        last_stage = start + count            # == prob.N for the last slice: its terminal rows belong to it
        if is_last:
            ncs[count] = prob.ncs[prob.N]
        c0, c1 = int(coff[start]), int(coff[last_stage + 1] if is_last else coff[last_stage])
        d0, d1 = int(doff[start]), int(doff[last_stage + 1] if is_last else doff[last_stage])
        out.ncs = ncs
        out.D = np.ascontiguousarray(prob.D[:, d0:d1])
        out.e_lb = None if prob.e_lb is None else np.ascontiguousarray(prob.e_lb[:, c0:c1])
        out.e_ub = None if prob.e_ub is None else np.ascontiguousarray(prob.e_ub[:, c0:c1])
        out.cones = [(k - start, r0, d, t) for (k, r0, d, t) in prob.cones
                     if start <= k < last_stage or (is_last and k == prob.N)]
        out.con_slice = (c0, c1)              # where this slice's ys / zs / rho live in the full vectors
    return out


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


def all_gather_rows(local_row, world: int, group=None, out=None):
    """all_gather of one fixed-size row per rank -> [world, len] tensor (NCCL for CUDA tensors, gloo for CPU).  `out`
    (persistent, contiguous) avoids a per-call allocation; on NCCL one all_gather_into_tensor, no list of views."""
    import torch
    import torch.distributed as dist
    if out is None:
        out = torch.empty((world,) + tuple(local_row.shape), dtype=local_row.dtype, device=local_row.device)
    if world == 1:
        out[0].copy_(local_row)
        return out
    if local_row.is_cuda:
        dist.all_gather_into_tensor(out, local_row, group=group)
    else:
        dist.all_gather([out[r] for r in range(world)], local_row.contiguous(), group=group)
    return out


class HorizonShardedSolver:
    """One long-horizon problem split into per-rank time slices (BASELINE.json config 5).

    solve_device(ws_prev_local, sigma, ws_out_local): update_problem_data + backward on the local slice, one
    all_gather of the slice summary, redundant interface solve, forward on the local slice.

    Stream discipline: the slice handle, the coupler, the collective and every torch op run on ONE stream -- torch's
    current stream of `device` at construction (or the one given to set_stream) -- so their order is the program order.
    All buffers that cross the phases are persistent members (nothing is handed back to torch's caching allocator while
    work that reads it is still queued)."""

    PHASES = ("local_sweep_and_tree", "all_gather", "coupler", "rollout")

    def __init__(self, prob, rank: int, world: int, num_segments: int = 0, device: int = 0, local=None):
        import torch
        from .solver import Coupler, LQRCudaSolver
        from . import capi
        self.rank, self.world = rank, world
        start, count = horizon_slices(prob.N, world)[rank] if local is None else (local.start, local.N)
        self.start, self.count = start, count
        self.is_last = rank == world - 1
        self.local = slice_problem(prob, start, count, self.is_last) if local is None else local
        self.dev = torch.device("cuda", device)
        nx, nu = self.local.nx, self.local.nu
        self.sol = LQRCudaSolver(nx, nu, count, batch=1, num_segments=num_segments, load_balancing=2,
                                 ncs=self.local.ncs, device=device)
        if not self.is_last:
            self.sol.set_option(capi.OPT_INTERIOR_SHARD, 1)
        self.sol.set_model(self.local)
        self.coupler = Coupler(nx, nu, world, batch=1, device=device)
        self.nx = nx
        self.srec = self.sol.summary_doubles()
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.my_sum = torch.empty(1, self.srec, **f64)
        self.allsum = torch.empty(world, self.srec, **f64)          # == [batch = 1][world][srec] for the coupler
        self.xhat = torch.empty(1, world, nx, **f64)
        self.lam = torch.empty(1, world, nx, **f64)
        self.x0 = torch.from_numpy(np.ascontiguousarray(prob.x0 if local is None else local.x0_global)).to(self.dev)
        self.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)

    def set_stream(self, ptr: int):
        """`ptr` must be the stream torch treats as current when solve_device runs (the collective is issued there)."""
        self.sol.set_stream(ptr)
        self.coupler.set_stream(ptr)

    def solve_device(self, ws_prev, sigma, ws_out, ys=None, zs=None, rho=None, inv_rho=None, events=None):
        """`events` (optional list) receives one torch.cuda.Event after every phase (PHASES) -- bench.py's breakdown."""
        import torch

        def mark():
            if events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                events.append(e)
        mark()
        self.sol.update_problem_data_device(ws_prev, ys, zs, inv_rho, sigma=sigma)
        self.sol.backward_device(rho)
        self.sol.root_summary_device(self.my_sum)
        mark()
        all_gather_rows(self.my_sum[0], self.world, out=self.allsum)       # the only collective
        mark()
        self.coupler.solve_device(self.allsum, self.x0, self.xhat, self.lam)
        self.sol.set_root_boundary_device(self.xhat[0, self.rank], self.lam[0, self.rank])
        mark()
        self.sol.forward_device(self.x0, ws_out)
        mark()
