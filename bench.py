#!/usr/bin/env python
"""bench.py -- LQ solves/sec of the B200-native PDP-LQR hot path (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c5|c1] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic problems:
    update_problem_data + backward (factorising) + forward          (lqr_solver_parallel.hpp:115-238)
Default workload = BASELINE.json configs[2] ("c3"): 65,536 independent cart-pole LQ problems, nx=4 nu=1 N=128,
per GPU -- the configuration the headline metric "LQ solves/sec (batched)" is quoted on and the one that shards
across 1/2/4/8 GPUs (independent problems, no data-path collective; weak scaling: 65,536 problems per rank).
`value`   : whole-job solves/s with model + iterates already resident in HBM (CUDA events on the launch stream).
`e2e`     : the same metric through the host-buffer C-ABI call pdplqr_solve() -- H2D of the iterate ws / x0 from
            pinned host memory and D2H of the solution inside the timed region (model resident, uploaded once by
            pdplqr_set_model like the reference's constructor builds its workspaces once).
`roofline`: dominant kernel (backward sweep) against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
`cpu_baseline` / `--impl reference`: the CPU oracle port (the reference needs Eigen3, absent from this image) on
            the box's host cores, bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout must carry exactly ONE line (the JSON).  Libraries (NCCL prints a version banner) write to fd 1 directly, so
# fd 1 is pointed at stderr for the whole run and emit() writes the JSON line to the saved, real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def host_cores() -> int:
    """Cores this process may use (torchrun sets OMP_NUM_THREADS=1, which must not throttle the CPU baseline)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


METRIC = "lq_solves_per_sec"
UNIT = "solves/s"


# --------------------------------------------------------------------------------------------- workloads
def make_workload(name: str, rank: int):
    import pdplqr_b200 as P
    if name == "c3":
        prob = P.problems.cartpole_batch(batch=65536, N=128, seed=1234 + 7919 * rank)
        return prob, dict(num_segments=1), "C3: 65536 x cart-pole LQR nx=4 nu=1 N=128 per GPU (BASELINE.json configs[2])"
    if name == "c3small":
        prob = P.problems.cartpole_batch(batch=4096, N=128, seed=1234 + 7919 * rank)
        return prob, dict(num_segments=1), "C3-small: 4096 x cart-pole LQR nx=4 nu=1 N=128 per GPU (debug size)"
    if name == "c2":
        prob = P.problems.quadrotor_ltv(1024)
        return prob, dict(num_segments=0), "C2: quadrotor LQR nx=12 nu=4 N=1024, single problem, segment-parallel (configs[1])"
    if name == "c5":
        prob = P.problems.quadrotor_ltv(1 << 20)
        return prob, dict(num_segments=wave_aligned((1 << 20) // 250), load_balancing=2), \
            "C5: quadrotor LQR nx=12 nu=4 N=2^20, single problem, ~250-stage segments in 2 whole waves of 148x14 CTAs (configs[4])"
    if name == "c5small":
        prob = P.problems.quadrotor_ltv(1 << 16)
        return prob, dict(num_segments=(1 << 16) // 64, load_balancing=False), \
            "C5-small: quadrotor LQR nx=12 nu=4 N=2^16 (debug size)"
    if name == "c1":
        prob = P.problems.quadrotor_example()
        return prob, dict(num_segments=4), "C1: examples/lqr_example.cpp as shipped (configs[0])"
    raise SystemExit(f"unknown workload {name}")


def c4_traffic(batch: int):
    """Measured DRAM bytes per launch of the affine sweep (profiles/traffic.json, taken at batch 4096), scaled to `batch`:
    well below the algorithmic figure because the sweep reads only [E|c], h and the cached affine terms."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    t = json.load(open(path)).get("c4_affine_batch4096")
    return None if t is None else t * batch / 4096.0


WAVE = 148 * 14   # resident one-warp stage-kernel CTAs per GPU at nx12/nu4 (16 KB of shared memory each)


def wave_aligned(num_segments: int, wave: int = WAVE) -> int:
    """Round a segment count down to whole waves of the stage kernel's resident CTAs (148 SMs x 14 one-warp CTAs at
    nx12/nu4): a trailing partial wave costs a full wave of time (2048 segments = 1.06 waves ran as 2)."""
    if num_segments <= wave:
        return max(1, num_segments)
    return (num_segments // wave) * wave


def algorithmic_bytes_per_stage(nx, nu, pdp: bool):
    """doubles per stage, LTV storage, carrying P,p,F,f,C on chip (nc = 0).
    Returns (backward_bytes, forward_bytes); their sum is the survey's B_stage (760 B for nx4/nu1 sequential,
    7424 B for nx12/nu4 PDP)."""
    s = nx + nu
    model = nx * s + nx + s * s + s
    fac = s * nu + nu + (nu * nx if pdp else 0)
    bwd = 8 * (model + fac)
    fwd = 8 * (fac + nx * s + nx + s)
    return bwd, fwd


# --------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms from before the warm-up until the last timed loop
    (B200_PROFILING.md recipe); `summary(windows)` keeps the samples that fall inside the timed windows."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       "50", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def summary(self, windows):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]),
                             [n for n, v in zip(names, parts[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if any(a - 0.03 <= r[0] <= b + 0.03 for a, b in windows)]
        if not inside and rows:   # timed windows shorter than the sampling period: nearest samples under load
            lo, hi = min(a for a, _ in windows), max(b for _, b in windows)
            inside = [r for r in rows if lo - 0.25 <= r[0] <= hi + 0.25]
        if inside:
            out.update(sm_mhz=statistics.median(r[1] for r in inside), sm_max_mhz=max(r[2] for r in inside),
                       reasons=sorted({n for r in inside for n in r[3]}), samples=len(inside))
        return out


def cpu_baseline(prob, seconds=6.0, max_problems=16384):
    """Time the CPU oracle port (OpenMP over problems, sequential Riccati each -- or the PDP solver for a single
    long problem) on a bounded sample of the workload.  Returns (solves_per_s, cores, sample_description)."""
    from oracle import oracle as O
    cores = host_cores()
    if prob.batch > 1:
        nb = min(prob.batch, max_problems)
        sub = prob.select(slice(0, nb))
        pool = O.OracleBatch(sub)
        ws_in = sub.zeros_ws()
        out = np.empty_like(ws_in)
        pool.solve(ws_in=ws_in, ws_out=out, nthreads=cores)  # warm-up
        best, reps, t_end = 1e30, 0, time.perf_counter() + seconds
        while reps < 2 or time.perf_counter() < t_end:
            t0 = time.perf_counter()
            pool.solve(ws_in=ws_in, ws_out=out, nthreads=cores)
            best = min(best, time.perf_counter() - t0)
            reps += 1
        return nb / best, cores, f"first {nb} of {prob.batch} problems, full solve, best of {reps} reps, {cores} OpenMP threads"
    S = max(1, min(8, cores))
    n_stages = min(prob.N, 1 << 17)
    sub = prob if n_stages == prob.N else None
    if sub is None:
        import pdplqr_b200 as P
        sub = P.problems.quadrotor_ltv(n_stages)
    o = O.OracleSolver(sub, parallel=S > 1, num_segments=S, load_balancing=True, condensed=O.CHOLESKY, nthreads=S)
    ws_in = np.zeros(sub.ws_len)
    out = np.zeros(sub.ws_len)

    def one():
        o.update_problem_data(ws_in, sigma=1e-6)
        o.backward()
        o.forward(sub.x0[0], out)
    one()
    best, reps, t_end = 1e30, 0, time.perf_counter() + seconds
    while reps < 3 or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        one()
        best = min(best, time.perf_counter() - t0)
        reps += 1
    scale = prob.N / n_stages  # a longer horizon costs proportionally more stage steps
    return 1.0 / (best * scale), S, (f"PDP oracle, S={S} segments on {S} threads, N={n_stages}"
                                     + (f" (time scaled x{scale:.0f} to N={prob.N})" if scale != 1 else "")
                                     + f", best of {reps} reps")


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    prob, _, desc = make_workload(args.workload, 0)
    times = []
    val, cores, sample = None, None, None
    per_step = min(1.0, 120.0 / max(1, args.warmup + args.steps))   # whole arm bounded to ~2 minutes of CPU timing
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_baseline(prob, seconds=per_step, max_problems=8192)
        if i >= args.warmup:
            times.append(v)
    val = statistics.median(times)
    solves_per_step = prob.batch
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * solves_per_step / val, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "cpu_only": True},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference needs Eigen3 (absent): timed the CPU oracle port of the same algorithm"}
    emit(line)



# --------------------------------------------------------------------------------------------- config 4 (conic ADMM)
def run_c4(args, rank, world, local_rank):
    """BASELINE.json configs[3]: conic-constrained (box + SOC) LQ MPC nx=30 nu=10 N=256, batch 4096 per GPU, full outer
    iterations.  A step = one ADMM solve with a FIXED number of outer iterations (1 factorising + ITERS-1 affine-only
    LQ solves, projections, residuals), all device-resident.  The outer iteration is not in the reference (hooks only)."""
    import torch
    import torch.distributed as dist
    import pdplqr_b200 as P
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = int(os.environ.get("C4_BATCH", "4096"))
    N = int(os.environ.get("C4_N", "256"))
    ITERS = int(os.environ.get("C4_ITERS", "50"))
    base = 64 if B % 64 == 0 else B
    sampler = ClockSampler(local_rank)
    hp = P.problems.random_conic_batch(batch=base, N=N, seed=99 + rank)
    rep = B // base
    nx, nu, s = hp.nx, hp.nu, hp.s
    sol = P.LQRCudaSolver(nx, nu, N, batch=B, num_segments=1, ncs=hp.ncs, device=local_rank)
    stream = torch.cuda.current_stream()
    sol.set_stream(stream.cuda_stream)

    def up(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        return t.repeat(rep, *([1] * (t.dim() - 1))).contiguous() if rep > 1 else t
    dE, dc, dH, dh, dHN, dhN, dD = (up(a) for a in (hp.E, hp.c, hp.H, hp.h, hp.HN, hp.hN, hp.D))
    sol.set_model_device(dE, dc, dH, dh, dHN, dhN, dD)
    torch.cuda.synchronize()
    del dE, dc, dH, dh, dD
    lb = np.where(np.isfinite(hp.e_lb), hp.e_lb, -1e20)
    ub = np.where(np.isfinite(hp.e_ub), hp.e_ub, 1e20)
    sol.admm_set_cones(hp.cones, np.tile(lb, (rep, 1)), np.tile(ub, (rep, 1)))
    nct = hp.nc_total
    x0 = up(hp.x0)
    rho = torch.full((B, nct), 0.1, dtype=torch.float64, device=dev)
    inv_rho = 1.0 / rho
    w = torch.zeros(B, hp.ws_len, dtype=torch.float64, device=dev)
    z = torch.zeros(B, nct, dtype=torch.float64, device=dev)
    y = torch.zeros(B, nct, dtype=torch.float64, device=dev)

    def step():
        w.zero_(); z.zero_(); y.zero_()
        return sol.admm_solve_device(x0, w, z, y, rho, inv_rho, sigma=1e-6, alpha=1.6, max_iter=ITERS, eps_abs=0.0,
                                     eps_rel=0.0, check_every=ITERS)
    steps = max(1, min(args.steps, int(os.environ.get("C4_STEPS", "3"))))
    for _ in range(min(args.warmup, 1)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_w0 = time.time()
    l0 = sol.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        it, res = step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_step = e0.elapsed_time(e1) / steps
    launches = sol.launch_count() - l0
    windows = [(t_w0, time.time())]
    # affine-only iteration alone (the common ADMM iteration): CUDA events around update + backward_without_factorization
    ka = 10
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(stream)
    for _ in range(ka):
        sol.update_problem_data_device(w, y, z, inv_rho, sigma=1e-6)
        sol.backward_without_factorization_device(rho)
    a1.record(stream)
    torch.cuda.synchronize()
    ms_aff = a0.elapsed_time(a1) / ka
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    sol.update_problem_data_device(w, y, z, inv_rho, sigma=1e-6)
    sol.backward_device(rho)
    f1.record(stream)
    torch.cuda.synchronize()
    ms_fact = f0.elapsed_time(f1)
    clocks = sampler.summary(windows)
    t = torch.tensor([ms_step, ms_aff, ms_fact], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, ms_aff, ms_fact = (float(v) for v in t.cpu())
    if rank == 0:
        nc = int(hp.ncs[1])
        aff_bytes = 8 * (nx * s + s + nu * nx + nu * nu + 2 * nx + nc * s + 3 * nc + s)
        fact_bytes = 55488
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = aff_bytes * B * N / (ms_aff * 1e-3) / 1e9
        line = {"metric": METRIC, "value": world * B * 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": min(args.warmup, 1), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C4: conic (box + SOC) LQ MPC nx=30 nu=10 N=%d, batch %d per GPU, %d fixed ADMM outer iterations per solve (configs[3])" % (N, B, ITERS),
                           "problems_per_gpu": B, "nx": nx, "nu": nu, "N": N, "nc": nc, "admm_iterations": ITERS,
                           "problem_iterations_per_sec": world * B * ITERS * 1e3 / ms_step,
                           "ms_factorizing_backward": ms_fact, "ms_affine_backward": ms_aff,
                           "final_residuals": [float(res[0]), float(res[1])],
                           "l2": "inputs larger than L2", "sharding": "independent problems per rank, no data-path collective"},
                "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"], "samples": clocks["samples"]},
                "e2e": None, "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "seg_affine_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": c4_traffic(B), "algorithmic_bytes_per_launch": aff_bytes * B * N,
                             "kernel_ms": ms_aff,
                             "factorizing_kernel": {"kernel": "seg_backward_kernel<30,10,128>", "ms": ms_fact,
                                                    "achieved": fact_bytes * B * N / (ms_fact * 1e-3) / 1e9,
                                                    "frac": fact_bytes * B * N / (ms_fact * 1e-3) / 1e9 / peak}}}
        if not args.no_cpu_baseline and world == 1:
            from oracle import oracle as O
            nb = min(base, 2 * host_cores())
            sub = hp.select(slice(0, nb))
            pool = O.OracleBatch(sub)
            rng = np.random.default_rng(0)
            ys_, zs_ = rng.standard_normal((nb, nct)), rng.standard_normal((nb, nct))
            rh = np.full((nb, nct), 0.1)
            nt = host_cores()
            t0 = time.perf_counter(); pool.solve(ys=ys_, zs=zs_, rho=rh, inv_rho=1.0 / rh, factorize=True, nthreads=nt); tf = time.perf_counter() - t0
            t0 = time.perf_counter(); pool.solve(ys=ys_, zs=zs_, rho=rh, inv_rho=1.0 / rh, factorize=False, nthreads=nt); tn = time.perf_counter() - t0
            v = nb / (tf + (ITERS - 1) * tn)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": host_cores(), "kind": "port",
                                    "sample": "%d problems: 1 factorising + 1 affine-only LQ solve timed, extrapolated to %d iterations (projections not counted)" % (nb, ITERS)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()

# --------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import pdplqr_b200 as P

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if args.workload == "c4":
        run_c4(args, rank, world, local_rank)
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sampler = ClockSampler(local_rank)
    windows = []
    horizon_sharded = args.workload in ("c5", "c5small") and world > 1
    prob, kw, desc = make_workload(args.workload, 0 if horizon_sharded else rank)
    stream = torch.cuda.current_stream()
    if horizon_sharded:
        # one long problem, contiguous time slices per rank, ONE all_gather of a 3,648-byte summary per solve
        from pdplqr_b200.sharding import HorizonShardedSolver
        # per rank: the single-GPU segment length, but never less than one full wave of (problem, segment) CTAs -- with a
        # partial wave every SM runs fewer warps than it can hold and the sweep is a pure latency chain
        hs = HorizonShardedSolver(prob, rank, world, num_segments=wave_aligned(max(kw["num_segments"] // world, WAVE)),
                                  device=local_rank)
        hs.set_stream(stream.cuda_stream)
        sol = hs.sol
        full_N = prob.N
        prob = hs.local
        prob.x0 = hs.x0.cpu().numpy()
    else:
        sol = P.LQRCudaSolver.from_problem(prob, device=local_rank, **kw)
        sol.set_stream(stream.cuda_stream)
        full_N = prob.N
    pdp = sol.num_segments > 1 or horizon_sharded
    B, N, nx, nu, s = prob.batch, prob.N, prob.nx, prob.nu, prob.s

    # device-resident iterates (ADMM would update ws between solves; here a fixed seeded iterate)
    rng = np.random.default_rng(17 + rank)
    ws_host = torch.from_numpy(0.01 * rng.standard_normal((B, prob.ws_len))).pin_memory()
    x0_host = torch.from_numpy(np.ascontiguousarray(prob.x0)).pin_memory()
    out_host = torch.empty_like(ws_host).pin_memory()
    ws_dev = ws_host.to(dev)
    x0_dev = x0_host.to(dev)
    out_dev = torch.empty_like(ws_dev)
    sigma = 1e-6

    def step_device():
        if horizon_sharded:
            hs.solve_device(ws_dev, sigma, out_dev)
            return
        sol.update_problem_data_device(ws_dev, sigma=sigma)
        sol.backward_device()
        sol.forward_device(x0_dev, out_dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    t_w0 = time.time()
    l0 = sol.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    launches = sol.launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    windows.append((t_w0, time.time()))
    bad, _ = sol.last_status()

    # dominant kernel alone (backward sweep): CUDA events around back-to-back launches on the same stream
    kb = max(5, args.steps)
    t_w0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l_b0 = sol.launch_count()
    e0.record(stream)
    for _ in range(kb):
        sol.update_problem_data_device(ws_dev, sigma=sigma)
        sol.backward_device()
    e1.record(stream)
    torch.cuda.synchronize()
    bwd_launches = (sol.launch_count() - l_b0) // kb
    ms_bwd = e0.elapsed_time(e1) / kb
    windows.append((t_w0, time.time()))
    # leave the handle in a consistent state (one forward per backward)
    if horizon_sharded:
        hs.solve_device(ws_dev, sigma, out_dev)
    else:
        sol.forward_device(x0_dev, out_dev)
    torch.cuda.synchronize()

    # end-to-end through the host-buffer C ABI (pinned host memory; H2D + kernels + D2H inside the timed region)
    ws_np, x0_np, out_np = ws_host.numpy(), x0_host.numpy(), out_host.numpy()
    e2e_steps = max(3, min(args.steps, 10))

    def step_e2e():
        if horizon_sharded:   # host slice in, sharded solve, host slice out (pinned buffers)
            ws_dev.copy_(ws_host, non_blocking=True)
            hs.solve_device(ws_dev, sigma, out_dev)
            out_host.copy_(out_dev, non_blocking=True)
            torch.cuda.synchronize()
        else:
            sol.solve(ws_np, x0_np, out_np, sigma=sigma)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    windows.append((time.time() - t_e2e * e2e_steps, time.time()))
    clocks = sampler.summary(windows)
    same = bool(np.array_equal(out_np, out_dev.cpu().numpy()))

    t = torch.tensor([ms_total, t_e2e * 1e3, ms_bwd], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_bwd = (float(v) for v in t.cpu())

    if rank == 0:
        ms_step = ms_total / args.steps
        solves_per_step = B if horizon_sharded else world * B   # a sharded long horizon is ONE solve across all ranks
        value = solves_per_step * 1e3 / ms_step
        bwd_b, fwd_b = algorithmic_bytes_per_stage(nx, nu, pdp)
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = bwd_b * B * N / (ms_bwd * 1e-3) / 1e9
        traffic = None
        tr_path = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr_path):
            traffic = json.load(open(tr_path)).get(args.workload)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if horizon_sharded else "weak",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "problems_per_gpu": B, "nx": nx, "nu": nu, "N": N,
                       "num_segments": sol.num_segments, "sigma": sigma,
                       "l2": "inputs larger than L2 (model records %.2f GB per GPU)" % (B * N * sol.record_doubles()[0] * 8 / 1e9),
                       "sharding": ("horizon: contiguous time slices per rank, one all_gather of a %d-byte slice summary per solve"
                                    % (sol.summary_doubles() * 8)) if horizon_sharded else
                                   "independent problems per rank, no data-path collective",
                       "full_horizon": full_N},
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                       "samples": clocks["samples"]},
            "e2e": {"value": solves_per_step / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(ws_host.numel() * 8 + x0_host.numel() * 8),
                    "d2h_bytes_per_step": int(out_host.numel() * 8), "ms_per_step": ms_e2e,
                    "model": "resident (uploaded once by pdplqr_set_model)", "matches_device_path": same},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "batch_backward_kernel" if sol.num_segments == 1 and nx + nu <= 8 else "seg_backward_kernel",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "traffic": traffic, "algorithmic_bytes_per_launch": bwd_b * B * N,
                         "kernel_ms": ms_bwd, "launches_per_backward": int(bwd_launches),
                         "step_hbm_frac": (bwd_b + fwd_b) * B * N / (ms_step * 1e-3) / 1e9 / peak},
            "non_pd_problems": int(bad),
        }
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample = cpu_baseline(prob)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
