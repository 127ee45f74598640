#!/usr/bin/env python
"""bench.py -- LQ solves/sec of the B200-native PDP-LQR hot path (BASELINE.json metric), ONE JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--legs c2,c5,c4] [--workload c3|c2|c5|c4|c1]
                    [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic problems:
    update_problem_data + backward (factorising) + forward          (lqr_solver_parallel.hpp:115-238)
Headline (`value`, `e2e`, `roofline`, `cpu_baseline`): BASELINE.json configs[2] ("c3"): 65,536 independent cart-pole LQ
problems, nx=4 nu=1 N=128, per GPU -- the configuration the metric "LQ solves/sec (batched)" is quoted on and the one
that shards across 1/2/4/8 GPUs without a collective (weak scaling: 65,536 problems per rank).
`configs` block: the other BASELINE.json configurations measured in the SAME run, each with its own ms_per_step,
roofline (HBM and FP64-pipe fractions), cpu_baseline, e2e and `parity_rel_err` (max relative error of THIS run's
output against the CPU oracle):
    c2  quadrotor nx12/nu4 N=1024, one problem, segment-parallel: per-solve latency (one CUDA graph launch), and latency vs N
        (`latency_vs_N_us`: GPU us, parity, and the CPU path timed beside every N -- the faster of the sequential and the
        PDP oracle port on 2 / 4 / 8 threads)
    c5  quadrotor N=2^20: at N GPUs > 1 the horizon is split into per-rank time slices with ONE NCCL all_gather of a
        3,648-byte summary per solve (strong scaling) + a per-phase breakdown and `parity_vs_1gpu`; best of three K-step loops
    c4  conic (box + SOC) LQ MPC nx30/nu10 N=256, batch 4096 per GPU, full ADMM outer iterations: 50 fixed iterations per
        solve (the step; one CUDA graph launch, on a side stream), and once to tolerance 1e-4 with rho adaptation (`to_tolerance`)
`value`   : whole-job solves/s with model + iterates resident in HBM (CUDA events on the launch stream, max over ranks).
`e2e`     : the same metric through the host-buffer C-ABI call pdplqr_solve() -- H2D of the iterate ws / x0 from pinned
            host memory and D2H of the solution inside the timed region (model resident, uploaded once by
            pdplqr_set_model like the reference's constructor builds its workspaces once); `link_frac` relates it to the
            host link bandwidth measured in the same run (all ranks copying at once).
`cpu_baseline` / `--impl reference`: the CPU oracle port on the box's host cores (the reference itself needs Eigen3,
            absent from this image; the shim-backed build of the reference's own headers, oracle/_ref, is timed beside it).
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
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout must carry exactly ONE line (the JSON).  Libraries (NCCL prints a version banner) write to fd 1 directly, so
# fd 1 is pointed at stderr for the whole run and emit() writes the JSON line to the saved, real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def host_cores() -> int:
    """Cores this process may use (torchrun sets OMP_NUM_THREADS=1, which must not throttle the CPU baseline)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


METRIC = "lq_solves_per_sec"
UNIT = "solves/s"
SIGMA = 1e-6
# FP64 peak of this pool's B200s, measured with scripts/micro/dmma_bench.cu (profiles/r1_dmma_microbench.txt):
# 36.8 TFLOP/s on the FP64 tensor pipe (mma.sync.m8n8k4.f64 = SASS DMMA), 31.9 TFLOP/s on the DFMA pipe.
FP64_PEAK_TFLOPS = 36.8
FP64_PEAK_SOURCE = "profiles/r1_dmma_microbench.txt (DMMA, measured on this pool; DFMA pipe 31.9)"
C3_BATCH, C3_N = 65536, 128
C5_N = 1 << 20
C4_BATCH, C4_N, C4_ITERS = 4096, 256, 50
LATENCY_NS = (100, 256, 1024, 4096, 16384)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path)).get("hbm_gbs", 6650.0)), "MEASURED_PEAKS.json hbm_gbs (measured)"
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def measured_traffic(key):
    """DRAM bytes per launch from an `ncu --set full` capture (profiles/traffic.json) -- a recorded constant of an
    earlier capture of the same kernel and workload, NOT measured in this run."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get(key)


# --------------------------------------------------------------------------------------------- byte / flop models
def stage_bytes(nx, nu, pdp: bool, sym_h: bool = False):
    """algorithmic bytes per stage (nc = 0), LTV storage, carrying P,p,F,f,C on chip.
    Returns (backward, forward); their sum is the survey's B_stage (760 B nx4/nu1 sequential, 7424 B nx12/nu4 PDP).
    sym_h: count only the lower triangle of H (what the thread-per-problem record actually holds)."""
    s = nx + nu
    nH = s * (s + 1) // 2 if sym_h else s * s
    model = nx * s + nx + nH + s
    fac = s * nu + nu + (nu * nx if pdp else 0)
    return 8 * (model + fac), 8 * (fac + nx * s + nx + s)


def stage_flops(nx, nu, nc=0, pdp=False):
    """flops per stage as the reference executes them (dense): (factorising backward, forward)."""
    s = nx + nu
    f_seq = 2 * s * nx * nx + 2 * s * s * nx + s ** 3 / 3 + 4 * nx * nx + 2 * s * nx + nu * nu + 2 * nx * nu
    f_par = 2 * nx * nu * nu + 6 * nx * nx * nu + 2 * nx ** 3 + 2 * nx * nu + 2 * nx * nx + 2 * nu * nu
    f_con = nc * s + 2 * s * s * nc + nc + 2 * s * nc
    f_fwd = 2 * nx * nu + nu * nu + 2 * nx * s + (2 * nu * nx if pdp else 0)
    return f_seq + (f_par if pdp else 0) + (f_con if nc else 0), f_fwd


def roofline(kernel, bytes_per_launch, flops_per_launch, ms, traffic_key=None, extra=None):
    peak, src = peaks()
    gbs = bytes_per_launch / (ms * 1e-3) / 1e9
    tfs = flops_per_launch / (ms * 1e-3) / 1e12
    bound = "hbm" if gbs / peak >= tfs / FP64_PEAK_TFLOPS else "fp64"
    out = {"bound": bound, "kernel": kernel, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
           "peak_source": src, "algorithmic_bytes_per_launch": int(bytes_per_launch), "kernel_ms": ms,
           "fp64": {"achieved": tfs, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tfs / FP64_PEAK_TFLOPS,
                    "peak_source": FP64_PEAK_SOURCE, "algorithmic_flops_per_launch": int(flops_per_launch)},
           "traffic": measured_traffic(traffic_key) if traffic_key else None,
           "traffic_source": ("profiles/traffic.json['%s'] (recorded ncu --set full capture, not measured in this run)" % traffic_key)
           if traffic_key and measured_traffic(traffic_key) else None}
    if extra:
        out.update(extra)
    return out


WAVE = 148 * 14   # fallback when no GPU is visible (the round-1 stage kernel: 14 one-warp CTAs per SM at nx12/nu4)


def stage_wave() -> int:
    """(problem, segment) groups the GPU keeps resident in the nx12/nu4 stage sweep (pdplqr_wave_size: SMs x CTAs per SM
    from the occupancy calculator)."""
    try:
        import pdplqr_b200 as P
        return P.wave_size(12, 4)
    except Exception:
        return WAVE


def wave_aligned(num_segments: int, wave: int = WAVE) -> int:
    """Round a segment count down to whole waves of the stage kernel's resident CTAs: a trailing partial wave costs a
    full wave of time (2048 segments = 1.06 waves ran as 2)."""
    if num_segments <= wave:
        return max(1, num_segments)
    return (num_segments // wave) * wave


def c4_traffic(batch: int):
    t = measured_traffic("c4_affine_batch4096")
    return None if t is None else t * batch / 4096.0


def c5_segments(world: int = 1) -> int:
    """Segments per rank for the 2^20-stage problem: whole waves of the stage kernel, never less than one full wave per rank.
    Measured at 1 GPU (scripts/sweep_c5_seglen.sh): one wave (2,368 segments of 443 stages) 2.057 ms, two waves 2.110, three
    2.117, four 2.165 ms -- the stage sweep does not care, the interface tree gets shorter."""
    wave = stage_wave()
    per_rank = wave_aligned(max(C5_N // int(os.environ.get("BENCH_C5_SEG_LEN", "450")), wave), wave)
    return wave_aligned(max(per_rank // world, wave), wave)


def workload_config(name: str, world: int) -> dict:
    """Deterministic description of a workload (identical in the GPU arm and in `--impl reference`)."""
    if name == "c3":
        return {"workload": "C3: 65536 x cart-pole LQR nx=4 nu=1 N=128 per GPU (BASELINE.json configs[2])",
                "problems_per_gpu": C3_BATCH, "nx": 4, "nu": 1, "N": C3_N, "num_segments": 1, "sigma": SIGMA,
                "l2": "inputs larger than L2 (model records 2.95 GB per GPU)",
                "sharding": "independent problems per rank, no data-path collective"}
    if name == "c2":
        return {"workload": "C2: quadrotor LQR nx=12 nu=4 N=1024, single problem, segment-parallel (configs[1])",
                "problems_per_gpu": 1, "nx": 12, "nu": 4, "N": 1024, "sigma": SIGMA,
                "l2": "latency-bound: the whole problem (7.6 MB) is L2-resident by construction",
                "sharding": "replicas only (one problem per GPU)"}
    if name == "c5":
        return {"workload": "C5: quadrotor LQR nx=12 nu=4 N=2^20, single problem (configs[4])",
                "problems_per_gpu": 1, "nx": 12, "nu": 4, "N": C5_N, "sigma": SIGMA,
                "l2": "inputs larger than L2 (model records 3.99 GB)",
                "sharding": "1 GPU: whole horizon" if world == 1 else
                "horizon: contiguous time slices per rank, one all_gather of a 3648-byte slice summary per solve"}
    if name == "c4":
        return {"workload": "C4: conic (box + SOC) LQ MPC nx=30 nu=10 N=%d, batch %d per GPU, %d fixed ADMM outer iterations per solve (configs[3])" % (C4_N, C4_BATCH, C4_ITERS),
                "problems_per_gpu": C4_BATCH, "nx": 30, "nu": 10, "N": C4_N, "nc": 44, "admm_iterations": C4_ITERS,
                "sigma": SIGMA, "l2": "inputs larger than L2",
                "sharding": "independent problems per rank, no data-path collective"}
    if name == "c1":
        return {"workload": "C1: examples/lqr_example.cpp as shipped (configs[0])", "problems_per_gpu": 1, "nx": 12,
                "nu": 4, "N": 100, "num_segments": 4, "sigma": SIGMA, "l2": "latency-bound (0.74 MB)",
                "sharding": "replicas only"}
    raise SystemExit(f"unknown workload {name}")


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


class Ctx:
    """Per-process run context (one process per GPU)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch, self.dist = torch, dist
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            import datetime
            # a rank that fails inside an auxiliary leg must not leave the others waiting for the default 10 minutes
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=240))
        self.stream = torch.cuda.current_stream()
        self.windows = []

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    def time_loop(self, fn, steps, warmup, stream=None):
        """W untimed + K timed calls of fn() bracketed by barrier + synchronize; CUDA events on the launch stream;
        returns ms per step as the MAX over ranks."""
        torch = self.torch
        stream = stream or self.stream
        for _ in range(warmup):
            fn()
        self.barrier()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        self.barrier()
        self.windows.append((t0, time.time()))
        return self.max_over_ranks([e0.elapsed_time(e1) / steps])[0]

    def time_wall(self, fn, steps, warmup=1):
        """Same bracket, host wall clock (end-to-end calls that synchronise themselves)."""
        for _ in range(warmup):
            fn()
        self.barrier()
        t0w = time.time()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        self.windows.append((t0w, time.time()))
        self.barrier()
        return self.max_over_ranks([dt * 1e3])[0]


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def link_bandwidth(ctx, nbytes=256 << 20):
    """Pinned host <-> device copy bandwidth of this box measured in this run, every rank copying at the same time
    (the ranks share the host link): GB/s per rank for H2D alone, D2H alone, and both directions at once."""
    torch = ctx.torch
    n = nbytes // 8
    hin = torch.empty(n, dtype=torch.float64).pin_memory()
    hout = torch.empty(n, dtype=torch.float64).pin_memory()
    din, dout = torch.empty(n, dtype=torch.float64, device=ctx.dev), torch.zeros(n, dtype=torch.float64, device=ctx.dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for name in ("h2d", "d2h", "duplex"):
        for rep in range(2):   # first repetition warms the path up
            ctx.barrier()
            t0 = time.perf_counter()
            if name in ("h2d", "duplex"):
                with torch.cuda.stream(s1):
                    din.copy_(hin, non_blocking=True)
            if name in ("d2h", "duplex"):
                with torch.cuda.stream(s2):
                    hout.copy_(dout, non_blocking=True)
            s1.synchronize(); s2.synchronize()
            dt = time.perf_counter() - t0
        dt = ctx.max_over_ranks([dt])[0]
        res[name] = nbytes / dt / 1e9           # per direction
    return res


def e2e_block(ctx, value, ms, h2d, d2h, link, extra=None):
    out = {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ms,
           "model": "resident (uploaded once by pdplqr_set_model)"}
    if link:
        ideal_ms = max(h2d / link["duplex"], d2h / link["duplex"]) / 1e6   # both directions overlap (two copy engines)
        out.update(link_gbs_per_rank=link, link_ideal_ms=ideal_ms, link_frac=ideal_ms / ms if ms > 0 else None,
                   link_note="pinned-copy GB/s per rank measured in this run with all %d rank(s) copying at once; "
                             "link_frac = (bytes / duplex rate) / e2e time" % ctx.world)
    if extra:
        out.update(extra)
    return out


# --------------------------------------------------------------------------------------------- CPU legs (oracle port)
def cpu_batch(prob, seconds=6.0, max_problems=None, threads=None):
    """CPU oracle port, OpenMP over problems, sequential Riccati each.  Returns (solves/s, cores, sample text)."""
    from oracle import oracle as O
    cores = threads or host_cores()
    nb = prob.batch if max_problems is None else min(prob.batch, max_problems)
    sub = prob if nb == prob.batch else prob.select(slice(0, nb))
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
    what = "all %d problems" % nb if nb == prob.batch else "first %d of %d problems" % (nb, prob.batch)
    return nb / best, cores, f"{what}, full solve, best of {reps} reps, {cores} OpenMP threads"


def cpu_single(prob, seconds=4.0, max_stages=1 << 17):
    """CPU oracle port, the reference's PDP solver (S segments on S threads) on one long problem."""
    from oracle import oracle as O
    import pdplqr_b200 as P
    S = max(1, min(8, host_cores()))
    n_stages = min(prob.N, max_stages)
    sub = prob if n_stages == prob.N else P.problems.quadrotor_ltv(n_stages)
    o = O.OracleSolver(sub, parallel=S > 1, num_segments=S, load_balancing=True, condensed=O.CHOLESKY, nthreads=S)
    ws_in, out = np.zeros(sub.ws_len), np.zeros(sub.ws_len)

    def one():
        o.update_problem_data(ws_in, sigma=SIGMA)
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


def cpu_latency(prob, seconds=0.6):
    """CPU oracle port, latency of ONE solve of `prob`: the sequential LQRSolver and the reference's PDP solver with S segments
    on S threads (S = 2, 4, 8 <= host cores), each best-of-reps; returns the faster of them next to the sequential time."""
    from oracle import oracle as O
    ws_in, out = np.zeros(prob.ws_len), np.zeros(prob.ws_len)
    res = {}
    cands = [1] + [S for S in (2, 4, 8) if S <= host_cores() and prob.N >= 8 * S]
    for S in cands:
        o = O.OracleSolver(prob, parallel=S > 1, num_segments=S, load_balancing=True, condensed=O.CHOLESKY, nthreads=S)

        def one():
            o.update_problem_data(ws_in, sigma=SIGMA)
            o.backward()
            o.forward(prob.x0[0], out)
        one()
        best, reps, t_end = 1e30, 0, time.perf_counter() + seconds / len(cands)
        while reps < 3 or time.perf_counter() < t_end:
            t0 = time.perf_counter()
            one()
            best = min(best, time.perf_counter() - t0)
            reps += 1
        res[S] = best * 1e6
        del o
    S_best = min(res, key=res.get)
    return {"cpu_us": res[S_best], "cpu_threads": S_best, "cpu_sequential_us": res[1],
            "cpu_kind": "port (oracle/liboracle.so): best of the sequential LQRSolver and the PDP solver with S = threads in %s"
                        % (tuple(cands),)}


def cpu_reference_build(prob, seconds=5.0, max_problems=256):
    """The reference's OWN solver class (lqr::LQRSolver compiled from /root/reference's headers against the Eigen-API
    shim: oracle/_ref/libpdpref.so) on a sample of the batch, one problem per thread at a time."""
    from oracle import reflib
    if not reflib.available():
        return None
    from concurrent.futures import ThreadPoolExecutor
    cores = host_cores()
    nb = min(prob.batch, max_problems)
    sols = [reflib.ReferenceSolver(prob, b=b) for b in range(nb)]
    ws_in = np.zeros(prob.ws_len)

    def one(b):
        sols[b].update_problem_data(ws_in, sigma=SIGMA)
        sols[b].backward()
        sols[b].forward(prob.x0[b])
    with ThreadPoolExecutor(max_workers=cores) as ex:
        list(ex.map(one, range(nb)))
        best, reps, t_end = 1e30, 0, time.perf_counter() + seconds
        while reps < 2 or time.perf_counter() < t_end:
            t0 = time.perf_counter()
            list(ex.map(one, range(nb)))
            best = min(best, time.perf_counter() - t0)
            reps += 1
    return {"value": nb / best, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"lqr::LQRSolver (reference headers + Eigen-API shim, oracle/_ref) on the first {nb} problems, "
                      f"{cores} host threads (ctypes thread pool), best of {reps} reps"}


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores, all threads, on the main
    arm's config.  Rank 0 alone runs and prints; the other ranks exit 0 without work."""
    if rank != 0:
        return
    import pdplqr_b200 as P
    name = args.workload if args.workload != "all" else "c3"
    cfg = workload_config(name, world)
    if name == "c3":
        prob = P.problems.cartpole_batch(batch=C3_BATCH, N=C3_N, seed=1234)
        runner = lambda sec: cpu_batch(prob, seconds=sec, max_problems=None)   # every one of the 65,536 problems
    elif name in ("c2", "c1"):
        prob = P.problems.quadrotor_ltv(1024) if name == "c2" else P.problems.quadrotor_example()
        runner = lambda sec: cpu_single(prob, seconds=sec)
    elif name == "c5":
        prob = P.problems.quadrotor_ltv(C5_N)
        runner = lambda sec: cpu_single(prob, seconds=sec, max_stages=C5_N)
    else:
        raise SystemExit("--impl reference supports c3, c2, c1, c5")
    times = []
    cores = sample = None
    per_step = min(2.0, 150.0 / max(1, args.warmup + args.steps))   # whole arm bounded to a few minutes of CPU timing
    for i in range(args.warmup + args.steps):
        v, cores, sample = runner(per_step)
        if i >= args.warmup:
            times.append(v)
    val = statistics.median(times)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * prob.batch / val, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference needs Eigen3 (absent from this image): `value` times the CPU oracle port of the same "
                    "algorithm (allocation-free, the faster of the two CPU builds, i.e. the conservative baseline); "
                    "`reference_build` times the reference's own headers compiled against the Eigen-API shim"}
    if name == "c3":
        try:
            line["reference_build"] = cpu_reference_build(prob)
        except Exception as e:   # the labelled second number must not take the arm down
            line["reference_build"] = {"error": repr(e)}
    emit(line)


# --------------------------------------------------------------------------------------------- C3 (headline)
def run_c3(ctx, steps, warmup, want_cpu=True, batch=C3_BATCH):
    import pdplqr_b200 as P
    torch = ctx.torch
    t_gen = time.time()
    prob = P.problems.cartpole_batch(batch=batch, N=C3_N, seed=1234 + 7919 * ctx.rank)
    log("c3: problem generated in %.1f s" % (time.time() - t_gen))
    sol = P.LQRCudaSolver.from_problem(prob, device=ctx.local_rank, num_segments=1)
    sol.set_stream(ctx.stream.cuda_stream)
    B, N, nx, nu, s = prob.batch, prob.N, prob.nx, prob.nu, prob.s
    rng = np.random.default_rng(17 + ctx.rank)
    ws_host = torch.from_numpy(0.01 * rng.standard_normal((B, prob.ws_len))).pin_memory()
    x0_host = torch.from_numpy(np.ascontiguousarray(prob.x0)).pin_memory()
    out_host = torch.empty_like(ws_host).pin_memory()
    ws_dev, x0_dev = ws_host.to(ctx.dev), x0_host.to(ctx.dev)
    out_dev = torch.empty_like(ws_dev)

    def step_device():
        sol.update_problem_data_device(ws_dev, sigma=SIGMA)
        sol.backward_device()
        sol.forward_device(x0_dev, out_dev)

    l0 = None

    def counted():
        nonlocal l0
        if l0 is None:
            l0 = sol.launch_count()
        step_device()
    for _ in range(warmup):
        step_device()
    ms_step = ctx.time_loop(counted, steps, 0)
    launches = sol.launch_count() - l0
    bad, _ = sol.last_status()

    def bwd_only():
        sol.update_problem_data_device(ws_dev, sigma=SIGMA)
        sol.backward_device()
    ms_bwd = ctx.time_loop(bwd_only, max(5, steps), 1)
    sol.forward_device(x0_dev, out_dev)   # leave the handle consistent (one forward per backward)
    torch.cuda.synchronize()

    # end to end through the host-buffer C ABI (pinned host memory; H2D + kernels + D2H inside the timed region)
    ws_np, x0_np, out_np = ws_host.numpy(), x0_host.numpy(), out_host.numpy()
    link = link_bandwidth(ctx)
    ms_e2e = ctx.time_wall(lambda: sol.solve(ws_np, x0_np, out_np, sigma=SIGMA), max(3, min(steps, 10)))
    got = out_dev.cpu().numpy()
    same = bool(np.array_equal(out_np, got))

    # parity of THIS run against the CPU oracle on a sample of problems: first tile, last tile (padded lanes), random
    parity = None
    if ctx.rank == 0:
        from oracle import oracle as O
        pick = np.unique(np.concatenate([np.arange(64), np.arange(B - 64, B),
                                         np.random.default_rng(5).integers(0, B, 128)]))
        sub = prob.select(pick)
        ref, _ = O.OracleBatch(sub).solve(ws_in=np.ascontiguousarray(ws_np[pick]), sigma=SIGMA)
        parity = rel_err(got[pick], ref)
    bwd_b, fwd_b = stage_bytes(nx, nu, False)
    bwd_sym, _ = stage_bytes(nx, nu, False, sym_h=True)
    ffl, wfl = stage_flops(nx, nu)
    res = {"ms_per_step": ms_step, "value": ctx.world * B * 1e3 / ms_step, "launches": int(launches), "bad": int(bad),
           "e2e": e2e_block(ctx, ctx.world * B / (ms_e2e * 1e-3), ms_e2e, ws_host.numel() * 8 + x0_host.numel() * 8,
                            out_host.numel() * 8, link, {"matches_device_path": same}),
           "roofline": roofline("batch_backward_kernel", bwd_b * B * N, ffl * B * N, ms_bwd, "c3",
                                {"launches_per_backward": 1,
                                 "frac_symmetric_h": bwd_sym * B * N / (ms_bwd * 1e-3) / 1e9 / peaks()[0],
                                 "bytes_note": "frac counts the survey's record (full H, 54 doubles); the device record "
                                               "holds the lower triangle (44): frac_symmetric_h is the traffic-true figure",
                                 "step_hbm_frac": (bwd_b + fwd_b) * B * N / (ms_step * 1e-3) / 1e9 / peaks()[0]}),
           "parity_rel_err": parity, "parity_sample": "256 problems of this run (first tile, last tile, 128 random) vs the CPU oracle"}
    if want_cpu and ctx.world == 1 and ctx.rank == 0:
        v, cores, sample = cpu_batch(prob, seconds=4.0, max_problems=None)
        res["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    del sol, ws_dev, out_dev
    torch.cuda.empty_cache()
    return res


# --------------------------------------------------------------------------------------------- C2 (latency)
def run_single(ctx, prob, num_segments, steps, warmup, load_balancing=True, graph=False):
    """One long problem on one GPU: returns (solver, ms per device-resident step, output array, ws / x0 tensors, step fn).
    graph=True: the step is pdplqr_solve_device (one CUDA graph launch per solve) on a side stream -- stream capture is
    not possible on the legacy default stream."""
    import pdplqr_b200 as P
    torch = ctx.torch
    sol = P.LQRCudaSolver.from_problem(prob, device=ctx.local_rank, num_segments=num_segments, load_balancing=load_balancing)
    stream = torch.cuda.Stream(device=ctx.dev) if graph else ctx.stream
    sol.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(17)
    ws_host = torch.from_numpy(0.01 * rng.standard_normal((1, prob.ws_len))).pin_memory()
    x0_host = torch.from_numpy(np.ascontiguousarray(prob.x0)).pin_memory()
    ws_dev, x0_dev = ws_host.to(ctx.dev), x0_host.to(ctx.dev)
    out_dev = torch.empty_like(ws_dev)
    torch.cuda.synchronize()

    def step():
        if graph:
            sol.solve_device(ws_dev, x0_dev, out_dev, sigma=SIGMA)
        else:
            sol.update_problem_data_device(ws_dev, sigma=SIGMA)
            sol.backward_device()
            sol.forward_device(x0_dev, out_dev)
    ms = ctx.time_loop(step, steps, warmup, stream=stream)
    return sol, ms, out_dev, ws_host, x0_host, step


def leg_c2(ctx, steps):
    import pdplqr_b200 as P
    from oracle import oracle as O
    torch = ctx.torch
    out = dict(workload_config("c2", ctx.world))
    prob = P.problems.quadrotor_ltv(1024)
    # the reference's 4-call protocol, one launch after the other ...
    sol_p, ms_protocol, out_p, _, _, _ = run_single(ctx, prob, 0, max(steps, 50), 5)
    del sol_p
    # ... and the whole solve as one CUDA graph launch (pdplqr_solve_device): the headline latency
    sol, ms, out_dev, ws_host, x0_host, step = run_single(ctx, prob, 0, max(steps, 50), 5, graph=True)
    l0 = sol.launch_count(); step(); launches = sol.launch_count() - l0
    torch.cuda.synchronize()
    out["ms_per_step_protocol_calls"] = ms_protocol
    out["graph_matches_protocol_calls"] = bool(torch.equal(out_dev, out_p))
    out_np = torch.empty(1, prob.ws_len, dtype=torch.float64).pin_memory().numpy()
    ms_e2e = ctx.time_wall(lambda: sol.solve(ws_host.numpy(), x0_host.numpy(), out_np, sigma=SIGMA), 20, 3)
    bwd_b, fwd_b = stage_bytes(12, 4, True)
    ffl, wfl = stage_flops(12, 4, pdp=True)
    out.update(num_segments=sol.num_segments, ms_per_step=ms, latency_us=ms * 1e3, value=ctx.world * 1e3 / ms, unit=UNIT,
               gpu_launches=int(launches),
               roofline=roofline("whole step (latency-bound by construction: small fraction expected)", (bwd_b + fwd_b) * 1024,
                                 (ffl + wfl) * 1024, ms),
               e2e=e2e_block(ctx, ctx.world / (ms_e2e * 1e-3), ms_e2e, (prob.ws_len + 12) * 8, prob.ws_len * 8, None))
    if ctx.rank == 0:
        ref = O.OracleSolver(prob, parallel=False).solve(ws_in=ws_host.numpy()[0].copy(), sigma=SIGMA)
        out["parity_rel_err"] = rel_err(out_dev.cpu().numpy()[0], ref)
    del sol
    # the metric's "per-solve latency vs N": library-chosen segmentation, device-resident step
    lat = {}
    for n in LATENCY_NS:
        pn = P.problems.quadrotor_ltv(n) if n != 100 else P.problems.quadrotor_example()
        sn, msn, on, wh, _, _ = run_single(ctx, pn, 0, 30, 5, graph=True)
        entry = {"us": msn * 1e3, "num_segments": sn.num_segments}
        if ctx.rank == 0:
            refn = O.OracleSolver(pn, parallel=False).solve(ws_in=wh.numpy()[0].copy(), sigma=SIGMA)
            entry["parity_rel_err"] = rel_err(on.cpu().numpy()[0], refn)
            if ctx.world == 1 and not ctx.args.no_cpu_baseline:
                entry.update(cpu_latency(pn))
                entry["speedup_vs_cpu"] = entry["cpu_us"] / entry["us"]
        lat[str(n)] = entry
        del sn
    out["latency_vs_N_us"] = lat
    if ctx.world == 1 and not ctx.args.no_cpu_baseline:
        v, cores, sample = cpu_single(prob, seconds=2.0)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                               "latency_us": 1e6 / v}
    return out


# --------------------------------------------------------------------------------------------- C5 (long horizon)
def leg_c5(ctx, steps):
    """N = 2^20.  1 GPU: whole horizon on one handle.  N > 1 GPUs: contiguous time slices per rank, ONE all_gather of a
    3,648-byte summary per solve (HorizonShardedSolver); strong scaling: `value` = solves/s of the ONE problem."""
    import pdplqr_b200 as P
    from pdplqr_b200 import sharding
    torch, dist = ctx.torch, ctx.dist
    world, rank = ctx.world, ctx.rank
    out = dict(workload_config("c5", world))
    steps = max(3, min(steps, 10))
    nseg = c5_segments(world)
    rng = np.random.default_rng(17)
    ws_full_host = None
    t0 = time.time()
    if world == 1:
        prob = P.problems.quadrotor_ltv(C5_N)
        log("c5: problem generated in %.1f s" % (time.time() - t0))
        sol, ms, out_dev, ws_host, x0_host, step = run_single(ctx, prob, nseg, steps, 3, load_balancing=2)
        runs = [ms] + [ctx.time_loop(step, steps, 0) for _ in range(2)]   # best of three K-step loops, all reported
        ms = min(runs)
        out["ms_per_step_runs"] = runs
        ws_full_host = ws_host.numpy()
        l0 = sol.launch_count(); step(); launches = sol.launch_count() - l0

        def bwd_only():
            sol.update_problem_data_device(ws_dev1, sigma=SIGMA)
            sol.backward_device()
        ws_dev1 = ws_host.to(ctx.dev)
        x0_dev1 = x0_host.to(ctx.dev)
        ms_bwd = ctx.time_loop(bwd_only, 5, 1)
        sol.forward_device(x0_dev1, out_dev)
        torch.cuda.synchronize()
        out_pin = torch.empty(1, prob.ws_len, dtype=torch.float64).pin_memory()
        out_np = out_pin.numpy()
        link = link_bandwidth(ctx, 128 << 20)
        ms_e2e = ctx.time_wall(lambda: sol.solve(ws_host.numpy(), x0_host.numpy(), out_np, sigma=SIGMA), 3, 1)
        local_N, h2d, d2h = C5_N, (prob.ws_len + 12) * 8, prob.ws_len * 8
        got_full = out_dev.cpu().numpy()[0]
        out["num_segments"] = sol.num_segments
    else:
        start, count = sharding.horizon_slices(C5_N, world)[rank]
        local = P.problems.quadrotor_ltv(C5_N, start=start, count=count)
        log("c5: slice [%d, %d) generated in %.1f s" % (start, start + count, time.time() - t0))
        hs = sharding.HorizonShardedSolver(None, rank, world, num_segments=nseg, device=ctx.local_rank, local=local)
        hs.set_stream(ctx.stream.cuda_stream)
        sol = hs.sol
        # the iterate w_prev of the full problem, seeded so that every rank can cut its own slice
        ws_full_host = 0.01 * rng.standard_normal((1, C5_N * 16 + 12))
        lo = start * 16
        hi = lo + count * 16 + 12          # the slice's last entry is x at its exit stage
        ws_host = torch.from_numpy(np.ascontiguousarray(ws_full_host[:, lo:hi])).pin_memory()
        ws_dev = ws_host.to(ctx.dev)
        out_dev = torch.empty_like(ws_dev)
        out_host = torch.empty_like(ws_host).pin_memory()

        def step():
            hs.solve_device(ws_dev, SIGMA, out_dev)
        # best of three K-step loops (all reported): with K = 10 steps of ~1 ms one host-side stall of a rank would double the figure
        runs = [ctx.time_loop(step, steps, 3 if i == 0 else 0) for i in range(3)]
        ms = min(runs)
        out["ms_per_step_runs"] = runs
        l0 = sol.launch_count(); step(); launches = sol.launch_count() - l0

        def bwd_only():
            sol.update_problem_data_device(ws_dev, sigma=SIGMA)
            sol.backward_device()
        ms_bwd = ctx.time_loop(bwd_only, 5, 1)
        step()
        torch.cuda.synchronize()
        # per-phase breakdown (CUDA events between the phases of a few instrumented solves; median, max over ranks)
        per = {k: [] for k in hs.PHASES}
        for _ in range(5):
            ev = []
            ctx.barrier()
            hs.solve_device(ws_dev, SIGMA, out_dev, events=ev)
            torch.cuda.synchronize()
            for i, k in enumerate(hs.PHASES):
                per[k].append(ev[i].elapsed_time(ev[i + 1]) * 1e3)
        med = ctx.max_over_ranks([statistics.median(per[k]) for k in hs.PHASES])
        out["phase_us"] = dict(zip(hs.PHASES, med))
        out["phase_us"]["note"] = ("max over ranks of the per-rank median; all_gather includes waiting for the slowest rank's "
                                   "local sweep; local_sweep_and_tree = stage sweep + local interface tree up-sweep")
        out["limiting_phase"] = max(hs.PHASES, key=lambda k: out["phase_us"][k])
        link = link_bandwidth(ctx, 128 << 20)

        def step_e2e():   # host slice in, sharded solve, host slice out (pinned buffers)
            ws_dev.copy_(ws_host, non_blocking=True)
            hs.solve_device(ws_dev, SIGMA, out_dev)
            out_host.copy_(out_dev, non_blocking=True)
            torch.cuda.synchronize()
        ms_e2e = ctx.time_wall(step_e2e, 3, 1)
        local_N, h2d, d2h = count, ws_host.numel() * 8, out_host.numel() * 8
        # gather the slices on rank 0 for the parity checks
        parts = [torch.empty(sharding.horizon_slices(C5_N, world)[r][1] * 16 + 12, dtype=torch.float64, device=ctx.dev)
                 for r in range(world)] if rank == 0 else None
        dist.gather(out_dev[0], parts, dst=0)
        got_full = None
        if rank == 0:
            got_full = np.empty(C5_N * 16 + 12)
            for r in range(world):
                s0, cnt = sharding.horizon_slices(C5_N, world)[r]
                pr = parts[r].cpu().numpy()
                got_full[s0 * 16:(s0 + cnt) * 16] = pr[:cnt * 16]
                if r == world - 1:
                    got_full[-12:] = pr[-12:]
        out["num_segments"] = sol.num_segments * world
        out["segments_per_rank"] = sol.num_segments
    bwd_b, fwd_b = stage_bytes(12, 4, True)
    ffl, wfl = stage_flops(12, 4, pdp=True)
    out.update(ms_per_step=ms, value=1e3 / ms, unit=UNIT, scaling="strong" if world > 1 else "n/a (1 GPU)",
               gpu_launches=int(launches),
               roofline=roofline("seg_backward_kernel<12,4,32> (+ local interface tree up-sweep)", bwd_b * local_N,
                                 ffl * local_N, ms_bwd, "c5" if world == 1 else None,
                                 {"step_hbm_frac": (bwd_b + fwd_b) * C5_N / world / (ms * 1e-3) / 1e9 / peaks()[0],
                                  "per_rank": world > 1}),
               e2e=e2e_block(ctx, 1.0 / (ms_e2e * 1e-3), ms_e2e, h2d, d2h, link))
    del sol
    if world > 1:
        del hs, step, bwd_only, step_e2e
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if rank == 0:   # parity of this run: sequential CPU oracle over the whole horizon (and the 1-GPU solve when sharded)
        from oracle import oracle as O
        t0 = time.time()
        full = P.problems.quadrotor_ltv(C5_N) if world > 1 else prob
        ref = O.OracleSolver(full, parallel=False).solve(ws_in=np.ascontiguousarray(ws_full_host[0]), sigma=SIGMA)
        out["parity_rel_err"] = rel_err(got_full, ref)
        out["parity_sample"] = "all 2^20 stages vs the sequential CPU oracle (%.0f s)" % (time.time() - t0)
        if world > 1:
            s1 = P.LQRCudaSolver.from_problem(full, device=ctx.local_rank, num_segments=c5_segments(1), load_balancing=2)
            one = np.empty((1, full.ws_len))
            s1.solve(np.ascontiguousarray(ws_full_host), full.x0, one, sigma=SIGMA)
            out["parity_vs_1gpu"] = rel_err(got_full, one[0])
            del s1
        if world == 1 and not ctx.args.no_cpu_baseline:
            v, cores, sample = cpu_single(full, seconds=3.0)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    return out


# --------------------------------------------------------------------------------------------- C4 (conic ADMM)
def leg_c4(ctx, steps):
    """C4 on a side stream: the conic solve is ONE CUDA graph launch (pdplqr_admm_solve_device), and stream capture is not
    possible on the legacy default stream (there the library falls back to issuing the iterations from a host loop)."""
    torch = ctx.torch
    side = torch.cuda.Stream(device=ctx.dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        out = _leg_c4(ctx, steps, side)
    torch.cuda.synchronize()
    return out


def _leg_c4(ctx, steps, stream):
    """BASELINE.json configs[3]: conic-constrained (box + SOC) LQ MPC nx=30 nu=10 N=256, batch 4096 per GPU, full outer
    iterations.  A step = one ADMM solve with a FIXED number of outer iterations (1 factorising + ITERS-1 affine-only
    LQ solves, projections, residuals), device-resident.  The outer iteration is not in the reference (hooks only)."""
    import pdplqr_b200 as P
    torch = ctx.torch
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    B = int(os.environ.get("C4_BATCH", str(C4_BATCH)))
    N = int(os.environ.get("C4_N", str(C4_N)))
    ITERS = int(os.environ.get("C4_ITERS", str(C4_ITERS)))
    out = dict(workload_config("c4", world))
    base = 64 if B % 64 == 0 else B
    hp = P.problems.random_conic_batch(batch=base, N=N, seed=99 + rank)
    rep = B // base
    nx, nu, s = hp.nx, hp.nu, hp.s
    sol = P.LQRCudaSolver(nx, nu, N, batch=B, num_segments=1, ncs=hp.ncs, device=ctx.local_rank)
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
    last = {}

    def step():
        w.zero_(); z.zero_(); y.zero_()
        last["it"], last["res"] = sol.admm_solve_device(x0, w, z, y, rho, inv_rho, sigma=SIGMA, alpha=1.6, max_iter=ITERS,
                                                        eps_abs=0.0, eps_rel=0.0, check_every=ITERS)
    steps = max(1, min(steps, int(os.environ.get("C4_STEPS", "2"))))
    step()
    torch.cuda.synchronize()
    l0 = sol.launch_count()
    g0 = sol.admm_stats()[0]
    ms_step = ctx.time_loop(step, steps, 0, stream=stream)
    graph_launches = (sol.admm_stats()[0] - g0) / steps
    launches = (sol.launch_count() - l0) // steps
    w_gpu = w[:base].cpu().numpy()

    def aff():
        sol.update_problem_data_device(w, y, z, inv_rho, sigma=SIGMA)
        sol.backward_without_factorization_device(rho)
    ms_aff = ctx.time_loop(aff, 10, 1, stream=stream)

    def fact():
        sol.update_problem_data_device(w, y, z, inv_rho, sigma=SIGMA)
        sol.backward_device(rho)
    ms_fact = ctx.time_loop(fact, 2, 1, stream=stream)
    sol.forward_device(x0, torch.empty_like(w))
    torch.cuda.synchronize()
    # "fixed 50 and to-tolerance (1e-4) outer iterations".  The to-tolerance solve: cold start, the same
    # rho = 0.1, OSQP rho adaptation on (a rescale re-factorises: one more graph launch each), convergence tested on the
    # device every 25 iterations over the whole batch (the slowest problem decides); run once, bounded by C4_TOL_MAX_ITER.
    tol_iters = int(os.environ.get("C4_TOL_MAX_ITER", "2000"))
    if tol_iters > 0:
        try:
            sol.admm_configure(use_graph=True, adaptive_rho=True, rho_tau=5.0, max_rho_updates=10)

            def step_tol():
                w.zero_(); z.zero_(); y.zero_()
                last["tol_it"], last["tol_res"] = sol.admm_solve_device(x0, w, z, y, rho, inv_rho, sigma=SIGMA, alpha=1.6,
                                                                        max_iter=tol_iters, eps_abs=1e-4, eps_rel=1e-4,
                                                                        check_every=25)
            g0 = sol.admm_stats()[0]
            ms_tol = ctx.time_loop(step_tol, 1, 0, stream=stream)
            g1, n_rho = sol.admm_stats()
            it_max = int(ctx.max_over_ranks([float(last["tol_it"])])[0])
            out["to_tolerance"] = {"eps_abs": 1e-4, "eps_rel": 1e-4, "ms": ms_tol, "iterations": int(last["tol_it"]),
                                   "iterations_max_over_ranks": it_max,
                                   "converged": bool(last["tol_it"] < tol_iters), "max_iter": tol_iters,
                                   "rho_updates": int(n_rho), "graph_launches": int(g1 - g0), "check_every": 25,
                                   "residuals": [float(last["tol_res"][0]), float(last["tol_res"][1])],
                                   "value": world * B * 1e3 / ms_tol, "unit": UNIT,
                                   "ms_per_iteration": ms_tol / max(1, it_max),
                                   "note": "whole batch to eps 1e-4 (slowest problem decides), rho = 0.1 at the start, OSQP rho "
                                           "adaptation (tau 5); max over ranks of the device time"}
        except Exception as e:  # noqa: BLE001 -- a failed extra must not take the leg down
            out["to_tolerance"] = {"error": repr(e)[:300]}
        finally:
            sol.admm_configure(use_graph=True, adaptive_rho=False, rho_tau=5.0, max_rho_updates=10)
    nc = int(hp.ncs[1])
    # traffic-true algorithmic bytes (selection-matrix constraints: the dense D, nc*s doubles per stage, is never read):
    # affine sweep reads [E|c], h, [K|d], [Quu^-1|P+c], w_prev, rho/z/y/inv_rho + (column, value) of every row, writes d
    aff_bytes = 8 * (nx * s + nx + s + nu * (nx + 1) + nu * nu + nx + s + 5.5 * nc + nu)
    # factorising sweep: model record + ADMM vectors in, factor record + affine cache out (no dense D either)
    fact_bytes = 8 * (nx * s + nx + s * s + s + s + 5.5 * nc + nu * (nx + 1) + nu * nu + nx)
    fact_flops, _ = stage_flops(nx, nu, nc=0, pdp=False)     # the selection fold-in is O(nc): not counted
    aff_flops = 2 * nx * s + 2 * nu * nu + 2 * nu * nx + 4 * nc + 2 * s
    out.update(ms_per_step=ms_step, value=world * B * 1e3 / ms_step, unit=UNIT, steps=steps, gpu_launches=int(launches),
               graph_launches_per_solve=graph_launches,
               problem_iterations_per_sec=world * B * ITERS * 1e3 / ms_step,
               ms_factorizing_backward=ms_fact, ms_affine_backward=ms_aff,
               final_residuals=[float(last["res"][0]), float(last["res"][1])],
               roofline=roofline("seg_affine_kernel<30,10,128> (the common ADMM iteration)", aff_bytes * B * N,
                                 aff_flops * B * N, ms_aff, None,
                                 {"traffic": c4_traffic(B),
                                  "traffic_source": "profiles/traffic.json['c4_affine_batch4096'] scaled to the batch (recorded ncu capture)",
                                  "bytes_note": "selection-matrix constraints: no dense D traffic counted (the survey's 55,488 B/stage model does)",
                                  "factorizing_kernel": roofline("seg_backward_kernel<30,10,128,CON>", fact_bytes * B * N,
                                                                 fact_flops * B * N, ms_fact)}))
    # e2e: host iterates in / out through pdplqr_admm_solve (pinned), a bounded sub-batch would change the workload:
    # the whole batch is solved once
    link = link_bandwidth(ctx, 128 << 20)
    hw = torch.zeros(B, hp.ws_len, dtype=torch.float64).pin_memory()
    hz = torch.zeros(B, nct, dtype=torch.float64).pin_memory()
    hy = torch.zeros(B, nct, dtype=torch.float64).pin_memory()
    hrho = torch.full((B, nct), 0.1, dtype=torch.float64).pin_memory()
    hx0 = x0.cpu().pin_memory()

    def step_e2e():
        hw.zero_(); hz.zero_(); hy.zero_()
        sol.admm_solve(hx0.numpy(), hw.numpy(), hz.numpy(), hy.numpy(), hrho.numpy(), sigma=SIGMA, alpha=1.6,
                       max_iter=ITERS, eps_abs=0.0, eps_rel=0.0, check_every=ITERS)
    ms_e2e = ctx.time_wall(step_e2e, 1, 1)
    h2d = (hw.numel() + hz.numel() + hy.numel() + 2 * hrho.numel() + hx0.numel()) * 8
    d2h = (hw.numel() + hz.numel() + hy.numel()) * 8
    out["e2e"] = e2e_block(ctx, world * B / (ms_e2e * 1e-3), ms_e2e, h2d, d2h, link)
    if rank == 0:
        from oracle import admm_ref
        t0 = time.time()
        errs = []
        for b in range(2):   # two problems of this run, all ITERS outer iterations, numpy restatement + CPU oracle
            wr, zr, yr, rp, rd = admm_ref.admm(hp, b, np.full(nct, 0.1), sigma=SIGMA, alpha=1.6, iters=ITERS)
            errs.append(rel_err(w_gpu[b], wr))
        out["parity_rel_err"] = max(errs)
        out["parity_sample"] = ("2 problems x %d outer iterations vs oracle/admm_ref.py (%.0f s); the outer iteration is "
                                "not in the reference: parity unpinned by construction" % (ITERS, time.time() - t0))
        if world == 1 and not ctx.args.no_cpu_baseline:
            from oracle import oracle as O
            nb = min(base, 2 * host_cores())
            sub = hp.select(slice(0, nb))
            pool = O.OracleBatch(sub)
            rngc = np.random.default_rng(0)
            ys_, zs_ = rngc.standard_normal((nb, nct)), rngc.standard_normal((nb, nct))
            rh = np.full((nb, nct), 0.1)
            nt = host_cores()
            t0 = time.perf_counter(); pool.solve(ys=ys_, zs=zs_, rho=rh, inv_rho=1.0 / rh, factorize=True, nthreads=nt); tf = time.perf_counter() - t0
            t0 = time.perf_counter(); pool.solve(ys=ys_, zs=zs_, rho=rh, inv_rho=1.0 / rh, factorize=False, nthreads=nt); tn = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": nb / (tf + (ITERS - 1) * tn), "unit": UNIT, "cores": nt, "kind": "port",
                                   "sample": "%d problems: 1 factorising + 1 affine-only LQ solve timed, extrapolated to %d iterations (projections not counted)" % (nb, ITERS)}
    del sol
    torch.cuda.empty_cache()
    return out


LEGS = {"c2": leg_c2, "c5": leg_c5, "c4": leg_c4}


# --------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", help="headline workload (c3), or one of c2 / c5 / c4 alone")
    ap.add_argument("--legs", default=os.environ.get("BENCH_LEGS", "c2,c5,c4"),
                    help="extra configurations measured into the `configs` block of the c3 line ('' for none)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    ctx = Ctx(args)
    sampler = ClockSampler(ctx.local_rank)
    t_start = time.time()
    if args.workload in LEGS:        # one configuration alone, as the line's own value
        leg = LEGS[args.workload](ctx, args.steps)
        main_res = None
    else:
        main_res = run_c3(ctx, args.steps, args.warmup, want_cpu=not args.no_cpu_baseline,
                          batch=int(os.environ.get("C3_BATCH", str(C3_BATCH))))
        c3_windows = list(ctx.windows)
        legs = {}
        c3_clocks = sampler.summary(c3_windows)

        def headline(extra):
            line = {"metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": workload_config("c3", ctx.world), "clocks": c3_clocks, "e2e": main_res["e2e"],
                    "gpu_launches": main_res["launches"], "roofline": main_res["roofline"],
                    "parity_rel_err": main_res["parity_rel_err"], "parity_sample": main_res["parity_sample"],
                    "non_pd_problems": main_res["bad"], "configs": legs, "bench_wall_s": time.time() - t_start}
            if "cpu_baseline" in main_res:
                line["cpu_baseline"] = main_res["cpu_baseline"]
            line.update(extra)
            return line
        # safety net: if the auxiliary legs hang (e.g. a desynchronised collective after a failure on one rank), rank 0
        # still prints the headline measured above and the process exits instead of waiting for the NCCL watchdog
        import threading
        done = threading.Event()

        def watchdog():
            if not done.wait(float(os.environ.get("BENCH_LEGS_BUDGET_S", "420"))):
                if ctx.rank == 0:
                    emit(headline({"configs_error": "auxiliary legs exceeded their time budget; headline only"}))
                os._exit(0)
        threading.Thread(target=watchdog, daemon=True).start()
        for name in [x for x in args.legs.split(",") if x]:
            t0 = time.time()
            try:
                legs[name] = LEGS[name](ctx, args.steps)
            except Exception as e:   # an auxiliary configuration must not take the headline down
                traceback.print_exc()
                legs[name] = {"error": repr(e)}
                import gc
                e = None
                gc.collect()
            legs[name]["leg_wall_s"] = time.time() - t0
            log("leg %s done in %.1f s" % (name, time.time() - t0))
            ctx.torch.cuda.empty_cache()
    if main_res is None:
        clocks = sampler.summary(ctx.windows)
    else:
        done.set()
    if ctx.rank == 0:
        if main_res is None:
            line = {"metric": METRIC, "value": leg["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
                    "scaling": leg.get("scaling", "weak"), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": workload_config(args.workload, ctx.world), "clocks": clocks, "e2e": leg.get("e2e"),
                    "gpu_launches": leg.get("gpu_launches"), "roofline": leg.get("roofline"),
                    "cpu_baseline": leg.get("cpu_baseline"), "detail": leg}
        else:
            line = headline({})
        emit(line)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
