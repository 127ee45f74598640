"""Multi-rank host logic on CPU (gloo, world_size 2): batch / horizon partition, the summary all_gather and the
interface solve that couples the ranks' time slices.  The per-slice numbers come from the oracle (no GPU here);
the GPU version of the same flow is tests/test_parity_gpu.py::test_horizon_shards_on_one_gpu."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import pdplqr_b200 as P
from pdplqr_b200 import sharding
from conftest import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slices():
    assert sharding.batch_slices(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert sharding.horizon_slices(1 << 20, 8)[-1] == (7 * (1 << 17), 1 << 17)
    assert sum(n for _, n in sharding.batch_slices(65536, 8)) == 65536
    with pytest.raises(ValueError):
        sharding.horizon_slices(3, 4)


def _pack_summary(o, seg):
    Pm, p, F, f, C = o.summary(seg)
    return np.concatenate([Pm, F, C, p, f])


def test_couple_numpy_matches_oracle_interface(oracle):
    p = P.problems.quadrotor_ltv(96)
    G = 4
    o = oracle.OracleSolver(p, parallel=True, num_segments=G, load_balancing=False, condensed=oracle.LU)
    ws = o.solve()
    xo, uo = o.interface()
    S = np.stack([_pack_summary(o, i) for i in range(G)])
    S[-1, 144:3 * 144] = 0.0   # the terminal slice is a pure value function (F = C = 0)
    S[-1, 3 * 144 + 12:] = 0.0
    xh, lam = sharding.couple_numpy(S, p.x0[0])
    assert rel_err(xh, xo) < 1e-11
    assert np.max(np.abs(lam[:G - 1] - uo[:G - 1])) < 1e-9
    assert ws is not None


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import torch.distributed as dist
    import pdplqr_b200 as P2
    from pdplqr_b200 import sharding as sh
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prob = P2.problems.quadrotor_ltv(64)
        # every rank reduces ITS time slice to a summary (here with the oracle), then one all_gather
        o = O.OracleSolver(prob, parallel=True, num_segments=world, load_balancing=False, condensed=O.LU)
        o.update_problem_data(np.zeros(prob.ws_len), sigma=1e-6)
        o.backward()
        Pm, p, F, f, C = o.summary(rank)
        if rank == world - 1:
            F[:], f[:], C[:] = 0.0, 0.0, 0.0
        mine = torch.from_numpy(np.concatenate([Pm, F, C, p, f]))
        allsum = sh.all_gather_rows(mine, world).numpy()
        xh, lam = sh.couple_numpy(allsum, prob.x0[0])
        o.forward(prob.x0[0], np.zeros(prob.ws_len))   # the oracle computes xhat/uhat in forward
        xo, uo = o.interface()
        ok = bool(np.max(np.abs(xh - xo)) < 1e-10 and np.max(np.abs(lam[:world - 1] - uo[:world - 1])) < 1e-9)
        start, count = sh.horizon_slices(prob.N, world)[rank]
        st, ln = o.partition()
        ok = ok and (start, count) == (int(st[rank]), int(ln[rank]))
        # batch sharding: every rank solves its slice, no collective; gather only to check coverage
        bs = sh.batch_slices(10, world)[rank]
        cover = sh.all_gather_rows(torch.tensor([bs[0], bs[1]], dtype=torch.int64), world).numpy()
        ok = ok and int(cover[:, 1].sum()) == 10 and int(cover[-1, 0] + cover[-1, 1]) == 10
        q.put((rank, ok))
    except Exception as e:  # report instead of hanging the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_horizon_and_batch_sharding(oracle):
    world, port = 2, 29631
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=90) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_bench_wave_alignment_helpers():
    """bench.py rounds segment counts to whole waves of resident stage-kernel CTAs and scales measured traffic."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    W = bench.WAVE
    assert bench.wave_aligned(5) == 5 and bench.wave_aligned(W) == W
    assert bench.wave_aligned(W + 1) == W and bench.wave_aligned(3 * W - 1) == 2 * W
    assert bench.wave_aligned((1 << 20) // 250) == 2 * W
    t = bench.c4_traffic(4096)
    assert t is None or abs(bench.c4_traffic(2048) - t / 2) < 1.0


def test_bench_cpu_latency_leg(oracle):
    """bench.py's per-N CPU latency (the number printed beside every entry of latency_vs_N_us): the faster of the sequential
    LQRSolver port and the PDP port on 2 / 4 / 8 threads, with the sequential time kept next to it."""
    import importlib.util
    import os
    import pdplqr_b200 as P
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    r = bench.cpu_latency(P.problems.quadrotor_example(), seconds=0.2)
    assert r["cpu_us"] > 0 and r["cpu_us"] <= r["cpu_sequential_us"] and r["cpu_threads"] in (1, 2, 4, 8)
    assert 10 < r["cpu_sequential_us"] < 1e6      # a 100-stage nx12/nu4 solve is tens to hundreds of microseconds on any host
