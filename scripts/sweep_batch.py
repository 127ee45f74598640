"""Tuning sweep for the thread-per-problem kernels (C3): launch variants x {backward, forward}, CUDA-event timed,
each checked against the oracle on a few problems.  Run on the GPU box: python scripts/sweep_batch.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pdplqr_b200 as P
from oracle import oracle as O

batch = int(os.environ.get("SWEEP_BATCH", "65536"))
prob = P.problems.cartpole_batch(batch=batch, N=128)
dev = torch.device("cuda", 0)
rng = np.random.default_rng(17)
ws_dev = torch.from_numpy(0.01 * rng.standard_normal((batch, prob.ws_len))).to(dev)
x0_dev = torch.from_numpy(prob.x0).to(dev)
out_dev = torch.empty_like(ws_dev)
ref = {}
for b in (0, batch // 3, batch - 1):
    ref[b] = O.OracleSolver(prob, b=b).solve(ws_in=ws_dev[b].cpu().numpy(), sigma=1e-6)
stream = torch.cuda.current_stream()


def time_it(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


variants = [int(v) for v in os.environ.get("SWEEP_VARIANTS", "0,1,2,3,4").split(",")]
for v in variants:
    os.environ["PDPLQR_BWD_VARIANT"] = str(v)
    os.environ["PDPLQR_FWD_VARIANT"] = str(v)
    sol = P.LQRCudaSolver.from_problem(prob)
    sol.set_stream(stream.cuda_stream)

    def bwd():
        sol.update_problem_data_device(ws_dev, sigma=1e-6)
        sol.backward_device()

    def fwd():
        sol._lib.pdplqr_update_problem_data_device  # noqa
        sol.update_problem_data_device(ws_dev, sigma=1e-6)
        sol.backward_device()
        sol.forward_device(x0_dev, out_dev)

    tb = time_it(bwd)
    tall = time_it(fwd)
    torch.cuda.synchronize()
    o = out_dev.cpu().numpy()
    err = max(np.max(np.abs(o[b] - r)) / np.max(np.abs(r)) for b, r in ref.items())
    print(f"variant {v}: backward {tb*1e3:8.1f} us   step {tall*1e3:8.1f} us   forward ~{(tall-tb)*1e3:8.1f} us   rel err {err:.1e}", flush=True)
    sol.close()
