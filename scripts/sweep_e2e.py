"""e2e (host-buffer pdplqr_solve) time vs number of pipeline chunks, C3 workload."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pdplqr_b200 as P
prob = P.problems.cartpole_batch(batch=65536, N=128)
rng = np.random.default_rng(17)
ws = torch.from_numpy(0.01 * rng.standard_normal((prob.batch, prob.ws_len))).pin_memory()
x0 = torch.from_numpy(np.ascontiguousarray(prob.x0)).pin_memory()
out = torch.empty_like(ws).pin_memory()
for chunks in [int(c) for c in os.environ.get("SWEEP_CHUNKS", "1,4,8,16,32").split(",")]:
    os.environ["PDPLQR_PIPELINE_CHUNKS"] = str(chunks)
    sol = P.LQRCudaSolver.from_problem(prob)
    for _ in range(2):
        sol.solve(ws.numpy(), x0.numpy(), out.numpy(), sigma=1e-6)
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        sol.solve(ws.numpy(), x0.numpy(), out.numpy(), sigma=1e-6)
    dt = (time.perf_counter() - t0) / n
    print(f"chunks={chunks:3d}  e2e {dt*1e3:7.2f} ms/step  {prob.batch/dt/1e6:6.2f} M solves/s", flush=True)
    sol.close()
