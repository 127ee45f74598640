"""C2 latency sweep: single quadrotor problem N=1024, step time vs number of segments (CUDA events)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pdplqr_b200 as P

N = int(os.environ.get("SWEEP_N", "1024"))
prob = P.problems.quadrotor_ltv(N)
dev = torch.device("cuda", 0)
ws_dev = torch.zeros(1, prob.ws_len, dtype=torch.float64, device=dev)
x0_dev = torch.from_numpy(prob.x0).to(dev)
out_dev = torch.empty_like(ws_dev)
stream = torch.cuda.current_stream()
for S in [int(v) for v in os.environ.get("SWEEP_S", "16,32,64,128,256,512").split(",")]:
    sol = P.LQRCudaSolver.from_problem(prob, num_segments=S, load_balancing=int(os.environ.get("SWEEP_LB", "0")))
    sol.set_stream(stream.cuda_stream)

    def step():
        sol.update_problem_data_device(ws_dev, sigma=1e-6)
        sol.backward_device()
        sol.forward_device(x0_dev, out_dev)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"S={S:5d}  step {e0.elapsed_time(e1)/n*1e3:8.1f} us   launches/step {sol.launch_count()//55}", flush=True)
    sol.close()
