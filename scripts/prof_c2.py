"""C2 (quadrotor N = 1024, one problem) solves for ncu launch lists (warm per-kernel durations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pdplqr_b200 as P
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
p = P.problems.quadrotor_ltv(N)
sol = P.LQRCudaSolver.from_problem(p, num_segments=0)
dev = torch.device("cuda", 0)
ws = torch.zeros(1, p.ws_len, dtype=torch.float64, device=dev); out = torch.zeros_like(ws); x0 = torch.from_numpy(p.x0).to(dev)
sol.set_stream(torch.cuda.current_stream().cuda_stream)
for _ in range(6):
    sol.update_problem_data_device(ws, sigma=1e-6); sol.backward_device(); sol.forward_device(x0, out)
torch.cuda.synchronize()
print("segments", sol.num_segments, "launches", sol.launch_count())
