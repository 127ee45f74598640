"""clock64() phase breakdown of the stage kernel (instrumented build: PDPLQR_VARIANT=prof
PDPLQR_CFLAGS=-DPDPLQR_PHASE_CLOCKS).  Prints cycles per stage and phase for the C5 throughput configuration (one warp
per segment, whole waves) and the C2 latency configuration (128 threads per segment)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pdplqr_b200 as P

for name, N, S, lb in (("c5-like (2 waves of 14 CTAs/SM)", 1 << 18, 4144, 2), ("c2 (latency mode)", 1024, 128, 2)):
    p = P.problems.quadrotor_ltv(N)
    sol = P.LQRCudaSolver.from_problem(p, num_segments=S, load_balancing=lb)
    ws = p.zeros_ws()
    print("==", name, "segments", sol.num_segments, flush=True)
    for _ in range(2):
        sol.update_problem_data(ws, sigma=1e-6)
        sol.backward()
        sol.forward(p.x0, np.zeros_like(ws))
    sol.synchronize()
    del sol
