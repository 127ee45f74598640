"""One long-horizon solve (BASELINE.json config 5 shape) for ncu captures: N = 2^20 (or argv[1]), bench segmentation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pdplqr_b200 as P

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
wave = P.wave_size(12, 4)
S = max(1, (max(N // 450, wave) // wave) * wave)
p = P.problems.quadrotor_ltv(N)
sol = P.LQRCudaSolver.from_problem(p, num_segments=S, load_balancing=2)
ws = 0.01 * np.random.default_rng(17).standard_normal((1, p.ws_len))
out = np.zeros_like(ws)
for _ in range(3):
    sol.solve(ws, p.x0, out, sigma=1e-6)
print("segments", sol.num_segments, "wave", wave, "launches", sol.launch_count(), "checksum", float(np.abs(out).sum()))
