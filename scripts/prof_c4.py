"""Short conic ADMM run (C4 shape, reduced batch / iterations) for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pdplqr_b200 as P
B, N, ITERS = int(os.environ.get("C4_BATCH", "4096")), 256, int(os.environ.get("C4_ITERS", "6"))
hp = P.problems.random_conic_batch(batch=64, N=N, seed=99)
rep = B // 64
dev = torch.device("cuda", 0)
sol = P.LQRCudaSolver(hp.nx, hp.nu, N, batch=B, num_segments=1, ncs=hp.ncs)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev).repeat(rep, *([1] * (a.ndim - 1))).contiguous()
sol.set_model_device(*[up(a) for a in (hp.E, hp.c, hp.H, hp.h, hp.HN, hp.hN, hp.D)])
lb = np.where(np.isfinite(hp.e_lb), hp.e_lb, -1e20); ub = np.where(np.isfinite(hp.e_ub), hp.e_ub, 1e20)
sol.admm_set_cones(hp.cones, np.tile(lb, (rep, 1)), np.tile(ub, (rep, 1)))
nct = hp.nc_total
x0 = up(hp.x0); rho = torch.full((B, nct), 0.1, dtype=torch.float64, device=dev); inv = 1.0 / rho
w = torch.zeros(B, hp.ws_len, dtype=torch.float64, device=dev); z = torch.zeros(B, nct, dtype=torch.float64, device=dev); y = torch.zeros_like(z)
for _ in range(2):
    w.zero_(); z.zero_(); y.zero_()
    it, res = sol.admm_solve_device(x0, w, z, y, rho, inv, sigma=1e-6, alpha=1.6, max_iter=ITERS, eps_abs=0.0, eps_rel=0.0, check_every=ITERS)
torch.cuda.synchronize()
print("iters", it, "res", res, "launches", sol.launch_count())
