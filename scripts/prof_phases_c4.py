"""clock64() phase breakdown of seg_backward_kernel<30,10,128,CON> at the C4 shape (instrumented build: PDPLQR_VARIANT=prof
PDPLQR_CFLAGS=-DPDPLQR_PHASE_CLOCKS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pdplqr_b200 as P
hp = P.problems.random_conic_batch(batch=592, N=64, seed=99)   # 2 CTAs per SM resident: one full wave
sol = P.LQRCudaSolver.from_problem(hp, num_segments=1)
nct = hp.nc_total
rho = np.full((hp.batch, nct), 0.1); inv = 1.0 / rho
z = np.zeros((hp.batch, nct)); y = np.zeros((hp.batch, nct))
for _ in range(2):
    sol.update_problem_data(hp.zeros_ws(), y, z, inv, sigma=1e-6)
    sol.backward(rho)
    sol.forward(hp.x0, hp.zeros_ws())
sol.synchronize()
