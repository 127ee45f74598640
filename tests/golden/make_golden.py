"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/pdp_oracle.cpp) -- run here, committed with the
fixtures.  The reference ships no golden vectors, so these pin OUR oracle's answers; the
`survey_probe` values inside c1_quadrotor.npz come from an independent numpy restatement made during the survey
 and from the independent KKT solve in tests/kkt_ref.py.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import pdplqr_b200 as P  # noqa: E402
from kkt_ref import kkt_solve  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    # C1: the shipped example (examples/lqr_example.cpp), sigma = 1e-6, ws = 0, x0 = 0
    p = P.problems.quadrotor_example()
    o = O.OracleSolver(p)
    ws = o.solve()
    K, d, _ = o.gains()
    Pv, pv = o.value()
    op = O.OracleSolver(p, parallel=True, num_segments=4, load_balancing=True, condensed=O.CHOLESKY)
    wsp = op.solve()
    xh, uh = op.interface()
    np.savez_compressed(os.path.join(HERE, "c1_quadrotor.npz"), ws_seq=ws, K=K, d=d, P0=Pv[0], p0=pv[0],
                        ws_par4=wsp, xhat4=xh, uhat4=uh, ws_kkt=kkt_solve(p),
                        survey_probe=np.array([-2.898056669662, 0.050699574998, 1.023422557959, 0.988485801933,
                                               0.627371337009]))
    # random LTV problem with dense cost, cross terms, nonzero affine terms, sigma-term active
    q = P.problems.random_lq(6, 3, 40, batch=1, seed=11)
    rng = np.random.default_rng(5)
    wprev = rng.standard_normal(q.ws_len)
    oq = O.OracleSolver(q)
    wsq = oq.solve(ws_in=wprev, sigma=0.05)
    np.savez_compressed(os.path.join(HERE, "random_6_3_40.npz"), wprev=wprev, ws_seq=wsq,
                        ws_kkt=kkt_solve(q, ws_prev=wprev, sigma=0.05))
    print("golden fixtures written")


if __name__ == "__main__":
    main()
