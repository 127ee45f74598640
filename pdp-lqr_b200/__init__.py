"""pdplqr_b200 -- B200-native (sm_100a, FP64) parallel dynamic-programming LQ solve.

Contents: csrc/ (hand-written CUDA kernels + the C ABI of include/pdplqr.h), capi.py (ctypes binding),
solver.py (host-side mirror of the reference's solver protocol), problems.py (synthetic problem generators),
mpc.py (receding-horizon driver over the conic ADMM solve).
The directory is named `pdp-lqr_b200`; import it as `pdplqr_b200` (shim module at the repo root)."""
from . import capi, mpc, problems  # noqa: F401
from ._build import build  # noqa: F401
from .solver import CHOLESKY, LU, LQRCudaSolver, PdplqrError, wave_size  # noqa: F401

__all__ = ["capi", "mpc", "problems", "build", "LQRCudaSolver", "PdplqrError", "LU", "CHOLESKY", "wave_size"]
