import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_err(a, b):
    import numpy as np
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(autouse=True)
def _guard_bands_clean():
    """PDPLQR_DEBUG_GUARDS=1 (the memcheck substitute, include/pdplqr.h: pdplqr_debug_check_guards): every handle a test
    created is checked when it is closed; a test that overwrote a guard band fails here."""
    yield
    import pdplqr_b200 as P
    st = P.solver.GUARD_STATS
    if st["enabled"]:
        import gc
        gc.collect()
        bad, ung = st["corrupted_bytes"], st["unguarded"]
        st["corrupted_bytes"] = st["unguarded"] = 0
        assert bad == 0, f"{bad} guard bytes around device allocations were overwritten (out-of-bounds write)"
        assert ung == 0, "a handle was created without guard bands although PDPLQR_DEBUG_GUARDS=1"
