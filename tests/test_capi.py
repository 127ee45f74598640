"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports exactly the
symbols include/pdplqr.h declares.  No compute calls (no GPU here)."""
import os
import re

import pytest

import pdplqr_b200 as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "pdplqr.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pdplqr_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound():
    lib = P.capi.load()
    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pdplqr.h but not exported"
        assert name in P.capi.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(P.capi.SIGNATURES) == declared
    assert lib.pdplqr_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(P.PdplqrError) as e:
        P.LQRCudaSolver(12, 4, 100)
    assert e.value.code == P.capi.ERR_CUDA


def test_product_does_not_import_oracle():
    """The product path must not route through oracle/ (only tests, smoke() and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, "pdp-lqr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle-free", ""), f"{f} mentions the oracle"
    for f in ("include/pdplqr.h",):
        assert "liboracle" not in open(os.path.join(ROOT, f)).read()


def test_sass_has_tma_bulk_copies():
    """The kernels stage stage-records with TMA 1-D bulk copies: SASS must contain UBLKCP (B200_PROFILING.md)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-sass", P.capi.lib_path()], capture_output=True, text=True).stdout
    assert "UBLKCP" in out and "DFMA" in out and "sm_100a" in out
