"""In-tree build of libpdplqr.so (sm_100a only).  nvcc cross-compiles without a GPU.

The kernels are heavy templates, so every instantiated (nx, nu) pair of csrc/inst_list.h is its own translation unit
(csrc/inst.cu compiled with -DINST_NX/-DINST_NU/-DINST_T); the units are compiled in parallel and linked with
csrc/pdplqr.cu (C ABI + orchestration).  Objects are cached under pdp-lqr_b200/build/ (git-ignored)."""
from __future__ import annotations

import os
import re
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# PDPLQR_VARIANT=name + PDPLQR_CFLAGS="-D..." build an instrumented copy (libpdplqr_<name>.so) next to the product library
VARIANT = os.environ.get("PDPLQR_VARIANT", "")
OBJ = os.path.join(HERE, "build", "variant_" + VARIANT) if VARIANT else os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpdplqr" + ("_" + VARIANT if VARIANT else "") + ".so")
NVCC = os.environ.get("PDPLQR_NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
         "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC", "-diag-suppress", "177"] + os.environ.get("PDPLQR_CFLAGS", "").split()


def sources():
    out = [os.path.join(ROOT, "include", "pdplqr.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def instances():
    """(nx, nu, T) triples of csrc/inst_list.h."""
    txt = open(os.path.join(CSRC, "inst_list.h")).read()
    txt = re.sub(r"//.*", "", txt)
    return [tuple(int(v) for v in m) for m in re.findall(r"X\(\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*\)", txt)]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def _units():
    units = [("pdplqr", os.path.join(CSRC, "pdplqr.cu"), []), ("sharded", os.path.join(CSRC, "sharded.cu"), [])]
    for nx, nu, t in instances():
        units.append((f"inst_{nx}_{nu}", os.path.join(CSRC, "inst.cu"), [f"-DINST_NX={nx}", f"-DINST_NU={nu}", f"-DINST_T={t}"]))
    return units


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> str:
    if not (force or stale()):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    newest = max(os.path.getmtime(s) for s in sources())
    todo, objs = [], []
    for name, src, defs in _units():
        obj = os.path.join(OBJ, name + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest:
            todo.append([NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + defs + ["-c", "-o", obj, src])
    jobs = jobs or min(len(todo) or 1, os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        for rc in ex.map(lambda cmd: subprocess.run(cmd).returncode, todo):
            if rc != 0:
                raise subprocess.CalledProcessError(rc, "nvcc")
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-ldl"])
    return LIB
