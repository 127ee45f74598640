"""The C++17 host class lqr::LQRCudaSolver (include/pdplqr/lqr_cuda_solver.hpp) mirrors the reference's solver
interface on top of the C ABI.  CPU: the example driver compiles and links against libpdplqr.so and fails loudly
without a GPU.  GPU: it reproduces config 1 (examples/lqr_example.cpp as shipped) to 1e-9."""
import os
import re
import subprocess

import numpy as np
import pytest

import pdplqr_b200 as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "lqr_example_bin")


def _compile():
    libdir = os.path.dirname(P.capi.lib_path())
    P.capi.load()
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "lqr_example.cpp"), "-L" + libdir, "-lpdplqr",
                           "-Wl,-rpath," + libdir, "-o", EXE])


def test_cpp_example_compiles_and_has_no_cpu_fallback():
    import torch
    _compile()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_example_reproduces_config1():
    _compile()
    r = subprocess.run([EXE], capture_output=True, text=True, check=True)
    g = np.load(os.path.join(ROOT, "tests", "golden", "c1_quadrotor.npz"))
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 6
    for k in range(5):
        u = np.array([float(v) for v in re.findall(r"-?\d+\.\d+", lines[k].split(":")[1])])
        assert np.max(np.abs(u - g["ws_seq"][k * 16:k * 16 + 4])) < 1e-9
    xN = np.array([float(v) for v in lines[5].split(":")[1].split()])
    assert np.max(np.abs(xN - g["ws_seq"][-12:])) < 1e-9


# ---------------------------------------------------------------- receding-horizon driver
MPC_EXE = os.path.join(ROOT, "examples", "mpc_example_bin")


def _compile_mpc():
    libdir = os.path.dirname(P.capi.lib_path())
    P.capi.load()
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "mpc_example.cpp"), "-L" + libdir, "-lpdplqr",
                           "-Wl,-rpath," + libdir, "-o", MPC_EXE])


def test_cpp_mpc_example_compiles_and_has_no_cpu_fallback():
    import torch
    _compile_mpc()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([MPC_EXE], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_mpc_example_closed_loop():
    """lqr::RecedingHorizon over the C ABI: 25 periods of the box-constrained quadrotor; the applied inputs respect
    their box up to the ADMM tolerance and the height moves towards the reference.  Cross-checked against the Python
    driver (same plant, same warm-start rule) on the first applied input."""
    _compile_mpc()
    r = subprocess.run([MPC_EXE], capture_output=True, text=True, check=True)
    out = r.stdout
    height = float(re.search(r"height after 25 periods: (-?\d+\.\d+)", out).group(1))
    viol = float(re.search(r"largest input-bound violation: (\S+)", out).group(1))
    u0 = [float(v) for v in re.search(r"period  0: u0 = (.*?)  z =", out).group(1).split()]
    assert viol < 2e-2 and 0.5 < height < 1.5
    from pdplqr_b200.mpc import RecedingHorizon
    p = P.problems.quadrotor_example(N=20, constrained=True)
    p.x0[0, 2] = 0.0
    sol = P.LQRCudaSolver.from_problem(p, num_segments=2)
    rh = RecedingHorizon(sol, p, rho=0.1, max_iter=400, eps_abs=1e-4, eps_rel=1e-4)
    u_py, _ = rh.step()
    assert np.max(np.abs(np.array(u0) - u_py[0])) < 5e-3


# ---------------------------------------------------------------- single-process multi-GPU solver (C ABI pdplqr_sharded_*)
SHARDED_EXE = os.path.join(ROOT, "examples", "sharded_example_bin")


def _compile_sharded():
    libdir = os.path.dirname(P.capi.lib_path())
    P.capi.load()
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "sharded_example.cpp"), "-L" + libdir, "-lpdplqr",
                           "-Wl,-rpath," + libdir, "-o", SHARDED_EXE])


def test_cpp_sharded_example_compiles_and_has_no_cpu_fallback():
    import torch
    _compile_sharded()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([SHARDED_EXE, "2", "256"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("G", [1, 2, 4])
def test_cpp_sharded_solver_matches_single_gpu(G):
    """lqr::LQRCudaShardedSolver (one process, G devices, NCCL all-gather of the slice summaries) against
    lqr::LQRCudaSolver on one device, from C++ (examples/sharded_example.cpp)."""
    import torch
    if torch.cuda.device_count() < G:
        pytest.skip(f"needs {G} GPUs")
    _compile_sharded()
    r = subprocess.run([SHARDED_EXE, str(G), "4096"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"devices {G}" in r.stdout
