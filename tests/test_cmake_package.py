"""The CMake package (reference: CMakeLists.txt:52-77, cmake/pdpLQRConfig.cmake.in): configure, build
and install libpdplqr + headers, then build a consumer with find_package(pdplqr) / pdplqr::pdplqr.  CPU only (nvcc
cross-compiles sm_100a); the consumer must fail loudly without a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cmake_package_builds_installs_and_is_consumable(tmp_path):
    cmake = shutil.which("cmake")
    if not cmake or not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("cmake / nvcc not available")
    build, prefix = tmp_path / "build", tmp_path / "prefix"
    gen = ["-G", "Ninja"] if shutil.which("ninja") else []
    subprocess.check_call([cmake, "-S", ROOT, "-B", str(build), "-DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc",
                           f"-DCMAKE_INSTALL_PREFIX={prefix}", "-DPDPLQR_BUILD_EXAMPLES=OFF"] + gen,
                          stdout=subprocess.DEVNULL)
    subprocess.check_call([cmake, "--build", str(build), "-j", "8"], stdout=subprocess.DEVNULL)
    subprocess.check_call([cmake, "--install", str(build)], stdout=subprocess.DEVNULL)
    assert (prefix / "include" / "pdplqr.h").exists() and (prefix / "include" / "pdplqr" / "lqr_cuda_solver.hpp").exists()
    assert (prefix / "lib" / "cmake" / "pdplqr" / "pdplqrConfig.cmake").exists()
    # consumer project, the way the reference's examples/CMakeLists.txt links pdpLQR::pdpLQR
    cons = tmp_path / "consumer"
    cons.mkdir()
    (cons / "CMakeLists.txt").write_text(
        "cmake_minimum_required(VERSION 3.24)\nproject(consumer LANGUAGES CXX)\nset(CMAKE_CXX_STANDARD 17)\n"
        "find_package(pdplqr REQUIRED)\n"
        f"add_executable(lqr_example {ROOT}/examples/lqr_example.cpp)\n"
        "target_link_libraries(lqr_example PRIVATE pdplqr::pdplqr)\n")
    subprocess.check_call([cmake, "-S", str(cons), "-B", str(cons / "b"), f"-DCMAKE_PREFIX_PATH={prefix}"] + gen,
                          stdout=subprocess.DEVNULL)
    subprocess.check_call([cmake, "--build", str(cons / "b")], stdout=subprocess.DEVNULL)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([str(cons / "b" / "lqr_example")], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr
