"""Applies the integration change of INTEGRATION.md section 1 to the reference's examples/lqr_example.cpp on the fly
(test infrastructure; the patched source is a temporary build input under oracle/_ref/ and is never committed):
  1. drop `#include "clqr/lqr/qdldl_solver.hpp"` and the QDLDLSolver block (QDLDL / Eigen::Sparse are absent here);
  2. add `#include "pdplqr/lqr_cuda_solver.hpp"`;
  3. `LQRParallelSolver lqr_parallel_solver(` -> `LQRCudaSolver lqr_parallel_solver(`   (the reference's line 213);
  4. print at full precision so that the GPU test can compare to 1e-9.
Everything else -- model construction with the reference's Node / LQRModel, initialize_vectors, the LQRSolver block,
the three protocol calls on the replaced solver -- is the reference's own text."""
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
s = open(src).read()
s = s.replace('#include "clqr/lqr/qdldl_solver.hpp"', '#include "pdplqr/lqr_cuda_solver.hpp"\n#include <iomanip>')
# cut the QDLDL block: from its construction up to (not including) the re-initialisation before LQRSolver
a = s.index("QDLDLSolver qdldl_solver(lqr_model);")
b = s.index("initialize_vectors(lqr_model, rho, ws, ys, zs, rho_vecs, inv_rho_vecs);", a)
s = s[:a] + "auto tic = std::chrono::high_resolution_clock::now();\n    auto toc = tic;\n    std::chrono::duration<double> elapsed = toc - tic;\n    std::cout << std::setprecision(15);\n    " + s[b:]
n = s.count("LQRParallelSolver lqr_parallel_solver(")
assert n >= 1
s = re.sub(r"(?m)^(\s*)LQRParallelSolver lqr_parallel_solver\(", r"\1LQRCudaSolver lqr_parallel_solver(", s)
s = s.replace("(LQRParallelSolver)", "(LQRCudaSolver)")
open(dst, "w").write(s)
