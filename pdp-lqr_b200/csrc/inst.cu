// One translation unit per instantiated (nx, nu) pair: compiled as  nvcc -DINST_NX=.. -DINST_NU=.. -DINST_T=..  (the list
// lives in inst_list.h; _build.py compiles the units in parallel and links them with pdplqr.cu).
#include "solver_impl.cuh"

#define PDPLQR_CAT3(a, b, c) a##b##_##c
#define PDPLQR_OPS_NAME(nx, nu) PDPLQR_CAT3(pdplqr_ops_, nx, nu)

namespace pdplqr_host {
const Ops* PDPLQR_OPS_NAME(INST_NX, INST_NU)() {
    static const Ops ops = make_ops<INST_NX, INST_NU, INST_T>();
    return &ops;
}
}  // namespace pdplqr_host
