// Host-side solver state + kernel launchers shared by pdplqr.cu (C ABI, orchestration) and inst.cu (one translation
// unit per instantiated (nx, nu) pair, compiled in parallel by _build.py -- the kernels are heavy templates).
#pragma once
#include "../../include/pdplqr.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "batch_kernels.cuh"
#include "seg_kernels.cuh"
#include "seg_warp_kernel.cuh"
#include "tree_kernels.cuh"
#include "tree_lat_kernels.cuh"
#include "costate_kernels.cuh"

using namespace pdplqr;   // kernel namespace (common.cuh)

namespace pdplqr_host {

struct TreeLevel {
    int count, R, groups;
    double *sum, *dd, *x, *lam;  // sum/x/lam of level 0 alias the segment arrays
    bool in_top;                 // handled inside the single-launch top kernels (binary)
};

struct Ops;

}  // namespace pdplqr_host
using pdplqr_host::Ops;
using pdplqr_host::TreeLevel;

struct pdplqr_solver {
    int nx = 0, nu = 0, N = 0, batch = 0, S = 1, s = 0, device = 0;   // nx, nu, s: KERNEL dimensions (an instantiated pair)
    // Caller's dimensions.  When (nxu, nuu) is not an instantiated pair the problem is embedded in the cheapest pair that
    // contains it (padded states: A = 0, B = 0, c = 0, Q = I, x0 = 0; padded inputs: R = I, zero columns of B and D), so
    // the padded components stay exactly zero and the caller's components are unchanged; every array that crosses the
    // C ABI is converted between the two layouts by the repack kernels (pdplqr.cu).
    int nxu = 0, nuu = 0, su = 0;
    bool padded = false;
    double *d_wsp_in = nullptr, *d_wsp_out = nullptr, *d_x0p = nullptr;   // kernel-layout staging (padded handles only)
    double *d_cost_ws = nullptr, *d_cost_lam = nullptr;                   // scratch of the host costate accessor (lazy)
    bool root_fresh = false;       // a root boundary was set since the last forward (consumed by forward)
    bool load_balancing = true;
    int condensed_type = 1;
    std::vector<int> ncs, seg_start, seg_len;
    long long nc_total = 0;
    const Ops* ops = nullptr;
    bool thread_path = false;  // thread-per-problem kernels (tiny nx+nu, S == 1)
    int frec = 0;              // doubles per stage in d_fac for the active path
    int mrec = 0;              // doubles per stage in d_model for the active path
    int bwd_variant = 0, fwd_variant = 0;
    int seg_mode = 0, seg_len0 = 0;   // closed-form partition handed to the kernels
    int lat_threads = 128;     // 0 disables the 128-thread latency mode of the segment backward kernel
    int seg_t = 0;             // PDPLQR_SEG_T: threads per (problem, segment) in throughput mode (0 = default 32)
    int warp_kernel = 1;       // PDPLQR_WARP_KERNEL: register-resident warp kernel (0 off, 1 throughput mode, 2 always)
    bool chain_tail = false;   // the last operation this library enqueued on the stream is a launch_chain kernel of the current call
    int fused = 0;             // inside pdplqr_solve_device: backward and forward are one chain (no API boundary in between)
    int pdl_mask = 0x1f;       // PDPLQR_PDL_MASK: launch sites that may carry the attribute (bit 0 stage sweep, 1 warp stage sweep,
                               // 2 rollout, 3 tree up, 4 tree down)
    int use_pdl = 0;           // programmatic dependent launch of the solve chain (launch_chain); set at create: on in the
                               // latency regime (batch * segments <= 2 x 148), PDPLQR_PDL = 0 / 1 overrides
    int tree_lat = 1;          // latency-mode tree kernels when a level has few groups (PDPLQR_TREE_LAT=0 disables)
    int tree_lat_max = 296;    // ... "few" = at most this many CTAs (PDPLQR_TREE_LAT_MAX)
    int lat_width = 0, lat_tt_cap = 0;   // PDPLQR_TREE_LAT_WIDTH / PDPLQR_TREE_LAT_TT: tuning overrides (0 = default)
    bool top_lat = false;      // the upper levels are binary and run on the latency-mode sub-tree kernels
    struct LatGroup { int l0, l1, width; };   // one launch: levels[l0] .. levels[l1], `width` nodes of l0 per CTA
    std::vector<LatGroup> lat_groups;         // bottom -> top
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // device memory
    double *d_model = nullptr, *d_HN = nullptr, *d_hN = nullptr;
    double *d_fac = nullptr, *d_sum = nullptr, *d_xhat = nullptr, *d_uhat = nullptr;
    double *d_ws_in = nullptr, *d_x0 = nullptr, *d_ws_out = nullptr;
    int *d_seg_start = nullptr, *d_seg_len = nullptr, *d_status = nullptr;
    // constraints
    int ncmax = 0;
    long long d_total_host = 0, d_total_dev = 0;
    std::vector<long long> coff, doff_host, doff_dev;
    int* d_ncs = nullptr;
    long long *d_coff = nullptr, *d_doff = nullptr, *d_doff_host = nullptr;
    // selection-matrix constraints (every row of D has at most one non-zero): compact (column, value) form
    int* d_sel_col = nullptr;
    double* d_sel_val = nullptr;
    int* d_sel_flag = nullptr;
    bool sel_mode = false;
    int allow_sel = 1;   // PDPLQR_SPARSE_D=0 keeps the dense path (A/B tests)
    double *d_D = nullptr, *d_ys = nullptr, *d_zs = nullptr, *d_rho = nullptr, *d_inv_rho = nullptr;
    const double *cur_ys = nullptr, *cur_zs = nullptr, *cur_inv_rho = nullptr, *cur_rho = nullptr;
    // affine cache for backward_without_factorization
    bool keep_affine = false;
    double* d_aff = nullptr;
    // conic ADMM outer loop (a11)
    int ncones = 0;
    int *d_cone_first = nullptr, *d_cone_type = nullptr, *d_cone_row = nullptr, *d_cone_dim = nullptr, *d_row_box = nullptr;
    double *d_elb = nullptr, *d_eub = nullptr, *d_wtilde = nullptr, *d_w = nullptr, *d_z = nullptr, *d_y = nullptr, *d_rho_admm = nullptr,
           *d_invrho_admm = nullptr;
    bool cones_set = false;
    // device-resident outer loop (pdplqr.cu, admm_kernels.cuh): control block, library-owned rho / 1/rho (rescaled by the
    // adaptation), kernel-layout iterate of padded handles, and the CUDA graph of one conic solve with the key it was
    // captured for
    void* d_ctl = nullptr;
    double *d_rho_work = nullptr, *d_invrho_work = nullptr, *d_wk = nullptr;
    int admm_use_graph = 1, admm_adaptive = 0, admm_max_rho_updates = 10;
    double admm_rho_tau = 5.0;
    cudaGraph_t admm_graph = nullptr;
    cudaGraphExec_t admm_exec = nullptr;
    struct AdmmKey { const void *x0, *w, *z, *y; double sigma, alpha; cudaStream_t stream; } admm_key{};
    int admm_graph_launches = 0, admm_rho_updates_last = 0;
    // CUDA graph of one update_problem_data + backward + forward (pdplqr_solve_device), and the pointers it was captured for
    int solve_use_graph = 1;
    cudaGraphExec_t solve_exec = nullptr;
    struct SolveKey { const void *ws, *ys, *zs, *rho, *inv_rho, *x0, *out; double sigma; cudaStream_t stream; } solve_key{};
    int solve_kernels = 0;
    int admm_fact_kernels = 0, admm_aff_kernels = 0;   // kernels per factorising / affine-only iteration inside the graph
    // pipelined host solve (H2D / compute / D2H overlapped over batch chunks)
    int chunk_b0 = 0, chunk_nb = 0;   // when chunk_nb > 0 the launchers work on problems [b0, b0 + nb)
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_cmp;
    cudaEvent_t ev_ready = nullptr;
    int pipeline_chunks = 8;
    // horizon sharding
    bool interior = false;         // slice ends at an interface (not the true terminal)
    bool is_coupler = false;       // handle created by pdplqr_coupler_create (interface tree only)
    double *d_root_x = nullptr, *d_root_lam = nullptr;
    bool have_root = false;
    std::vector<TreeLevel> levels;
    std::vector<void*> owned;  // everything to cudaFree
    // PDPLQR_DEBUG_GUARDS=1 (read at create): every device allocation of the handle is filled with 0xFF bytes (NaN as a
    // double, -1 as an int: a read of memory the library never wrote shows up in the results) and sits between two
    // GUARD_BYTES bands of 0xA5 that pdplqr_debug_check_guards inspects (out-of-bounds writes).  compute-sanitizer is closed
    // on the pool this was developed on; the parity suite run in this mode is the memcheck substitute.
    bool guards = false;
    static constexpr size_t GUARD_BYTES = 4096;
    struct GuardedAlloc { char* base; size_t bytes; };
    std::vector<GuardedAlloc> guarded;
    // per-iteration state
    const double* cur_ws = nullptr;  // device pointer used by the next backward (nullptr == zeros)
    double sigma = 0.0;
    bool model_set = false, updated = false, factorized = false, backward_done = false;
    long long launches = 0;
    std::string err;
};

namespace pdplqr_host {

using Solver = pdplqr_solver;

struct Ops {
    int nx, nu, T;
    int REC, FREC, SREC, DREC, FRECT, TREC, AREC;
    int top_lat_nodes;   // nodes one CTA of the latency-mode sub-tree kernels reduces (0: not built for this nx)
    int lat_tt_cap;      // threads per combine at most
    bool has_thread_path;
    int (*backward)(Solver&);
    int (*forward)(Solver&, const double* d_x0, double* d_ws_out);
    int (*tree_up)(Solver&, const TreeParams&);
    int (*tree_down)(Solver&, const TreeParams&);
    int (*affine)(Solver&);
    int (*tree_up_affine)(Solver&, const TreeParams&);
    int (*tree_top_up)(Solver&, const TreeTopParams&);
    int (*tree_top_down)(Solver&, const TreeTopParams&);
    int (*tree_sub_up)(Solver&, const TreeTopParams&);
    int (*tree_sub_down)(Solver&, const TreeTopParams&);
    int (*costates)(Solver&, const double* traj, double* lam);
    int (*wave)(int ncmax, bool sel);   // resident (problem, segment) CTAs per SM of the throughput-mode stage kernel
};

inline int fail(Solver* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}
#define CU_TRY(h, expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return fail(h, PDPLQR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    } while (0)

template <class K>
int set_smem(Solver& h, K kernel, size_t bytes) {
    if (bytes > 48 * 1024) CU_TRY(&h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return PDPLQR_OK;
}

// Launch of a solve-chain kernel whose body starts with pdl_wait() (common.cuh).  EXPERIMENTAL, off by default (PDPLQR_PDL=1):
// with h.use_pdl the launch carries the programmatic stream serialization attribute, so that the kernel is scheduled while
// its predecessor drains; stream capture turns the attribute into a programmatic edge of the solve graph.
// Measured on B200 (C2, N = 1024, S = 128; gpurun_out r14 / r15 / r16 of round 2, summarised in DESIGN.md section 9):
//   graph-launched solve 93.5 -> 92.3 us (graph edges are already tight), protocol calls 104.7 -> 92.6 us.
//   With the attribute on every chain launch, constrained handles with host buffers failed parity (rel. error 0.07 - 0.12, timing
//   dependent: nx12/nu4/nc8/S3 in test_backward_without_factorization and test_horizon_shards_with_constraints).  A per-site
//   mask (PDPLQR_PDL_MASK) pins it on ONE edge: stage sweep -> first tree launch (bit 3); rollout (bit 2) and tree-down
//   (bit 4) edges pass.  The stage sweep there is queued behind the asynchronous H2D copies of ys / zs / rho / 1/rho, and the
//   tree kernel, whose only dependency is the programmatic one, read summaries the sweep had not finished writing although it
//   starts with griddepcontrol.wait.  Not root-caused within the GPU budget of the round, so the default stays the ordinary
//   launch; the wait / trigger instructions in the kernels are no-ops then.
template <class K, class Prm>
inline void launch_chain(Solver& h, int site, K kern, int grid, int block, size_t smem, const Prm& prm) {
    // Only behind another launch_chain kernel of the same API call: griddepcontrol.wait waits for prerequisite GRIDS only, and
    // work the caller enqueued between two API calls is unknown -- every call starts its chain with an ordinary launch.
    const bool pdl = h.use_pdl && h.chain_tail && ((h.pdl_mask >> site) & 1);
    h.chain_tail = true;
    if (!pdl) {
        kern<<<grid, block, smem, h.stream>>>(prm);
        return;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h.stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, prm);   // the caller checks cudaGetLastError() as after a <<< >>> launch
}

inline SegParams seg_params(Solver& h) {
    SegParams p{};
    p.N = h.N; p.S = h.S; p.batch = h.batch; p.interior = h.interior ? 1 : 0;
    p.seg_start = h.d_seg_start; p.seg_len = h.d_seg_len;
    p.seg_mode = h.seg_mode; p.seg_len0 = h.seg_len0;
    p.model = h.d_model; p.HN = h.d_HN; p.hN = h.d_hN;
    p.ws_prev = h.cur_ws; p.sigma = h.sigma;
    p.fac = h.d_fac; p.sum = h.d_sum; p.status = h.d_status;
    p.xhat = h.d_xhat; p.uhat = h.d_uhat;
    p.aff = h.keep_affine ? h.d_aff : nullptr;
    p.ncmax = h.ncmax; p.nc_total = h.nc_total; p.d_total = h.d_total_dev;
    p.ncs = h.d_ncs; p.coff = h.d_coff; p.doff = h.d_doff; p.Dm = h.d_D;
    p.ys = h.cur_ys; p.zs = h.cur_zs; p.rho = h.cur_rho; p.inv_rho = h.cur_inv_rho;
    p.sel_col = h.sel_mode ? h.d_sel_col : nullptr;
    p.sel_val = h.sel_mode ? h.d_sel_val : nullptr;
    if (h.chunk_nb > 0) {   // batch chunk [b0, b0 + nb): offset every per-problem array (nc = 0 handles only)
        const size_t b0 = h.chunk_b0, wsl = (size_t)h.N * h.s + h.nx;
        p.batch = h.chunk_nb;
        p.model += b0 * h.N * h.mrec; p.HN += b0 * h.nx * h.nx; p.hN += b0 * h.nx;
        if (p.ws_prev) p.ws_prev += b0 * wsl;
        p.fac += b0 * h.N * h.frec; p.sum += b0 * h.S * h.ops->SREC; p.status += b0 * h.S;
        p.xhat += b0 * h.S * h.nx; p.uhat += b0 * h.S * h.nx;
    }
    return p;
}

// Thread-path launch variants (warps per CTA, TMA ring depth, min CTAs per SM).  The default is chosen so that a
// 65,536-problem batch runs in whole waves on 148 SMs (DESIGN.md "grid sizing"); PDPLQR_BWD_VARIANT /
// PDPLQR_FWD_VARIANT select another one for tuning sweeps.
template <int NX, int NU, int WARPS, int DEPTH, int MINB>
int launch_batch_bwd(Solver& h, const SegParams& p) {
    auto kern = batch_backward_kernel<NX, NU, WARPS, DEPTH, MINB>;
    constexpr size_t bytes = BatchBwdSmem<NX, NU, WARPS, DEPTH>::BYTES;
    int rc = set_smem(h, kern, bytes);
    if (rc) return rc;
    const int blocks = (p.batch + WARPS * 32 - 1) / (WARPS * 32);
    kern<<<blocks, WARPS * 32, bytes, h.stream>>>(p);
    h.chain_tail = false;
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}
template <int NX, int NU, int WARPS, int DEPTH, int MINB>
int launch_batch_fwd(Solver& h, const SegParams& p) {
    auto kern = batch_forward_kernel<NX, NU, WARPS, DEPTH, MINB>;
    constexpr size_t bytes = BatchFwdSmem<NX, NU, WARPS, DEPTH>::BYTES;
    int rc = set_smem(h, kern, bytes);
    if (rc) return rc;
    const int blocks = (p.batch + WARPS * 32 - 1) / (WARPS * 32);
    kern<<<blocks, WARPS * 32, bytes, h.stream>>>(p);
    h.chain_tail = false;
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}

template <int NX, int NU, int TT, bool CON>
int launch_seg_bwd(Solver& h, const SegParams& p, size_t bytes) {
    auto kern = seg_backward_kernel<NX, NU, TT, CON>;
    int rc = set_smem(h, kern, bytes);
    if (rc) return rc;
    launch_chain(h, 0, kern, h.batch * h.S, TT, bytes, p);
    return PDPLQR_OK;
}

template <int NX, int NU, int T>
int backward_impl(Solver& h) {
    SegParams p = seg_params(h);
    if constexpr (BatchDims<NX, NU>::ENABLED) {
        if (h.thread_path) {
            switch (h.bwd_variant) {
                case 1: return launch_batch_bwd<NX, NU, 2, 1, 7>(h, p);
                case 2: return launch_batch_bwd<NX, NU, 7, 2, 1>(h, p);
                case 3: return launch_batch_bwd<NX, NU, 14, 1, 1>(h, p);
                case 4: return launch_batch_bwd<NX, NU, 5, 3, 1>(h, p);
                default: return launch_batch_bwd<NX, NU, 4, 2, 2>(h, p);
            }
        }
    }
    // latency mode: with fewer (problem, segment) groups than SMs a whole 128-thread CTA works on each group
    constexpr int TL = (T < 128) ? 128 : T;
    const bool latency_mode = (T < 128) && h.lat_threads > 0 && (long long)h.batch * h.S <= 2 * 148;
    if constexpr (SegDims<NX, NU>::WLAY) {
        // register-resident warp kernel (seg_warp_kernel.cuh): unconstrained stages, no affine cache.
        // PDPLQR_WARP_KERNEL = 0: off, 1 (default): throughput mode, 2: latency mode as well
        if (h.ncmax == 0 && !h.keep_affine && h.seg_t == 0 && (h.warp_kernel >= 2 || (h.warp_kernel == 1 && !latency_mode))) {
            auto kern = seg_backward_warp_kernel<NX, NU>;
            constexpr size_t wbytes = WarpSmem<NX, NU>::BYTES;
            int rc = set_smem(h, kern, wbytes);
            if (rc) return rc;
            launch_chain(h, 1, kern, h.batch * h.S, 32, wbytes, p);
            h.launches++;
            CU_TRY(&h, cudaGetLastError());
            return PDPLQR_OK;
        }
    }
    const size_t bytes = BwdSmem<NX, NU>::bytes(h.ncmax, h.sel_mode);
    if constexpr (T == 32) {
        // One warp per (problem, segment) by default: with the products on register-blocked DMMA a single warp owns every
        // output tile and reuses its operand fragments most (C5: 3.46 ms, against 3.58 ms with two warps sharing the
        // working set).  PDPLQR_SEG_T = 32 / 64 / 128 overrides.
        const int seg_t = h.seg_t ? h.seg_t : 32;
        if (!latency_mode && seg_t == 64) {
            int rc = h.ncmax > 0 ? launch_seg_bwd<NX, NU, 64, true>(h, p, bytes) : launch_seg_bwd<NX, NU, 64, false>(h, p, bytes);
            if (rc) return rc;
            h.launches++;
            CU_TRY(&h, cudaGetLastError());
            return PDPLQR_OK;
        }
    }
    int rc;
    if (latency_mode || (T == 32 && h.seg_t == 128))
        rc = h.ncmax > 0 ? launch_seg_bwd<NX, NU, TL, true>(h, p, bytes) : launch_seg_bwd<NX, NU, TL, false>(h, p, bytes);
    else
        rc = h.ncmax > 0 ? launch_seg_bwd<NX, NU, T, true>(h, p, bytes) : launch_seg_bwd<NX, NU, T, false>(h, p, bytes);
    if (rc) return rc;
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}

template <int NX, int NU, int T>
int forward_impl(Solver& h, const double* d_x0, double* d_ws_out) {
    SegParams p = seg_params(h);
    p.ws_out = d_ws_out;
    if (h.S == 1) {
        p.xhat = d_x0;
        if (h.have_root) p.uhat = h.d_root_lam;
    }
    if (h.chunk_nb > 0) {
        p.ws_out += (size_t)h.chunk_b0 * ((size_t)h.N * h.s + h.nx);
        if (h.S == 1) p.xhat += (size_t)h.chunk_b0 * h.nx;
    }
    if constexpr (BatchDims<NX, NU>::ENABLED) {
        if (h.thread_path) {
            switch (h.fwd_variant) {
                case 1: return launch_batch_fwd<NX, NU, 2, 2, 7>(h, p);
                case 2: return launch_batch_fwd<NX, NU, 7, 3, 1>(h, p);
                case 3: return launch_batch_fwd<NX, NU, 10, 2, 1>(h, p);
                case 4: return launch_batch_fwd<NX, NU, 7, 2, 1>(h, p);
                default: return launch_batch_fwd<NX, NU, 4, 2, 3>(h, p);
            }
        }
    }
    constexpr int TF = (NX + NU >= 32) ? 128 : 32;   // latency-bound rollout: more warps per SM for big stages
    auto kern = seg_forward_kernel<NX, NU, TF>;
    constexpr size_t bytes = FwdSmem<NX, NU>::BYTES;
    int rc = set_smem(h, kern, bytes);
    if (rc) return rc;
    launch_chain(h, 2, kern, h.batch * h.S, TF, bytes, p);
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}

template <int NX>
int tree_up_impl(Solver& h, const TreeParams& p) {
    constexpr size_t bytes = TreeSmem<NX>::BYTES;
    auto kern = tree_up_kernel<NX, 32>;
    int rc = set_smem(h, kern, bytes);
    if (rc) return rc;
    kern<<<p.batch * p.groups, 32, bytes, h.stream>>>(p);
    h.chain_tail = false;
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}
template <int NX>
int tree_down_impl(Solver& h, const TreeParams& p) {
    auto kern = tree_down_kernel<NX>;
    kern<<<p.batch * p.groups, 32, 4 * NX * sizeof(double), h.stream>>>(p);
    h.chain_tail = false;
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}

template <int NX, int NU>
int affine_impl(Solver& h) {
    SegParams p = seg_params(h);
    constexpr int TA = (NX + NU >= 32) ? 128 : 32;   // the sweep is latency-bound: more warps per SM for big stages
    auto kern = seg_affine_kernel<NX, NU, TA>;
    const size_t bytes = AffSmem<NX, NU>::bytes(h.ncmax, h.sel_mode, p.S == 1 && !p.interior);
    int rc = set_smem(h, kern, bytes);
    if (rc) return rc;
    kern<<<h.batch * h.S, TA, bytes, h.stream>>>(p);
    h.chain_tail = false;
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}
template <int NX>
int tree_up_affine_impl(Solver& h, const TreeParams& p) {
    static_assert(NX <= 32, "tree_up_affine_kernel keeps one row per lane");
    auto kern = tree_up_affine_kernel<NX>;
    kern<<<p.batch * p.groups, 32, 5 * NX * sizeof(double), h.stream>>>(p);
    h.chain_tail = false;
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}

template <int NX>
int tree_top_up_impl(Solver& h, const TreeTopParams& p) {
    constexpr size_t bytes = TreeTopSmem<NX>::BYTES;
    constexpr int WARPS = TreeTopSmem<NX>::WARPS;
    auto kern = tree_top_up_kernel<NX, 32>;
    int rc = set_smem(h, kern, bytes);
    if (rc) return rc;
    kern<<<p.batch, WARPS * 32, bytes, h.stream>>>(p);
    h.chain_tail = false;
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}
template <int NX>
int tree_top_down_impl(Solver& h, const TreeTopParams& p) {
    auto kern = tree_top_down_kernel<NX>;
    kern<<<p.batch, TreeTopSmem<NX>::WARPS * 32, TreeTopSmem<NX>::WARPS * 4 * NX * sizeof(double), h.stream>>>(p);
    h.chain_tail = false;
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}

// latency-mode binary sub-tree launches (tree_lat_kernels.cuh)
template <int NX>
int tree_sub_up_impl(Solver& h, const TreeTopParams& p) {
    if constexpr (LatSmem<NX>::DOWN_OK) {
        auto kern = tree_sub_up_lat_kernel<NX>;
        int rc = set_smem(h, kern, LatSmem<NX>::TOP_BYTES);
        if (rc) return rc;
        launch_chain(h, 3, kern, p.batch * p.ngroups, LatSmem<NX>::TOP_THREADS, LatSmem<NX>::TOP_BYTES, p);
        h.launches++;
        CU_TRY(&h, cudaGetLastError());
        return PDPLQR_OK;
    } else
        return fail(&h, PDPLQR_ERR_UNSUPPORTED, "latency-mode tree kernels are not built for this nx");
}
template <int NX>
int tree_sub_down_impl(Solver& h, const TreeTopParams& p) {
    if constexpr (LatSmem<NX>::DOWN_OK) {
        auto kern = tree_sub_down_lat_kernel<NX>;
        int rc = set_smem(h, kern, LatSmem<NX>::DOWN_BYTES);
        if (rc) return rc;
        launch_chain(h, 4, kern, p.batch * p.ngroups, LatSmem<NX>::TOP_THREADS, LatSmem<NX>::DOWN_BYTES, p);
        h.launches++;
        CU_TRY(&h, cudaGetLastError());
        return PDPLQR_OK;
    } else
        return fail(&h, PDPLQR_ERR_UNSUPPORTED, "latency-mode tree kernels are not built for this nx");
}

// costates of the last solve (costate_kernels.cuh); segment-path handles only
template <int NX, int NU>
int costates_impl(Solver& h, const double* traj, double* lam) {
    CostateParams q{};
    q.sp = seg_params(h);
    q.traj = traj; q.lam = lam;
    q.lam_root = (h.interior && h.have_root) ? h.d_root_lam : nullptr;
    if constexpr (BatchDims<NX, NU>::ENABLED) {
        if (h.thread_path) {
            batch_costate_kernel<NX, NU><<<(h.batch + 127) / 128, 128, 0, h.stream>>>(q);
            h.chain_tail = false;
            h.launches++;
            CU_TRY(&h, cudaGetLastError());
            return PDPLQR_OK;
        }
    }
    const size_t bytes = (size_t)(NX + (NX + NU) + std::max(h.ncmax, 1)) * sizeof(double);
    auto kern = seg_costate_kernel<NX, NU>;
    int rc = set_smem(h, kern, bytes);
    if (rc) return rc;
    kern<<<h.batch * h.S, 32, bytes, h.stream>>>(q);
    h.chain_tail = false;
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}

// resident CTAs per SM of the stage kernel a throughput-mode backward would launch (unconstrained: the warp kernel where
// its layout is on, else seg_backward_kernel with T threads)
template <int NX, int NU, int T>
int wave_impl(int ncmax, bool sel) {
    int n = 0;
    if constexpr (SegDims<NX, NU>::WLAY) {
        if (ncmax == 0) {
            auto kern = seg_backward_warp_kernel<NX, NU>;
            constexpr size_t wbytes = WarpSmem<NX, NU>::BYTES;
            if (wbytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wbytes);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 32, wbytes) != cudaSuccess) n = 0;
            return n;
        }
    }
    const size_t bytes = BwdSmem<NX, NU>::bytes(ncmax, sel);
    if (ncmax > 0) {
        auto kern = seg_backward_kernel<NX, NU, T, true>;
        if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, T, bytes) != cudaSuccess) n = 0;
    } else {
        auto kern = seg_backward_kernel<NX, NU, T, false>;
        if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, T, bytes) != cudaSuccess) n = 0;
    }
    return n;
}

template <int NX, int NU, int T>
constexpr Ops make_ops() {
    return Ops{NX, NU, T, SegDims<NX, NU>::REC, SegDims<NX, NU>::FREC, SegDims<NX, NU>::SREC, TreeDims<NX>::DREC,
               BatchDims<NX, NU>::FRECT, BatchDims<NX, NU>::TREC, SegDims<NX, NU>::AREC,
               (LatSmem<NX>::DOWN_OK ? LatSmem<NX>::TOP_NODES : 0), LatSmem<NX>::TT_CAP,
               BatchDims<NX, NU>::ENABLED,
               &backward_impl<NX, NU, T>, &forward_impl<NX, NU, T>, &tree_up_impl<NX>, &tree_down_impl<NX>,
               &affine_impl<NX, NU>, &tree_up_affine_impl<NX>, &tree_top_up_impl<NX>, &tree_top_down_impl<NX>,
               &tree_sub_up_impl<NX>, &tree_sub_down_impl<NX>, &costates_impl<NX, NU>, &wave_impl<NX, NU, T>};
}

}  // namespace pdplqr_host
