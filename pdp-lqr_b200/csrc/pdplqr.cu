// libpdplqr.so -- C ABI (include/pdplqr.h) + host-side orchestration of the sm_100a kernels.
// One handle = one batch of identically-sized LQ problems resident on one GPU, one CUDA stream.
#include "solver_impl.cuh"
#include "inst_list.h"
#include "admm_kernels.cuh"

using namespace pdplqr_host;

namespace pdplqr_host {
#define PDPLQR_DECL_OPS(nx, nu, t) const Ops* pdplqr_ops_##nx##_##nu();
PDPLQR_INST_LIST(PDPLQR_DECL_OPS)
#undef PDPLQR_DECL_OPS
}  // namespace pdplqr_host

namespace {

thread_local std::string g_create_error;   // why the last pdplqr_create on this thread failed (pdplqr_last_error(NULL))

// registry of the instantiated (nx, nu) pairs (inst_list.h; one translation unit each)
const Ops* find_ops(int nx, int nu) {
#define PDPLQR_FIND_OPS(a, b, t) \
    if (nx == a && nu == b) return pdplqr_ops_##a##_##b();
    PDPLQR_INST_LIST(PDPLQR_FIND_OPS)
#undef PDPLQR_FIND_OPS
    return nullptr;
}
// cheapest instantiated pair with NX >= nx and NU >= nu (cost ~ the (nx + nu)^3 stage flops)
const Ops* find_padded_ops(int nx, int nu) {
    const Ops* best = nullptr;
    long best_cost = 0;
#define PDPLQR_PAD_OPS(a, b, t)                                                    \
    if (a >= nx && b >= nu) {                                                      \
        const long cost = (long)(a + b) * (a + b) * (a + b);                       \
        if (!best || cost < best_cost) { best = pdplqr_ops_##a##_##b(); best_cost = cost; } \
    }
    PDPLQR_INST_LIST(PDPLQR_PAD_OPS)
#undef PDPLQR_PAD_OPS
    return best;
}

// flat (E, c, H, h) -> device stage records.  sym = 0: [E | c | H | h] per (problem, stage) (segment kernels);
// sym = 1: thread-per-problem path: [E | c | lower(H) packed by columns | h] with the records of the 32 problems of a
// tile interleaved pair-wise (BatchDims::TR_*, tile_pos): one contiguous block per (tile, stage).  `nprob` is the
// batch size, the tile count is ceil(nprob / 32) and lanes past the batch replicate the last problem.
__global__ void pack_model_kernel(const double* __restrict__ E, const double* __restrict__ c,
                                  const double* __restrict__ H, const double* __restrict__ hv, double* __restrict__ rec,
                                  long long nprob, int N, int nx, int s, int REC, int sym, int nxu, int nuu) {
    // kernel dimensions (nx, s) vs the caller's (nxu, nuu): index i of w = [u; x] maps to the caller's index umap(i), or
    // -1 for a padded input / state (identity cost, zero dynamics -- see pdplqr_solver::padded)
    const int nu = s - nx, su = nxu + nuu;
    auto umap = [&](int i) { return i < nu ? (i < nuu ? i : -1) : (i - nu < nxu ? nuu + (i - nu) : -1); };
    // record order (common.cuh): row positions of [E c] and positions of the w-indices; identity on the thread path and
    // for the sizes without the warp kernel's layout
    const bool lay = !sym && warp_layout(nx, nu);
    const int nH = sym ? s * (s + 1) / 2 : s * s;
    const int oC = nx * s, oH = oC + nx, oh = oH + nH, oend = oh + s;
    const long long npad = sym ? ((nprob + 31) / 32) * 32 : nprob;
    const long long total = npad * N * REC;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long b, k;
        int e;
        if (!sym) {
            const long long st = idx / REC;
            e = (int)(idx - st * REC);
            b = st / N; k = st - b * N;
        } else {   // idx = ((tile*N + k)*REC + (e & ~1))*32 + lane*2 + (e & 1)
            const long long blk = idx / (32LL * REC);
            const int r = (int)(idx - blk * 32LL * REC);
            const int pair = r / 64, lane = (r % 64) / 2;
            e = pair * 2 + (r & 1);
            const long long tile = blk / N;
            k = blk - tile * N;
            b = tile * 32 + lane;
            if (b >= nprob) b = nprob - 1;
        }
        const long long st = b * N + k;
        double v = 0.0;
        if (e < oC) {
            const int i = erow_inv(e % nx, nx, lay), ju = umap(widx_inv(e / nx, nx, nu, lay));
            if (i < nxu && ju >= 0) v = E[st * (long long)(nxu * su) + i + ju * nxu];
        } else if (e < oH) {
            const int i = erow_inv(e - oC, nx, lay);
            if (i < nxu) v = c[st * nxu + i];
        } else if (e < oh) {
            int q = e - oH;
            if (sym) {  // q -> (i, j), i >= j, columns packed one after the other
                int j = 0;
                while (q >= s - j) { q -= s - j; ++j; }
                q = (j + q) + j * s;
            } else if (h_rotated(s)) {   // SegDims::h_off: rows of column j rotated by 4 (j / 2)
                const int di = q % s, j = q / s;
                q = (((di - 4 * (j >> 1)) % s + s) % s) + j * s;
            }
            const int iu = umap(widx_inv(q % s, nx, nu, lay)), ju = umap(widx_inv(q / s, nx, nu, lay));
            // the record's H is exactly symmetric: both halves come from the caller's LOWER triangle, the one the reference's
            // LLT of M reads (lqr_kernel.hpp:121-126); the kernels may then read H(i,j) as H(j,i)
            if (iu >= 0 && ju >= 0) v = H[st * (long long)(su * su) + (iu > ju ? iu : ju) + (iu > ju ? ju : iu) * su];
            else v = (q % s == q / s) ? 1.0 : 0.0;
        } else if (e < oend) {
            const int iu = umap(widx_inv(e - oh, nx, nu, lay));
            if (iu >= 0) v = hv[st * su + iu];
        }
        rec[idx] = v;
    }
}

// flat D (reference layout, stage chunks back to back, nc_k x dim_user column-major) -> device copy in kernel dimensions
// (zero columns for padded inputs / states) with every stage chunk padded to 16 bytes
__global__ void pad_D_kernel(const double* __restrict__ src, double* __restrict__ dst, const long long* __restrict__ soff,
                             const long long* __restrict__ doff, const int* __restrict__ ncs, long long s_total,
                             long long d_total, int nstages, int nx, int nu, int nxu, int nuu) {
    const int b = blockIdx.x / nstages, k = blockIdx.x % nstages;   // 1-D grid: grid.y is capped at 65535
    const long long nd = doff[k + 1] - doff[k];
    const int nc = ncs[k];
    const bool term = (k == nstages - 1);                            // terminal stage: columns are states only
    const double* s = src + (long long)b * s_total + soff[k];
    double* d = dst + (long long)b * d_total + doff[k];
    const int dim = term ? nx : nx + nu;
    for (long long e = threadIdx.x; e < nd; e += blockDim.x) {
        double v = 0.0;
        if (nc > 0 && e < (long long)nc * dim) {
            const int r = (int)(e % nc), j = (int)(e / nc);
            int ju;
            if (term) ju = j < nxu ? j : -1;
            else ju = j < nu ? (j < nuu ? j : -1) : (j - nu < nxu ? nuu + (j - nu) : -1);
            if (ju >= 0) v = s[r + (long long)ju * nc];
        }
        d[e] = v;
    }
}

// layout conversion of trajectories [batch][N (nu + nx) + nx] between caller and kernel dimensions (either direction:
// components the source does not have are zero, components the destination does not have are dropped)
__global__ void repack_ws_kernel(const double* __restrict__ src, double* __restrict__ dst, long long batch, int N,
                                 int nx_s, int nu_s, int nx_d, int nu_d) {
    const int s_s = nx_s + nu_s, s_d = nx_d + nu_d;
    const long long wl_s = (long long)N * s_s + nx_s, wl_d = (long long)N * s_d + nx_d, total = batch * wl_d;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / wl_d, r = idx - b * wl_d;
        const long long k = r / s_d < N ? r / s_d : N;
        const int i = (int)(r - k * s_d);
        int is;
        if (k == N) is = i < nx_s ? i : -1;
        else is = i < nu_d ? (i < nu_s ? i : -1) : (i - nu_d < nx_s ? nu_s + (i - nu_d) : -1);
        dst[idx] = is >= 0 ? src[b * wl_s + k * s_s + is] : 0.0;
    }
}
// same for arrays of vectors / square blocks: [items][n_s (x n_s)] -> [items][n_d (x n_d)], `fill` on the padded diagonal
__global__ void repack_vec_kernel(const double* __restrict__ src, double* __restrict__ dst, long long items, int n_s,
                                  int n_d, int square, double fill) {
    const long long per_d = square ? (long long)n_d * n_d : n_d, per_s = square ? (long long)n_s * n_s : n_s;
    const long long total = items * per_d;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long it = idx / per_d;
        const int e = (int)(idx - it * per_d);
        const int i = square ? e % n_d : e, j = square ? e / n_d : 0;
        double v = (square && i == j) ? fill : 0.0;
        if (i < n_s && j < (square ? n_s : 1)) v = src[it * per_s + i + (long long)j * n_s];
        dst[idx] = v;
    }
}

// factor records -> gains in the caller's layout for problems [b0, b0 + nb): K [nb][N][nuu x nxu], d [nb][N][nuu],
// Gt [nb][N][nuu x nxu] (zero in the last segment, which ends at the terminal cost, and on the thread-per-problem path)
__global__ void unpack_gains_kernel(const double* __restrict__ fac, double* __restrict__ K, double* __restrict__ d,
                                    double* __restrict__ Gt, int b0, int nb, int N, int nx, int nu, int nxu, int nuu,
                                    int FREC, int tiled, int last_start) {
    const int nK = nuu * nxu, per = 2 * nK + nuu;
    const long long total = (long long)nb * N * per;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long st = idx / per;
        const int e = (int)(idx - st * per);
        const long long bl = st / N, b = b0 + bl;
        const int k = (int)(st - bl * N);
        int which, q;                               // 0: K, 1: d, 2: Gt ; q = index inside the caller's block
        if (e < nK) { which = 0; q = e; } else if (e < nK + nuu) { which = 1; q = e - nK; } else { which = 2; q = e - nK - nuu; }
        const int i = which == 1 ? q : q % nuu, j = which == 1 ? 0 : q / nuu;
        const int src = which == 0 ? i + j * nu : (which == 1 ? nu * nx + i : nu * (nx + 1) + i + j * nu);
        double v = 0.0;
        if (which == 2 && (tiled || k >= last_start)) v = 0.0;
        else if (tiled) v = fac[((b / 32) * N + k) * (long long)FREC * 32 + (src & ~1) * 32 + (b % 32) * 2 + (src & 1)];
        else v = fac[(b * N + k) * (long long)FREC + src];
        (which == 0 ? K + st * nK : (which == 1 ? d + st * nuu : Gt + st * nK))[q] = v;
    }
}

// one block per (stage, problem): for every constraint row find its non-zeros; rows with at most one non-zero are
// recorded as (column, value) (column -1 for an all-zero row), any other row raises the flag (-> dense path)
__global__ void detect_selection_kernel(const double* __restrict__ D, const long long* __restrict__ doff,
                                        const long long* __restrict__ coff, const int* __restrict__ ncs, long long d_total,
                                        long long nc_total, int N, int nx, int s, int* __restrict__ col,
                                        double* __restrict__ val, int* flag) {
    const int k = blockIdx.x % (N + 1), b = blockIdx.x / (N + 1);   // 1-D grid: grid.y is capped at 65535
    const int nc = ncs[k], dim = (k < N) ? s : nx;
    const double* Dk = D + (long long)b * d_total + doff[k];
    for (int r = threadIdx.x; r < nc; r += blockDim.x) {
        int nnz = 0, cj = -1;
        double v = 0.0;
        for (int j = 0; j < dim; ++j) {
            const double d = Dk[r + (long long)j * nc];
            if (d != 0.0) { ++nnz; cj = j; v = d; }
        }
        col[(long long)b * nc_total + coff[k] + r] = cj;
        val[(long long)b * nc_total + coff[k] + r] = v;
        if (nnz > 1) atomicExch(flag, 1);
    }
}

template <class Tp>
int dev_alloc(Solver& h, Tp** p, size_t count) {
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(Tp);
    if (h.guards) {   // [guard | payload (0xFF) | guard], see Solver::guards
        constexpr size_t G = Solver::GUARD_BYTES;
        CU_TRY(&h, cudaMalloc(&q, bytes + 2 * G));
        h.owned.push_back(q);
        char* base = static_cast<char*>(q);
        CU_TRY(&h, cudaMemset(base, 0xA5, G));
        CU_TRY(&h, cudaMemset(base + G, 0xFF, bytes));
        CU_TRY(&h, cudaMemset(base + G + bytes, 0xA5, G));
        h.guarded.push_back({base, bytes});
        *p = reinterpret_cast<Tp*>(base + G);
        return PDPLQR_OK;
    }
    CU_TRY(&h, cudaMalloc(&q, bytes));
    h.owned.push_back(q);
    *p = static_cast<Tp*>(q);
    return PDPLQR_OK;
}

// grid for an element-wise helper kernel
inline int ew_blocks(long long total) { return (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, 148LL * 32)); }

// caller-layout trajectory -> kernel layout (or back): only used by padded handles
int repack_ws(Solver& h, const double* src, double* dst, bool to_kernel) {
    h.chain_tail = false;
    const long long total = (long long)h.batch * ((long long)h.N * (to_kernel ? h.s : h.su) + (to_kernel ? h.nx : h.nxu));
    if (to_kernel) repack_ws_kernel<<<ew_blocks(total), 256, 0, h.stream>>>(src, dst, h.batch, h.N, h.nxu, h.nuu, h.nx, h.nu);
    else repack_ws_kernel<<<ew_blocks(total), 256, 0, h.stream>>>(src, dst, h.batch, h.N, h.nx, h.nu, h.nxu, h.nuu);
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}
// [items][n_src (x n_src)] -> [items][n_dst (x n_dst)]
int repack_vec(Solver& h, const double* src, double* dst, long long items, int n_src, int n_dst, bool square = false,
               double fill = 0.0) {
    h.chain_tail = false;
    const long long total = items * (square ? (long long)n_dst * n_dst : n_dst);
    repack_vec_kernel<<<ew_blocks(total), 256, 0, h.stream>>>(src, dst, items, n_src, n_dst, square ? 1 : 0, fill);
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}

int set_model_common(Solver& h, const double* E, const double* c, const double* H, const double* hv,
                     const double* HN, const double* hN, const double* Dflat, bool on_device) {
    if (h.nc_total > 0 && !Dflat) return fail(&h, PDPLQR_ERR_INVALID, "set_model: D is required when ncs has non-zero entries");
    const size_t nst = (size_t)h.batch * h.N, B = h.batch;
    const size_t nE = nst * h.nxu * h.su, nc = nst * h.nxu, nH = nst * h.su * h.su, nh = nst * h.su;   // caller's layout
    const size_t nHN = B * h.nxu * h.nxu, nhN = B * h.nxu;
    const double *dE = E, *dc = c, *dH = H, *dh = hv, *dHN = HN, *dhN = hN;
    double* stage = nullptr;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (!on_device) {
        CU_TRY(&h, cudaMalloc((void**)&stage, (nE + nc + nH + nh + nHN + nhN) * sizeof(double)));
        cudaError_t e1 = cudaMemcpyAsync(stage, E, nE * 8, kind, h.stream);
        cudaError_t e2 = cudaMemcpyAsync(stage + nE, c, nc * 8, kind, h.stream);
        cudaError_t e3 = cudaMemcpyAsync(stage + nE + nc, H, nH * 8, kind, h.stream);
        cudaError_t e4 = cudaMemcpyAsync(stage + nE + nc + nH, hv, nh * 8, kind, h.stream);
        cudaError_t e5 = cudaMemcpyAsync(stage + nE + nc + nH + nh, HN, nHN * 8, kind, h.stream);
        cudaError_t e6 = cudaMemcpyAsync(stage + nE + nc + nH + nh + nHN, hN, nhN * 8, kind, h.stream);
        for (cudaError_t e : {e1, e2, e3, e4, e5, e6})
            if (e != cudaSuccess) {
                cudaFree(stage);
                return fail(&h, PDPLQR_ERR_CUDA, std::string("set_model: model upload failed: ") + cudaGetErrorString(e));
            }
        dE = stage; dc = stage + nE; dH = stage + nE + nc; dh = stage + nE + nc + nH;
        dHN = dh + nh; dhN = dHN + nHN;
    }
    const long long total = (long long)(h.thread_path ? ((h.batch + 31) / 32) * 32 : h.batch) * h.N * h.mrec;
    pack_model_kernel<<<ew_blocks(total), 256, 0, h.stream>>>(dE, dc, dH, dh, h.d_model, (long long)h.batch, h.N, h.nx, h.s,
                                                              h.mrec, h.thread_path ? 1 : 0, h.nxu, h.nuu);
    h.launches++;
    cudaError_t e = cudaGetLastError();
    // terminal cost; padded states get a unit diagonal (they are zero, the value function stays positive definite)
    if (e == cudaSuccess && repack_vec(h, dHN, h.d_HN, (long long)B, h.nxu, h.nx, true, 1.0)) e = cudaErrorUnknown;
    if (e == cudaSuccess && repack_vec(h, dhN, h.d_hN, (long long)B, h.nxu, h.nx)) e = cudaErrorUnknown;
    double* dstage = nullptr;
    if (e == cudaSuccess && h.nc_total > 0) {
        const double* dsrc = Dflat;
        if (!on_device) {
            e = cudaMalloc((void**)&dstage, (size_t)h.batch * h.d_total_host * 8);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dstage, Dflat, (size_t)h.batch * h.d_total_host * 8, kind, h.stream);
            dsrc = dstage;
        }
        const unsigned grid = (unsigned)((h.N + 1) * (long long)h.batch);   // 1-D: grid.y is capped at 65535
        if (e == cudaSuccess) {
            pad_D_kernel<<<grid, 128, 0, h.stream>>>(dsrc, h.d_D, h.d_doff_host, h.d_doff, h.d_ncs, h.d_total_host,
                                                     h.d_total_dev, h.N + 1, h.nx, h.nu, h.nxu, h.nuu);
            h.launches++;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) {   // structure detection: selection-matrix constraints need no dense D in the kernels
            e = cudaMemsetAsync(h.d_sel_flag, 0, sizeof(int), h.stream);
            detect_selection_kernel<<<grid, 64, 0, h.stream>>>(h.d_D, h.d_doff, h.d_coff, h.d_ncs, h.d_total_dev, h.nc_total, h.N,
                                                               h.nx, h.s, h.d_sel_col, h.d_sel_val, h.d_sel_flag);
            h.launches++;
            if (e == cudaSuccess) e = cudaGetLastError();
            int flag = 1;
            if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, h.d_sel_flag, sizeof(int), cudaMemcpyDeviceToHost, h.stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h.stream);
            h.sel_mode = (e == cudaSuccess) && flag == 0 && h.allow_sel;
        }
    }
    if (e == cudaSuccess && (stage || dstage)) e = cudaStreamSynchronize(h.stream);
    if (stage) cudaFree(stage);
    if (dstage) cudaFree(dstage);
    if (e != cudaSuccess) return fail(&h, PDPLQR_ERR_CUDA, std::string("set_model: ") + cudaGetErrorString(e));
    h.model_set = true;
    h.factorized = false;
    if (h.admm_exec) { cudaGraphExecDestroy(h.admm_exec); h.admm_exec = nullptr; }   // sel_mode etc. are baked into the graph
    if (h.solve_exec) { cudaGraphExecDestroy(h.solve_exec); h.solve_exec = nullptr; }
    return PDPLQR_OK;
}

TreeTopParams top_params(Solver& h, const double* d_x0, bool affine_only) {
    TreeTopParams tp{};
    tp.batch = h.batch; tp.x0 = d_x0; tp.lam0 = nullptr; tp.affine_only = affine_only ? 1 : 0;
    for (TreeLevel& lv : h.levels)
        if (lv.in_top) {
            const int i = tp.nlevels++;
            tp.count[i] = lv.count; tp.sum[i] = lv.sum; tp.dd[i] = lv.dd; tp.x[i] = lv.x; tp.lam[i] = lv.lam;
        }
    tp.width = tp.count[0]; tp.ngroups = 1; tp.is_root = 1;
    return tp;
}
// one latency-mode launch: levels[g.l0] .. levels[g.l1], one CTA per block of g.width nodes of level l0
TreeTopParams sub_params(Solver& h, const pdplqr_solver::LatGroup& g, const double* d_x0, const double* d_lam0,
                         bool affine_only) {
    TreeTopParams tp{};
    tp.batch = h.batch; tp.x0 = d_x0; tp.lam0 = d_lam0; tp.affine_only = affine_only ? 1 : 0;
    for (int l = g.l0; l <= g.l1; ++l) {
        TreeLevel& lv = h.levels[l];
        const int i = tp.nlevels++;
        tp.count[i] = lv.count; tp.sum[i] = lv.sum; tp.dd[i] = lv.dd; tp.x[i] = lv.x; tp.lam[i] = lv.lam;
    }
    tp.width = g.width;
    tp.ngroups = (tp.count[0] + g.width - 1) / g.width;
    tp.is_root = (g.l1 == (int)h.levels.size() - 1) ? 1 : 0;
    tp.tt_cap = h.lat_tt_cap > 0 ? h.lat_tt_cap : h.ops->lat_tt_cap;
    return tp;
}

// up-sweep of the interface tree: one launch per lower level, then one launch for all upper (binary) levels
int run_tree_up(Solver& h, bool affine_only) {
    for (size_t l = 0; l < h.levels.size(); ++l) {
        TreeLevel& lv = h.levels[l];
        if (lv.in_top) break;
        TreeParams tp{};
        tp.batch = h.batch; tp.count = lv.count; tp.R = lv.R; tp.groups = lv.groups;
        tp.sum_in = lv.sum; tp.dd = lv.dd;
        tp.sum_out = h.levels[l + 1].sum;
        int rc = affine_only ? h.ops->tree_up_affine(h, tp) : h.ops->tree_up(h, tp);
        if (rc) return rc;
    }
    if (h.top_lat) {
        for (const auto& g : h.lat_groups) {
            if (g.l1 == g.l0) continue;   // a lone root: nothing to combine
            int rc = h.ops->tree_sub_up(h, sub_params(h, g, nullptr, nullptr, affine_only));
            if (rc) return rc;
        }
        return PDPLQR_OK;
    }
    return h.ops->tree_top_up(h, top_params(h, nullptr, affine_only));
}

int run_backward(Solver& h) {
    if (!h.fused) h.chain_tail = false;   // work the caller enqueued since the last API call is unknown (launch_chain)
    if (!h.model_set) return fail(&h, PDPLQR_ERR_ORDER, "backward before set_model");
    if (!h.updated) return fail(&h, PDPLQR_ERR_ORDER, "backward before update_problem_data (lqr_solver_parallel.hpp:115)");
    int rc = h.ops->backward(h);
    if (rc) return rc;
    if (h.S > 1) {
        rc = run_tree_up(h, false);
        if (rc) return rc;
    }
    h.updated = false;  // backward consumes the staged data (in-place accumulation in the reference)
    h.factorized = true;
    h.backward_done = true;
    return PDPLQR_OK;
}

// backward_without_factorization: affine-only sweeps with the cached factors.  When the affine cache is not kept
// (thread-per-problem path, or PDPLQR_OPT_AFFINE_CACHE = 0) the full factorising sweep is run instead -- same
// result, since only affine data may have changed between the two calls.
int run_backward_nofact(Solver& h) {
    if (!h.fused) h.chain_tail = false;   // work the caller enqueued since the last API call is unknown (launch_chain)
    if (!h.factorized)
        return fail(&h, PDPLQR_ERR_ORDER, "backward_without_factorization before any backward (lqr_solver_parallel.hpp:148)");
    if (!h.keep_affine || h.thread_path) return run_backward(h);
    if (!h.updated) return fail(&h, PDPLQR_ERR_ORDER, "backward_without_factorization before update_problem_data");
    int rc = h.ops->affine(h);
    if (rc) return rc;
    if (h.S > 1) {
        rc = run_tree_up(h, true);
        if (rc) return rc;
    }
    h.updated = false;
    h.backward_done = true;
    return PDPLQR_OK;
}

// down-sweep of the interface tree: the upper levels in one launch, then one launch per lower level
int run_tree_down(Solver& h, const double* d_x0, const double* d_lam0) {
    int rc = PDPLQR_OK;
    if (h.top_lat) {
        for (int i = (int)h.lat_groups.size() - 1; i >= 0; --i) {
            rc = h.ops->tree_sub_down(h, sub_params(h, h.lat_groups[i], d_x0, d_lam0, false));
            if (rc) return rc;
        }
    } else {
        TreeTopParams ttp = top_params(h, d_x0, false);
        ttp.lam0 = d_lam0;
        rc = h.ops->tree_top_down(h, ttp);
        if (rc) return rc;
    }
    for (int l = (int)h.levels.size() - 1; l >= 0; --l) {
        TreeLevel& lv = h.levels[l];
        if (lv.in_top) continue;
        TreeParams tp{};
        tp.batch = h.batch; tp.count = lv.count; tp.R = lv.R; tp.groups = lv.groups;
        tp.dd = lv.dd;
        tp.x_parent = h.levels[l + 1].x;
        tp.lam_parent = h.levels[l + 1].lam;
        tp.x_node = lv.x; tp.lam_node = lv.lam;
        rc = h.ops->tree_down(h, tp);
        if (rc) return rc;
    }
    return PDPLQR_OK;
}

int run_forward(Solver& h, const double* d_x0, double* d_ws_out) {
    if (!h.fused) h.chain_tail = false;   // work the caller enqueued since the last API call is unknown (launch_chain)
    if (!h.backward_done) return fail(&h, PDPLQR_ERR_ORDER, "forward before backward (one forward per backward)");
    if (h.interior && !h.root_fresh)
        return fail(&h, PDPLQR_ERR_ORDER, "forward on an interior horizon shard needs a fresh pdplqr_set_root_boundary_device");
    // a root boundary is consumed by the forward that follows it: a handle reused without refreshing the boundary
    // rolls out from the caller's x0 again (have_root = "the last forward used the root boundary", read by the costates)
    h.have_root = h.root_fresh;
    h.root_fresh = false;
    if (h.have_root) d_x0 = h.d_root_x;
    if (h.S > 1) {
        int rc = run_tree_down(h, d_x0, h.have_root ? h.d_root_lam : nullptr);
        if (rc) return rc;
    }
    int rc = h.ops->forward(h, d_x0, d_ws_out);
    if (rc) return rc;
    h.backward_done = false;
    return PDPLQR_OK;
}

// forward with caller-layout device arrays (x0 [batch][nxu], ws_out [batch][N su + nxu])
int run_forward_user(Solver& h, const double* d_x0, double* d_ws_out) {
    if (!h.padded) return run_forward(h, d_x0, d_ws_out);
    int rc = repack_vec(h, d_x0, h.d_x0p, h.batch, h.nxu, h.nx);
    if (rc) return rc;
    rc = run_forward(h, h.d_x0p, h.d_wsp_out);
    if (rc) return rc;
    return repack_ws(h, h.d_wsp_out, d_ws_out, false);
}

}  // namespace

// =====================================================================================================
extern "C" {

int pdplqr_version(void) { return 200; }

// (problem, segment) groups one GPU keeps resident in the throughput-mode stage sweep = SMs x CTAs per SM of that kernel
// (occupancy calculator).  Segment counts that are whole multiples of it avoid a trailing partial wave.
int pdplqr_wave_size(int nx, int nu, int device) {
    const Ops* ops = find_ops(nx, nu);
    if (!ops) ops = find_padded_ops(nx, nu);
    if (!ops) return PDPLQR_ERR_UNSUPPORTED;
    int sms = 0;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
        return PDPLQR_ERR_CUDA;
    const int per_sm = ops->wave(0, false);
    return per_sm > 0 ? sms * per_sm : PDPLQR_ERR_CUDA;
}

int pdplqr_create(pdplqr_handle_t* out, int nx, int nu, int N, const int* ncs, int batch, int num_segments,
                  int load_balancing, int condensed_type, int device) {
    if (!out) return PDPLQR_ERR_INVALID;
    *out = nullptr;
    if (nx < 1 || nu < 1 || N < 1 || batch < 1 || num_segments < 0) return PDPLQR_ERR_INVALID;  // lqr_model.hpp:75-77
    if (condensed_type != PDPLQR_CONDENSED_LU && condensed_type != PDPLQR_CONDENSED_CHOLESKY)
        return PDPLQR_ERR_INVALID;  // lqr_solver_parallel.hpp:98-99
    const int nxu = nx, nuu = nu;
    const Ops* ops = find_ops(nx, nu);
    if (!ops) {   // not an instantiated pair: embed in the cheapest one that contains it (the reference takes any n, m:
                  // lqr_model.hpp:66-89)
        ops = find_padded_ops(nx, nu);
        if (!ops) {
            g_create_error = "pdplqr_create: (nx, nu) exceeds the largest instantiated kernel size (csrc/inst_list.h)";
            return PDPLQR_ERR_UNSUPPORTED;
        }
        nx = ops->nx; nu = ops->nu;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) return PDPLQR_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return PDPLQR_ERR_CUDA;

    Solver* h = new Solver();
    h->nx = nx; h->nu = nu; h->N = N; h->batch = batch; h->s = nx + nu; h->device = device;
    h->nxu = nxu; h->nuu = nuu; h->su = nxu + nuu; h->padded = (nxu != nx || nuu != nu);
    h->load_balancing = load_balancing == 1; h->condensed_type = condensed_type; h->ops = ops;
    h->ncs.assign(N + 1, 0);
    if (ncs) h->ncs.assign(ncs, ncs + N + 1);
    for (int v : h->ncs) {
        if (v < 0) { delete h; return PDPLQR_ERR_INVALID; }
        h->nc_total += v;
    }
    h->coff.assign(N + 2, 0); h->doff_host.assign(N + 2, 0); h->doff_dev.assign(N + 2, 0);
    for (int k = 0; k <= N; ++k) {
        const long long dim = (k < N) ? nx + nu : nx, dim_user = (k < N) ? nxu + nuu : nxu;
        h->coff[k + 1] = h->coff[k] + h->ncs[k];
        h->doff_host[k + 1] = h->doff_host[k] + h->ncs[k] * dim_user;
        h->doff_dev[k + 1] = h->doff_dev[k] + ((h->ncs[k] * dim + 1) & ~1LL);
        h->ncmax = std::max(h->ncmax, h->ncs[k]);
    }
    h->d_total_host = h->doff_host[N + 1];
    h->d_total_dev = h->doff_dev[N + 1];
    h->keep_affine = h->nc_total > 0;

    // ---- segmentation (lqr_solver_parallel.hpp:70-80)
    int S = num_segments;
    const bool auto_seg = (S == 0);
    const bool equal_split = auto_seg || load_balancing == 2;   // GPU-style partition: equal lengths
    if (auto_seg) {
        // GPU-appropriate default: one full wave of (problem, segment) groups (SMs x resident stage-kernel CTAs per SM,
        // from the occupancy calculator), segments >= 8 stages
        const int target_groups = std::max(1, pdplqr_wave_size(nx, nu, device));
        S = std::max(1, std::min(N / 8, (target_groups + batch - 1) / batch));
    }
    S = std::min(S, N);
    h->S = S;
    const double scale = h->load_balancing ? 1.55 : 1.0;
    h->seg_start.resize(S); h->seg_len.resize(S);
    for (int i = 0; i < S; ++i) {
        const int st = (i == 0) ? 0 : h->seg_start[i - 1] + h->seg_len[i - 1];
        int len;
        if (equal_split) len = N / S + (i < N % S ? 1 : 0);  // equal lengths (the 1.55 rule targets <= #cores segments)
        else len = (i < S - 1) ? int(N / (scale + S - 1)) : N - st;
        if (i < S - 1 && len < 1) len = 1;
        h->seg_start[i] = st; h->seg_len[i] = len;
    }
    if (h->seg_len[S - 1] < 1) { delete h; return PDPLQR_ERR_INVALID; }
    h->seg_mode = equal_split ? 0 : 1;
    h->seg_len0 = h->seg_len[0];
    if (!equal_split && S > 1) {   // the reference rule must be expressible in closed form (it is unless len was clamped)
        for (int i = 0; i < S - 1; ++i)
            if (h->seg_len[i] != h->seg_len0) { delete h; return PDPLQR_ERR_INVALID; }
    }
    h->thread_path = (S == 1) && ops->has_thread_path && h->nc_total == 0;   // (interior shards switch it off)
    h->frec = h->thread_path ? ops->FRECT : ops->FREC;
    h->mrec = h->thread_path ? ops->TREC : ops->REC;
    if (const char* e = getenv("PDPLQR_DEBUG_GUARDS")) h->guards = atoi(e) != 0;
    if (const char* e = getenv("PDPLQR_BWD_VARIANT")) h->bwd_variant = atoi(e);
    if (const char* e = getenv("PDPLQR_FWD_VARIANT")) h->fwd_variant = atoi(e);
    if (const char* e = getenv("PDPLQR_LAT_THREADS")) h->lat_threads = atoi(e);
    if (const char* e = getenv("PDPLQR_TREE_LAT")) h->tree_lat = atoi(e);
    if (const char* e = getenv("PDPLQR_TREE_LAT_MAX")) h->tree_lat_max = atoi(e);
    if (const char* e = getenv("PDPLQR_TREE_LAT_WIDTH")) h->lat_width = atoi(e);
    if (const char* e = getenv("PDPLQR_TREE_LAT_TT")) h->lat_tt_cap = atoi(e);
    if (const char* e = getenv("PDPLQR_SEG_T")) h->seg_t = atoi(e);
    if (const char* e = getenv("PDPLQR_WARP_KERNEL")) h->warp_kernel = atoi(e);
    if (const char* e = getenv("PDPLQR_ADMM_GRAPH")) h->admm_use_graph = atoi(e);
    if (const char* e = getenv("PDPLQR_SOLVE_GRAPH")) h->solve_use_graph = atoi(e);
    if (const char* e = getenv("PDPLQR_SPARSE_D")) h->allow_sel = atoi(e);
    if (const char* e = getenv("PDPLQR_PIPELINE_CHUNKS")) h->pipeline_chunks = std::max(1, std::min(64, atoi(e)));
    // Programmatic dependent launch of the solve chain: OFF unless PDPLQR_PDL=1 (experimental, see launch_chain in
    // solver_impl.cuh: one edge of the chain is not safe behind asynchronous H2D copies, and the graph-launched solve gains
    // only 1.3 us from it).
    h->use_pdl = 0;
    if (const char* e = getenv("PDPLQR_PDL")) h->use_pdl = atoi(e) != 0;
    if (const char* e = getenv("PDPLQR_PDL_MASK")) h->pdl_mask = atoi(e);

    auto bail = [&](int rc) {   // keep the error text: the handle does not survive
        g_create_error = h->err.empty() ? std::string("pdplqr_create: ") + cudaGetErrorString(cudaGetLastError()) : h->err;
        pdplqr_destroy(h);
        return rc;
    };
#define CREATE_TRY(expr)                                                                              \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            h->err = std::string("pdplqr_create: " #expr ": ") + cudaGetErrorString(_e);              \
            return bail(PDPLQR_ERR_CUDA);                                                             \
        }                                                                                             \
    } while (0)
    CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    const size_t B = batch, ws_len = (size_t)N * h->s + nx;
    int rc = 0;
    // thread-per-problem handles keep whole tiles of 32 problems; everything else exactly `batch` records (allocating the
    // padded count for every handle cost 32 x the model memory of a single long problem: 128 GB at N = 2^20)
    const size_t Bpad = h->thread_path ? ((B + 31) / 32) * 32 : B;
    rc |= dev_alloc(*h, &h->d_model, Bpad * N * ops->REC);
    rc |= dev_alloc(*h, &h->d_HN, B * nx * nx);
    rc |= dev_alloc(*h, &h->d_hN, B * nx);
    rc |= dev_alloc(*h, &h->d_fac, Bpad * N * ops->FREC);
    rc |= dev_alloc(*h, &h->d_sum, B * S * ops->SREC);
    rc |= dev_alloc(*h, &h->d_xhat, B * S * nx);
    rc |= dev_alloc(*h, &h->d_uhat, B * S * nx);
    rc |= dev_alloc(*h, &h->d_ws_in, ws_len * B);
    rc |= dev_alloc(*h, &h->d_ws_out, ws_len * B);
    rc |= dev_alloc(*h, &h->d_x0, B * nx);
    rc |= dev_alloc(*h, &h->d_seg_start, S);
    rc |= dev_alloc(*h, &h->d_seg_len, S);
    rc |= dev_alloc(*h, &h->d_status, B * S);
    if (h->padded) {
        rc |= dev_alloc(*h, &h->d_wsp_in, ws_len * B);
        rc |= dev_alloc(*h, &h->d_wsp_out, ws_len * B);
        rc |= dev_alloc(*h, &h->d_x0p, B * nx);
    }
    if (h->nc_total > 0) {
        rc |= dev_alloc(*h, &h->d_ncs, N + 1);
        rc |= dev_alloc(*h, &h->d_coff, N + 2);
        rc |= dev_alloc(*h, &h->d_doff, N + 2);
        rc |= dev_alloc(*h, &h->d_doff_host, N + 2);
        rc |= dev_alloc(*h, &h->d_D, B * h->d_total_dev);
        rc |= dev_alloc(*h, &h->d_ys, B * h->nc_total);
        rc |= dev_alloc(*h, &h->d_zs, B * h->nc_total);
        rc |= dev_alloc(*h, &h->d_rho, B * h->nc_total);
        rc |= dev_alloc(*h, &h->d_inv_rho, B * h->nc_total);
        rc |= dev_alloc(*h, &h->d_sel_col, B * h->nc_total);
        rc |= dev_alloc(*h, &h->d_sel_val, B * h->nc_total);
        rc |= dev_alloc(*h, &h->d_sel_flag, 1);
    }
    if (h->keep_affine) rc |= dev_alloc(*h, &h->d_aff, B * N * ops->AREC);
    if (rc) return bail(PDPLQR_ERR_CUDA);   // (dev_alloc left the CUDA error text in h->err)
    if (h->nc_total > 0) {
        CREATE_TRY(cudaMemcpy(h->d_ncs, h->ncs.data(), sizeof(int) * (N + 1), cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMemcpy(h->d_coff, h->coff.data(), sizeof(long long) * (N + 2), cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMemcpy(h->d_doff, h->doff_dev.data(), sizeof(long long) * (N + 2), cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMemcpy(h->d_doff_host, h->doff_host.data(), sizeof(long long) * (N + 2), cudaMemcpyHostToDevice));
    }
    CREATE_TRY(cudaMemcpy(h->d_seg_start, h->seg_start.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(h->d_seg_len, h->seg_len.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemset(h->d_fac, 0, Bpad * N * ops->FREC * sizeof(double)));
    CREATE_TRY(cudaMemset(h->d_uhat, 0, B * S * nx * sizeof(double)));
    CREATE_TRY(cudaMemset(h->d_status, 0, B * S * sizeof(int)));
    // the thread-per-problem kernels write only P, p of the summary: F, C, f of a one-segment terminal slice are zero
    CREATE_TRY(cudaMemset(h->d_sum, 0, B * S * ops->SREC * sizeof(double)));

    // ---- interface tree plan: lower levels (one launch each, fan-in 4) until <= 32 nodes, then binary levels
    //      that all run inside one launch (tree_top_*_kernel); the last level is the root (1 node)
    if (S > 1) {
        int cnt = S;
        // Few problems (latency mode): fan-in-4 throughput levels only while a level would still need more CTAs than
        // tree_lat_max, then binary levels, grouped from the top into launches of <= log2(top_lat_nodes) levels that
        // each reduce blocks of nodes inside one CTA.  Otherwise: fan-in-4 levels until <= 32 nodes, then the binary
        // upper tree in one launch of one CTA per problem.
        int W = ops->top_lat_nodes;
        if (h->lat_width >= 2 && h->lat_width < W) W = h->lat_width;
        h->top_lat = h->tree_lat && W >= 2 && B <= (size_t)h->tree_lat_max;
        auto add_level = [&](int count, int R, bool in_top) {
            TreeLevel lv{};
            lv.count = count; lv.R = R; lv.in_top = in_top;
            lv.groups = (count + R - 1) / R;
            if (h->levels.empty()) { lv.sum = h->d_sum; lv.x = h->d_xhat; lv.lam = h->d_uhat; }
            else {
                rc |= dev_alloc(*h, &lv.sum, B * count * ops->SREC);
                rc |= dev_alloc(*h, &lv.x, B * count * nx);
                rc |= dev_alloc(*h, &lv.lam, B * count * nx);
            }
            rc |= dev_alloc(*h, &lv.dd, B * count * ops->DREC);
            h->levels.push_back(lv);
        };
        if (h->top_lat) {
            while ((long long)B * ((cnt + W - 1) / W) > h->tree_lat_max) { add_level(cnt, 4, false); cnt = (cnt + 3) / 4; }
            const int first_bin = (int)h->levels.size();
            for (;; cnt = (cnt + 1) / 2) { add_level(cnt, 2, true); if (cnt == 1) break; }
            int logw = 0;
            while ((2 << logw) <= W) ++logw;                       // levels one CTA can climb
            int hi = (int)h->levels.size() - 1;
            std::vector<pdplqr_solver::LatGroup> groups;           // top -> bottom
            do {
                const int lo = std::max(first_bin, hi - logw);
                groups.push_back({lo, hi, 1 << (hi - lo)});
                hi = lo;
            } while (hi > first_bin);
            h->lat_groups.assign(groups.rbegin(), groups.rend());
        } else {
            for (;;) {
                const bool in_top = cnt <= TREE_TOP_MAX_NODES;
                add_level(cnt, in_top ? 2 : 4, in_top);
                if (cnt == 1) break;
                cnt = h->levels.back().groups;
            }
        }
        if (rc) return bail(PDPLQR_ERR_CUDA);
    }
    CREATE_TRY(cudaDeviceSynchronize());
#undef CREATE_TRY
    *out = h;
    return PDPLQR_OK;
}

int pdplqr_destroy(pdplqr_handle_t h) {
    if (!h) return PDPLQR_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->admm_exec) cudaGraphExecDestroy(h->admm_exec);
    if (h->admm_graph) cudaGraphDestroy(h->admm_graph);
    if (h->solve_exec) cudaGraphExecDestroy(h->solve_exec);
    for (void* p : h->owned) cudaFree(p);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    for (cudaEvent_t e : h->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_cmp) cudaEventDestroy(e);
    if (h->ev_ready) cudaEventDestroy(h->ev_ready);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return PDPLQR_OK;
}

int pdplqr_set_stream(pdplqr_handle_t h, void* cuda_stream) {
    if (!h) return PDPLQR_ERR_INVALID;
    if (h->own_stream && h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    h->own_stream = false;
    if (h->admm_exec) { cudaGraphExecDestroy(h->admm_exec); h->admm_exec = nullptr; }   // captured for the old stream
    if (h->solve_exec) { cudaGraphExecDestroy(h->solve_exec); h->solve_exec = nullptr; }
    return PDPLQR_OK;
}

int pdplqr_set_model(pdplqr_handle_t h, const double* E, const double* c, const double* H, const double* hvec,
                     const double* HN, const double* hN, const double* D) {
    if (!h || !E || !c || !H || !hvec || !HN || !hN) return fail(h, PDPLQR_ERR_INVALID, "set_model: null pointer");
    cudaSetDevice(h->device);
    return set_model_common(*h, E, c, H, hvec, HN, hN, D, false);
}
int pdplqr_set_model_device(pdplqr_handle_t h, const double* E, const double* c, const double* H,
                            const double* hvec, const double* HN, const double* hN, const double* D) {
    if (!h || !E || !c || !H || !hvec || !HN || !hN) return fail(h, PDPLQR_ERR_INVALID, "set_model: null pointer");
    cudaSetDevice(h->device);
    return set_model_common(*h, E, c, H, hvec, HN, hN, D, true);
}

int pdplqr_update_problem_data_device(pdplqr_handle_t h, const double* ws, const double* ys, const double* zs,
                                      const double* inv_rho, double sigma) {
    if (!h) return PDPLQR_ERR_INVALID;
    if (h->nc_total > 0 && (!ys || !zs || !inv_rho))
        return fail(h, PDPLQR_ERR_INVALID, "update_problem_data: ys, zs, inv_rho are required when the problem has constraints");
    if (h->padded && ws) {   // caller layout -> kernel layout (copied now: the caller's array may change afterwards)
        cudaSetDevice(h->device);
        int rc = repack_ws(*h, ws, h->d_wsp_in, true);
        if (rc) return rc;
        ws = h->d_wsp_in;
    }
    h->cur_ws = ws;
    h->cur_ys = ys; h->cur_zs = zs; h->cur_inv_rho = inv_rho;
    h->sigma = sigma;
    h->updated = true;
    return PDPLQR_OK;
}
int pdplqr_update_problem_data(pdplqr_handle_t h, const double* ws, const double* ys, const double* zs,
                               const double* inv_rho, double sigma) {
    if (!h) return PDPLQR_ERR_INVALID;
    cudaSetDevice(h->device);
    const size_t ws_len = (size_t)h->N * h->su + h->nxu;   // caller's layout
    if (ws) CU_TRY(h, cudaMemcpyAsync(h->d_ws_in, ws, ws_len * h->batch * 8, cudaMemcpyHostToDevice, h->stream));
    if (h->nc_total > 0) {
        if (!ys || !zs || !inv_rho)
            return fail(h, PDPLQR_ERR_INVALID, "update_problem_data: ys, zs, inv_rho are required when the problem has constraints");
        const size_t nb = (size_t)h->batch * h->nc_total * 8;
        CU_TRY(h, cudaMemcpyAsync(h->d_ys, ys, nb, cudaMemcpyHostToDevice, h->stream));
        CU_TRY(h, cudaMemcpyAsync(h->d_zs, zs, nb, cudaMemcpyHostToDevice, h->stream));
        CU_TRY(h, cudaMemcpyAsync(h->d_inv_rho, inv_rho, nb, cudaMemcpyHostToDevice, h->stream));
        ys = h->d_ys; zs = h->d_zs; inv_rho = h->d_inv_rho;
    }
    return pdplqr_update_problem_data_device(h, ws ? h->d_ws_in : nullptr, ys, zs, inv_rho, sigma);
}

static int stage_rho(pdplqr_handle_t h, const double* rho, const double** dev) {
    *dev = nullptr;
    if (h->nc_total == 0) return PDPLQR_OK;
    if (!rho) return fail(h, PDPLQR_ERR_INVALID, "backward: rho_vecs is required when the problem has constraints");
    CU_TRY(h, cudaMemcpyAsync(h->d_rho, rho, (size_t)h->batch * h->nc_total * 8, cudaMemcpyHostToDevice, h->stream));
    *dev = h->d_rho;
    return PDPLQR_OK;
}
int pdplqr_backward_device(pdplqr_handle_t h, const double* rho) {
    if (!h) return PDPLQR_ERR_INVALID;
    if (h->nc_total > 0 && !rho) return fail(h, PDPLQR_ERR_INVALID, "backward: rho_vecs is required when the problem has constraints");
    cudaSetDevice(h->device);
    h->cur_rho = rho;
    return run_backward(*h);
}
int pdplqr_backward(pdplqr_handle_t h, const double* rho) {
    if (!h) return PDPLQR_ERR_INVALID;
    cudaSetDevice(h->device);
    const double* dev = nullptr;
    int rc = stage_rho(h, rho, &dev);
    if (rc) return rc;
    return pdplqr_backward_device(h, dev);
}
int pdplqr_backward_without_factorization_device(pdplqr_handle_t h, const double* rho) {
    if (!h) return PDPLQR_ERR_INVALID;
    if (h->nc_total > 0 && !rho) return fail(h, PDPLQR_ERR_INVALID, "backward: rho_vecs is required when the problem has constraints");
    cudaSetDevice(h->device);
    h->cur_rho = rho;
    return run_backward_nofact(*h);
}
int pdplqr_backward_without_factorization(pdplqr_handle_t h, const double* rho) {
    if (!h) return PDPLQR_ERR_INVALID;
    cudaSetDevice(h->device);
    const double* dev = nullptr;
    int rc = stage_rho(h, rho, &dev);
    if (rc) return rc;
    return pdplqr_backward_without_factorization_device(h, dev);
}
int pdplqr_set_option(pdplqr_handle_t h, int option, int value) {
    if (!h) return PDPLQR_ERR_INVALID;
    if (h->solve_exec) { cudaGraphExecDestroy(h->solve_exec); h->solve_exec = nullptr; }   // options change the launch plan
    if (h->admm_exec) { cudaGraphExecDestroy(h->admm_exec); h->admm_exec = nullptr; }
    if (option == PDPLQR_OPT_AFFINE_CACHE) {
        if (value && !h->d_aff) {
            if (dev_alloc(*h, &h->d_aff, (size_t)h->batch * h->N * h->ops->AREC)) return PDPLQR_ERR_CUDA;
        }
        h->keep_affine = value != 0;
        h->factorized = false;
        return PDPLQR_OK;
    }
    if (option == PDPLQR_OPT_INTERIOR_SHARD) {
        if (!value && h->interior && h->model_set)
            return fail(h, PDPLQR_ERR_ORDER, "PDPLQR_OPT_INTERIOR_SHARD cannot be switched off after set_model (create a new handle)");
        h->interior = value != 0;
        h->have_root = h->root_fresh = false;
        if (h->interior) {
            if (h->thread_path) {   // the thread-per-problem records are packed differently: re-plan before set_model
                if (h->model_set) return fail(h, PDPLQR_ERR_ORDER, "set PDPLQR_OPT_INTERIOR_SHARD before set_model");
                h->thread_path = false;
                h->frec = h->ops->FREC;
                h->mrec = h->ops->REC;
                if (dev_alloc(*h, &h->d_model, (size_t)h->batch * h->N * h->ops->REC)) return PDPLQR_ERR_CUDA;
                if (dev_alloc(*h, &h->d_fac, (size_t)h->batch * h->N * h->ops->FREC)) return PDPLQR_ERR_CUDA;
            }
            if (!h->d_root_x) {
                if (dev_alloc(*h, &h->d_root_x, (size_t)h->batch * h->nx)) return PDPLQR_ERR_CUDA;
                if (dev_alloc(*h, &h->d_root_lam, (size_t)h->batch * h->nx)) return PDPLQR_ERR_CUDA;
            }
        }
        h->factorized = false;
        return PDPLQR_OK;
    }
    return fail(h, PDPLQR_ERR_INVALID, "unknown option");
}

// ---- horizon sharding (one handle per time slice / rank) --------------------------------------------------
int pdplqr_summary_doubles(pdplqr_handle_t h) { return h ? h->ops->SREC : PDPLQR_ERR_INVALID; }

int pdplqr_get_root_summary_device(pdplqr_handle_t h, double* out) {
    if (!h || !out) return fail(h, PDPLQR_ERR_INVALID, "get_root_summary: null pointer");
    if (!h->factorized) return fail(h, PDPLQR_ERR_ORDER, "get_root_summary before backward");
    cudaSetDevice(h->device);
    const double* src = (h->S > 1) ? h->levels.back().sum : h->d_sum;
    CU_TRY(h, cudaMemcpyAsync(out, src, (size_t)h->batch * h->ops->SREC * 8, cudaMemcpyDeviceToDevice, h->stream));
    return PDPLQR_OK;
}

int pdplqr_set_root_boundary_device(pdplqr_handle_t h, const double* xhat, const double* lam) {
    if (!h || !xhat) return fail(h, PDPLQR_ERR_INVALID, "set_root_boundary: null pointer");
    cudaSetDevice(h->device);
    if (!h->d_root_x) {
        if (dev_alloc(*h, &h->d_root_x, (size_t)h->batch * h->nx)) return PDPLQR_ERR_CUDA;
        if (dev_alloc(*h, &h->d_root_lam, (size_t)h->batch * h->nx)) return PDPLQR_ERR_CUDA;
    }
    const size_t nb = (size_t)h->batch * h->nx * 8;
    int rc = repack_vec(*h, xhat, h->d_root_x, h->batch, h->nxu, h->nx);   // (a plain copy when nothing is padded)
    if (rc) return rc;
    if (lam) rc = repack_vec(*h, lam, h->d_root_lam, h->batch, h->nxu, h->nx);
    else CU_TRY(h, cudaMemsetAsync(h->d_root_lam, 0, nb, h->stream));
    if (rc) return rc;
    h->root_fresh = true;
    return PDPLQR_OK;
}

// Coupler: solves the interface system of G time slices from their (all-gathered) root summaries.
int pdplqr_coupler_create(pdplqr_handle_t* out, int nx, int nu, int num_shards, int batch, int device) {
    if (!out) return PDPLQR_ERR_INVALID;
    *out = nullptr;
    if (num_shards < 1 || num_shards > TREE_TOP_MAX_NODES) return PDPLQR_ERR_INVALID;
    // a coupler is a handle whose "segments" are the shards: reuse the tree plan of an S = num_shards handle
    int rc = pdplqr_create(out, nx, nu, /*N=*/num_shards, nullptr, batch, /*num_segments=*/num_shards, 0,
                           PDPLQR_CONDENSED_LU, device);
    if (rc) return rc;
    (*out)->is_coupler = true;
    return PDPLQR_OK;
}

int pdplqr_coupler_solve_device(pdplqr_handle_t c, const double* summaries, const double* x0, double* xhat,
                                double* lam) {
    if (!c || !c->is_coupler || !summaries || !x0 || !xhat || !lam)
        return fail(c, PDPLQR_ERR_INVALID, "coupler_solve: bad arguments");
    cudaSetDevice(c->device);
    const int G = c->S;
    const long long items = (long long)c->batch * G;
    if (G == 1) {
        CU_TRY(c, cudaMemcpyAsync(xhat, x0, (size_t)c->batch * c->nxu * 8, cudaMemcpyDeviceToDevice, c->stream));
        CU_TRY(c, cudaMemsetAsync(lam, 0, (size_t)items * c->nxu * 8, c->stream));
        return PDPLQR_OK;
    }
    // summaries are laid out [batch][G][SREC] exactly like the level-0 array of the tree (kernel dimensions: for a padded
    // (nx, nu) they are opaque records, produced by pdplqr_get_root_summary_device of handles with the same (nx, nu))
    CU_TRY(c, cudaMemcpyAsync(c->d_sum, summaries, (size_t)items * c->ops->SREC * 8, cudaMemcpyDeviceToDevice, c->stream));
    int rc = PDPLQR_OK;
    if (c->padded) {
        rc = repack_vec(*c, x0, c->d_x0p, c->batch, c->nxu, c->nx);
        if (rc) return rc;
        x0 = c->d_x0p;
    }
    c->chain_tail = false;   // (the D2D copy above precedes the first tree kernel: ordinary launch, see launch_chain)
    rc = run_tree_up(*c, false);
    if (rc) return rc;
    rc = run_tree_down(*c, x0, nullptr);
    if (rc) return rc;
    rc = repack_vec(*c, c->d_xhat, xhat, items, c->nx, c->nxu);
    if (rc) return rc;
    return repack_vec(*c, c->d_uhat, lam, items, c->nx, c->nxu);
}

int pdplqr_forward_device(pdplqr_handle_t h, const double* x0, double* ws_out) {
    if (!h || !x0 || !ws_out) return fail(h, PDPLQR_ERR_INVALID, "forward: null pointer");
    cudaSetDevice(h->device);
    return run_forward_user(*h, x0, ws_out);
}
int pdplqr_forward(pdplqr_handle_t h, const double* x0, double* ws_out) {
    if (!h || !x0 || !ws_out) return fail(h, PDPLQR_ERR_INVALID, "forward: null pointer");
    cudaSetDevice(h->device);
    const size_t ws_len = (size_t)h->N * h->su + h->nxu;   // caller's layout
    CU_TRY(h, cudaMemcpyAsync(h->d_x0, x0, (size_t)h->batch * h->nxu * 8, cudaMemcpyHostToDevice, h->stream));
    int rc = run_forward_user(*h, h->d_x0, h->d_ws_out);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(ws_out, h->d_ws_out, ws_len * h->batch * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return PDPLQR_OK;
}
// Pipelined host solve for large batches on the thread-per-problem path: the batch is cut into chunks; chunk c's
// H2D copy (stream s_in), kernels (handle stream) and D2H copy (stream s_out) overlap with the neighbours', so the
// end-to-end time approaches max(H2D, D2H) instead of H2D + kernels + D2H.  Host buffers should be pinned.
static int solve_pipelined(Solver& h, const double* ws_in, double sigma, const double* x0, double* ws_out) {
    const int C = h.pipeline_chunks;
    if (!h.s_in) {
        CU_TRY(&h, cudaStreamCreateWithFlags(&h.s_in, cudaStreamNonBlocking));
        CU_TRY(&h, cudaStreamCreateWithFlags(&h.s_out, cudaStreamNonBlocking));
        h.ev_in.resize(C); h.ev_cmp.resize(C);
        for (int c = 0; c < C; ++c) {
            CU_TRY(&h, cudaEventCreateWithFlags(&h.ev_in[c], cudaEventDisableTiming));
            CU_TRY(&h, cudaEventCreateWithFlags(&h.ev_cmp[c], cudaEventDisableTiming));
        }
        CU_TRY(&h, cudaEventCreateWithFlags(&h.ev_ready, cudaEventDisableTiming));
    }
    const size_t wsl = (size_t)h.N * h.s + h.nx;
    const int per = (((h.batch + C - 1) / C) + 127) / 128 * 128;   // whole CTAs of the default 4-warp launch
    // copies must not start before earlier work on the handle's stream (e.g. set_model) has finished
    CU_TRY(&h, cudaEventRecord(h.ev_ready, h.stream));
    CU_TRY(&h, cudaStreamWaitEvent(h.s_in, h.ev_ready, 0));
    CU_TRY(&h, cudaStreamWaitEvent(h.s_out, h.ev_ready, 0));
    h.cur_ws = ws_in ? h.d_ws_in : nullptr;
    h.sigma = sigma;
    int rc = PDPLQR_OK;
    for (int c = 0, b0 = 0; b0 < h.batch && rc == PDPLQR_OK; ++c, b0 += per) {
        const int nb = std::min(per, h.batch - b0);
        if (ws_in) CU_TRY(&h, cudaMemcpyAsync(h.d_ws_in + b0 * wsl, ws_in + b0 * wsl, nb * wsl * 8, cudaMemcpyHostToDevice, h.s_in));
        CU_TRY(&h, cudaMemcpyAsync(h.d_x0 + (size_t)b0 * h.nx, x0 + (size_t)b0 * h.nx, (size_t)nb * h.nx * 8, cudaMemcpyHostToDevice, h.s_in));
        CU_TRY(&h, cudaEventRecord(h.ev_in[c], h.s_in));
        CU_TRY(&h, cudaStreamWaitEvent(h.stream, h.ev_in[c], 0));
        h.chunk_b0 = b0; h.chunk_nb = nb;
        rc = h.ops->backward(h);
        if (rc == PDPLQR_OK) rc = h.ops->forward(h, h.d_x0, h.d_ws_out);
        h.chunk_nb = 0; h.chunk_b0 = 0;
        if (rc) break;
        CU_TRY(&h, cudaEventRecord(h.ev_cmp[c], h.stream));
        CU_TRY(&h, cudaStreamWaitEvent(h.s_out, h.ev_cmp[c], 0));
        CU_TRY(&h, cudaMemcpyAsync(ws_out + b0 * wsl, h.d_ws_out + b0 * wsl, nb * wsl * 8, cudaMemcpyDeviceToHost, h.s_out));
    }
    if (rc) return rc;
    CU_TRY(&h, cudaStreamSynchronize(h.s_out));
    CU_TRY(&h, cudaStreamSynchronize(h.stream));
    h.updated = false; h.factorized = true; h.backward_done = false;
    return PDPLQR_OK;
}

int pdplqr_solve(pdplqr_handle_t h, const double* ws_in, const double* ys, const double* zs, const double* rho,
                 const double* inv_rho, double sigma, const double* x0, double* ws_out) {
    if (h && x0 && ws_out && h->thread_path && h->model_set && h->pipeline_chunks > 1 && h->batch >= 4096 && !h->root_fresh && !h->padded) {
        cudaSetDevice(h->device);
        return solve_pipelined(*h, ws_in, sigma, x0, ws_out);
    }
    int rc = pdplqr_update_problem_data(h, ws_in, ys, zs, inv_rho, sigma);
    if (rc) return rc;
    rc = pdplqr_backward(h, rho);
    if (rc) return rc;
    return pdplqr_forward(h, x0, ws_out);
}
// update_problem_data + backward + forward on device arrays, enqueued as ONE CUDA graph launch (captured on first use and
// whenever a pointer, sigma or the stream changes).  For a single latency-bound problem the five to eight dependent
// kernels of a solve then start back to back without a host round trip per launch.
int pdplqr_solve_device(pdplqr_handle_t h, const double* ws_in, const double* ys, const double* zs, const double* rho,
                        const double* inv_rho, double sigma, const double* x0, double* ws_out) {
    if (!h || !x0 || !ws_out) return fail(h, PDPLQR_ERR_INVALID, "solve_device: null pointer");
    if (!h->model_set) return fail(h, PDPLQR_ERR_ORDER, "solve_device before set_model");
    cudaSetDevice(h->device);
    auto plain = [&]() {
        // one chain from the stage sweep to the rollout: nothing the caller does can come between backward and forward here
        struct Fused { Solver* s; Fused(Solver* q) : s(q) { q->fused = 1; q->chain_tail = false; } ~Fused() { s->fused = 0; s->chain_tail = false; } } fused(h);
        int rc = pdplqr_update_problem_data_device(h, ws_in, ys, zs, inv_rho, sigma);
        if (rc) return rc;
        rc = pdplqr_backward_device(h, rho);
        if (rc) return rc;
        return pdplqr_forward_device(h, x0, ws_out);
    };
    if (!h->solve_use_graph || h->interior || h->root_fresh) return plain();   // (shard boundaries change per solve)
    const auto& k = h->solve_key;
    const bool same = h->solve_exec && k.ws == ws_in && k.ys == ys && k.zs == zs && k.rho == rho && k.inv_rho == inv_rho &&
                      k.x0 == x0 && k.out == ws_out && k.sigma == sigma && k.stream == h->stream;
    if (!same) {
        if (h->solve_exec) { cudaGraphExecDestroy(h->solve_exec); h->solve_exec = nullptr; }
        const long long l0 = h->launches;
        const bool was_factorized = h->factorized;
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed);
        if (e != cudaSuccess) { cudaGetLastError(); return plain(); }
        int rc = plain();
        e = cudaStreamEndCapture(h->stream, &g);
        h->solve_kernels = (int)(h->launches - l0);
        h->launches = l0;
        h->factorized = was_factorized; h->backward_done = false; h->updated = false;   // nothing ran during the capture
        if (rc != PDPLQR_OK || e != cudaSuccess || !g) {
            cudaGetLastError();
            if (g) cudaGraphDestroy(g);
            return rc != PDPLQR_OK ? rc : plain();
        }
        e = cudaGraphInstantiate(&h->solve_exec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { cudaGetLastError(); h->solve_exec = nullptr; return plain(); }
        h->solve_key = {ws_in, ys, zs, rho, inv_rho, x0, ws_out, sigma, h->stream};
    }
    CU_TRY(h, cudaGraphLaunch(h->solve_exec, h->stream));
    h->launches += h->solve_kernels;
    // host-side protocol state as after update_problem_data + backward + forward
    h->cur_ws = (h->padded && ws_in) ? h->d_wsp_in : ws_in;
    h->cur_ys = ys; h->cur_zs = zs; h->cur_inv_rho = inv_rho; h->cur_rho = rho; h->sigma = sigma;
    h->updated = false; h->factorized = true; h->backward_done = false; h->have_root = false;
    return PDPLQR_OK;
}
int pdplqr_synchronize(pdplqr_handle_t h) {
    if (!h) return PDPLQR_ERR_INVALID;
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return PDPLQR_OK;
}

int pdplqr_num_segments(pdplqr_handle_t h) { return h ? h->S : PDPLQR_ERR_INVALID; }
int pdplqr_get_partition(pdplqr_handle_t h, int* starts, int* lens) {
    if (!h) return PDPLQR_ERR_INVALID;
    for (int i = 0; i < h->S; ++i) {
        if (starts) starts[i] = h->seg_start[i];
        if (lens) lens[i] = h->seg_len[i];
    }
    return PDPLQR_OK;
}
int pdplqr_get_gains(pdplqr_handle_t h, double* K, double* d, double* Gt) {
    if (!h) return PDPLQR_ERR_INVALID;
    cudaSetDevice(h->device);
    // the factor records are unpacked on the device into the caller's layout, a bounded chunk of problems at a time
    // (this accessor used to copy the whole factor array to the host)
    const int nxu = h->nxu, nuu = h->nuu;
    const size_t per_prob = (size_t)h->N * (2 * (size_t)nuu * nxu + nuu);
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>(h->batch, (64u << 20) / (per_prob * 8 + 1) + 1));
    double* scratch = nullptr;
    CU_TRY(h, cudaMalloc(&scratch, chunk * per_prob * 8));
    int rc = PDPLQR_OK;
    for (int b0 = 0; b0 < h->batch && rc == PDPLQR_OK; b0 += chunk) {
        const int nb = std::min(chunk, h->batch - b0);
        const size_t nst = (size_t)nb * h->N;
        double *dK = scratch, *dd = dK + nst * nuu * nxu, *dG = dd + nst * nuu;
        const long long total = (long long)nst * (2 * nuu * nxu + nuu);
        unpack_gains_kernel<<<ew_blocks(total), 256, 0, h->stream>>>(h->d_fac, dK, dd, dG, b0, nb, h->N, h->nx, h->nu, nxu, nuu,
                                                                     h->frec, h->thread_path ? 1 : 0,
                                                                     h->interior ? h->N : h->seg_start[h->S - 1]);
        h->launches++;
        cudaError_t e = cudaGetLastError();
        const size_t o = (size_t)b0 * h->N;
        if (e == cudaSuccess && K) e = cudaMemcpyAsync(K + o * nuu * nxu, dK, nst * nuu * nxu * 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess && d) e = cudaMemcpyAsync(d + o * nuu, dd, nst * nuu * 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess && Gt) e = cudaMemcpyAsync(Gt + o * nuu * nxu, dG, nst * nuu * nxu * 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(h, PDPLQR_ERR_CUDA, std::string("get_gains: ") + cudaGetErrorString(e));
    }
    cudaFree(scratch);
    return rc;
}
int pdplqr_get_interface(pdplqr_handle_t h, double* xhat, double* uhat) {
    if (!h) return PDPLQR_ERR_INVALID;
    cudaSetDevice(h->device);
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    const size_t cnt = (size_t)h->batch * h->S;
    std::vector<double> host(cnt * h->nx);
    for (int which = 0; which < 2; ++which) {
        double* dst = which ? uhat : xhat;
        if (!dst) continue;
        CU_TRY(h, cudaMemcpy(host.data(), which ? h->d_uhat : h->d_xhat, host.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < cnt; ++i) std::memcpy(dst + i * h->nxu, host.data() + i * h->nx, 8 * h->nxu);
    }
    return PDPLQR_OK;
}
int pdplqr_get_summaries(pdplqr_handle_t h, double* P, double* p, double* F, double* f, double* C) {
    if (!h) return PDPLQR_ERR_INVALID;
    cudaSetDevice(h->device);
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    const int nx = h->nx, n2 = nx * nx, SREC = h->ops->SREC, nxu = h->nxu;
    const size_t cnt = (size_t)h->batch * h->S;
    std::vector<double> host(cnt * SREC);
    CU_TRY(h, cudaMemcpy(host.data(), h->d_sum, host.size() * 8, cudaMemcpyDeviceToHost));
    auto block = [&](double* dst, const double* src) {   // leading nxu x nxu block of an nx x nx column-major matrix
        for (int j = 0; j < nxu; ++j) std::memcpy(dst + (size_t)j * nxu, src + (size_t)j * nx, 8 * nxu);
    };
    for (size_t i = 0; i < cnt; ++i) {
        const double* r = host.data() + i * SREC;
        if (P) block(P + i * nxu * nxu, r);
        if (F) block(F + i * nxu * nxu, r + n2);
        if (C) block(C + i * nxu * nxu, r + 2 * n2);
        if (p) std::memcpy(p + i * nxu, r + 3 * n2, 8 * nxu);
        if (f) std::memcpy(f + i * nxu, r + 3 * n2 + nx, 8 * nxu);
    }
    return PDPLQR_OK;
}
// costates with kernel-layout device arrays (traj [batch][N s + nx], lam [batch][N][nx])
static int costates_raw(pdplqr_handle_t h, const double* traj, double* lam) {
    if (h->is_coupler) return fail(h, PDPLQR_ERR_INVALID, "get_costates: a coupler handle has no stages");
    if (!h->factorized || h->backward_done)
        return fail(h, PDPLQR_ERR_ORDER, "get_costates needs a completed backward + forward (interface costates)");
    if (h->interior && !h->have_root)
        return fail(h, PDPLQR_ERR_ORDER, "get_costates on an interior horizon shard needs its root boundary");
    return h->ops->costates(*h, traj, lam);
}
static int costate_scratch(pdplqr_handle_t h) {   // allocated once, on first use
    if (h->d_cost_ws) return PDPLQR_OK;
    const size_t wsl = (size_t)h->N * h->s + h->nx, B = h->batch;
    if (dev_alloc(*h, &h->d_cost_ws, B * wsl)) return PDPLQR_ERR_CUDA;
    if (dev_alloc(*h, &h->d_cost_lam, B * h->N * h->nx)) return PDPLQR_ERR_CUDA;
    return PDPLQR_OK;
}
int pdplqr_get_costates_device(pdplqr_handle_t h, const double* ws, double* lam) {
    if (!h || !ws || !lam) return fail(h, PDPLQR_ERR_INVALID, "get_costates: bad arguments");
    cudaSetDevice(h->device);
    if (!h->padded) return costates_raw(h, ws, lam);
    int rc = costate_scratch(h);
    if (rc) return rc;
    rc = repack_ws(*h, ws, h->d_cost_ws, true);
    if (rc) return rc;
    rc = costates_raw(h, h->d_cost_ws, h->d_cost_lam);
    if (rc) return rc;
    return repack_vec(*h, h->d_cost_lam, lam, (long long)h->batch * h->N, h->nx, h->nxu);
}
int pdplqr_get_costates(pdplqr_handle_t h, const double* ws, double* lam) {
    if (!h || !ws || !lam) return fail(h, PDPLQR_ERR_INVALID, "get_costates: bad arguments");
    cudaSetDevice(h->device);
    const size_t wsl_user = (size_t)h->N * h->su + h->nxu, B = h->batch;
    int rc = costate_scratch(h);
    if (rc) return rc;
    // d_ws_out holds nothing the costate kernel reads (w_prev lives in d_ws_in / d_wsp_in): use it for the upload
    CU_TRY(h, cudaMemcpyAsync(h->d_ws_out, ws, B * wsl_user * 8, cudaMemcpyHostToDevice, h->stream));
    const double* traj = h->d_ws_out;
    if (h->padded) {
        rc = repack_ws(*h, h->d_ws_out, h->d_cost_ws, true);
        if (rc) return rc;
        traj = h->d_cost_ws;
    }
    rc = costates_raw(h, traj, h->d_cost_lam);
    if (rc) return rc;
    const double* src = h->d_cost_lam;
    if (h->padded) {   // crop into the (now free) trajectory scratch
        rc = repack_vec(*h, h->d_cost_lam, h->d_cost_ws, (long long)B * h->N, h->nx, h->nxu);
        if (rc) return rc;
        src = h->d_cost_ws;
    }
    CU_TRY(h, cudaMemcpyAsync(lam, src, B * h->N * h->nxu * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return PDPLQR_OK;
}
int pdplqr_last_status(pdplqr_handle_t h, int* status) {
    if (!h) return PDPLQR_ERR_INVALID;
    cudaSetDevice(h->device);
    const size_t S = h->S;
    std::vector<int> host((size_t)h->batch * S);   // one slot per (problem, segment)
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    CU_TRY(h, cudaMemcpy(host.data(), h->d_status, sizeof(int) * host.size(), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int b = 0; b < h->batch; ++b) {
        int worst = 0;   // the latest stage with a non-positive pivot (the sweep runs backwards: the first one met)
        for (size_t sg = 0; sg < S; ++sg) worst = std::max(worst, host[(size_t)b * S + sg]);
        bad += worst != 0;
        if (status) status[b] = worst;
    }
    return bad;
}
// h == NULL: the reason the last pdplqr_create / pdplqr_coupler_create on this thread failed
const char* pdplqr_last_error(pdplqr_handle_t h) { return h ? h->err.c_str() : g_create_error.c_str(); }
long long pdplqr_launch_count(pdplqr_handle_t h) { return h ? h->launches : 0; }
int pdplqr_debug_check_guards(pdplqr_handle_t h, long long* corrupted_bytes) {
    if (!h || !corrupted_bytes) return PDPLQR_ERR_INVALID;
    const bool self_test = (*corrupted_bytes == PDPLQR_GUARD_SELF_TEST);
    *corrupted_bytes = -1;
    if (!h->guards) return PDPLQR_OK;
    if (self_test && !h->guarded.empty())   // prove the detector: 3 bytes written just past the first allocation
        CU_TRY(h, cudaMemset(h->guarded[0].base + Solver::GUARD_BYTES + h->guarded[0].bytes, 0, 3));
    cudaSetDevice(h->device);
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    CU_TRY(h, cudaDeviceSynchronize());   // the copy streams of the pipelined host solve as well
    constexpr size_t G = Solver::GUARD_BYTES;
    std::vector<unsigned char> host(2 * G);
    long long bad = 0;
    for (const Solver::GuardedAlloc& a : h->guarded) {
        CU_TRY(h, cudaMemcpy(host.data(), a.base, G, cudaMemcpyDeviceToHost));
        CU_TRY(h, cudaMemcpy(host.data() + G, a.base + G + a.bytes, G, cudaMemcpyDeviceToHost));
        for (unsigned char v : host) bad += (v != 0xA5);
    }
    *corrupted_bytes = bad;
    return PDPLQR_OK;
}
int pdplqr_record_doubles(pdplqr_handle_t h, int* model_rec, int* factor_rec) {
    if (!h) return PDPLQR_ERR_INVALID;
    if (model_rec) *model_rec = h->mrec;
    if (factor_rec) *factor_rec = h->frec;
    return PDPLQR_OK;
}

// ---- conic ADMM outer iteration (addition; NOT in the reference -- SURVEY.md section 8 rows a11 / f1) -----------------
int pdplqr_admm_set_cones(pdplqr_handle_t h, int ncones, const int* stage, const int* row0, const int* dim,
                          const int* type, const double* e_lb, const double* e_ub) {
    if (!h || h->nc_total == 0) return fail(h, PDPLQR_ERR_INVALID, "admm_set_cones: the problem has no constraints");
    if (ncones < 1 || !stage || !row0 || !dim || !type || !e_lb || !e_ub) return fail(h, PDPLQR_ERR_INVALID, "admm_set_cones: bad arguments");
    cudaSetDevice(h->device);
    // cones must be given stage by stage (non-decreasing), tile the rows of every stage exactly once
    std::vector<int> first(h->N + 2, 0);
    std::vector<int> covered(h->N + 1, 0);
    for (int c = 0; c < ncones; ++c) {
        if (stage[c] < 0 || stage[c] > h->N || (c > 0 && stage[c] < stage[c - 1])) return fail(h, PDPLQR_ERR_INVALID, "admm_set_cones: cones must be sorted by stage");
        if (row0[c] != covered[stage[c]] || dim[c] < 1 || type[c] < 0 || type[c] > 2) return fail(h, PDPLQR_ERR_INVALID, "admm_set_cones: cones must tile the rows of a stage in order");
        covered[stage[c]] += dim[c];
        first[stage[c] + 1] = c + 1;
    }
    for (int k = 0; k <= h->N; ++k) {
        if (covered[k] != h->ncs[k]) return fail(h, PDPLQR_ERR_INVALID, "admm_set_cones: rows of a stage not covered");
        if (first[k + 1] < first[k]) first[k + 1] = first[k];
    }
    const size_t B = h->batch, nct = (size_t)h->nc_total, wsl = (size_t)h->N * h->s + h->nx;   // (kernel layout: >= caller's)
    int rc = 0;
    if (!h->d_cone_first) {
        rc |= dev_alloc(*h, &h->d_cone_first, h->N + 2);
        rc |= dev_alloc(*h, &h->d_elb, B * nct);
        rc |= dev_alloc(*h, &h->d_eub, B * nct);
        rc |= dev_alloc(*h, &h->d_wtilde, B * wsl);
        rc |= dev_alloc(*h, &h->d_w, B * wsl);
        rc |= dev_alloc(*h, &h->d_z, B * nct);
        rc |= dev_alloc(*h, &h->d_y, B * nct);
        rc |= dev_alloc(*h, &h->d_rho_admm, B * nct);
        rc |= dev_alloc(*h, &h->d_invrho_admm, B * nct);
        rc |= dev_alloc(*h, &h->d_rho_work, B * nct);
        rc |= dev_alloc(*h, &h->d_invrho_work, B * nct);
        if (h->padded) rc |= dev_alloc(*h, &h->d_wk, B * wsl);
        AdmmCtl* ctl = nullptr;
        rc |= dev_alloc(*h, &ctl, 1);
        h->d_ctl = ctl;
    }
    rc |= dev_alloc(*h, &h->d_row_box, nct);
    rc |= dev_alloc(*h, &h->d_cone_type, ncones);
    rc |= dev_alloc(*h, &h->d_cone_row, ncones);
    rc |= dev_alloc(*h, &h->d_cone_dim, ncones);
    if (rc) return PDPLQR_ERR_CUDA;
    h->ncones = ncones;
    CU_TRY(h, cudaMemcpy(h->d_cone_first, first.data(), sizeof(int) * (h->N + 2), cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(h->d_cone_type, type, sizeof(int) * ncones, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(h->d_cone_row, row0, sizeof(int) * ncones, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(h->d_cone_dim, dim, sizeof(int) * ncones, cudaMemcpyHostToDevice));
    {   // per-row flag "belongs to a box cone" (admm_update_kernel finishes those rows in one pass)
        std::vector<long long> off(h->N + 2, 0);
        for (int k = 0; k <= h->N; ++k) off[k + 1] = off[k] + h->ncs[k];
        std::vector<int> box(nct, 0);
        for (int c = 0; c < ncones; ++c)
            if (type[c] == PDPLQR_CONE_BOX)
                for (int r = 0; r < dim[c]; ++r) box[off[stage[c]] + row0[c] + r] = 1;
        CU_TRY(h, cudaMemcpy(h->d_row_box, box.data(), sizeof(int) * nct, cudaMemcpyHostToDevice));
    }
    CU_TRY(h, cudaMemcpy(h->d_elb, e_lb, B * nct * 8, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(h->d_eub, e_ub, B * nct * 8, cudaMemcpyHostToDevice));
    h->cones_set = true;
    if (h->admm_exec) { cudaGraphExecDestroy(h->admm_exec); h->admm_exec = nullptr; }   // the cone tables are baked in
    if (h->admm_graph) { cudaGraphDestroy(h->admm_graph); h->admm_graph = nullptr; }
    return PDPLQR_OK;
}

int pdplqr_admm_configure(pdplqr_handle_t h, int use_graph, int adaptive_rho, double rho_tau, int max_rho_updates) {
    if (!h || rho_tau <= 1.0 || max_rho_updates < 0) return fail(h, PDPLQR_ERR_INVALID, "admm_configure: bad arguments");
    h->admm_use_graph = use_graph ? 1 : 0;
    h->admm_adaptive = adaptive_rho ? 1 : 0;
    h->admm_rho_tau = rho_tau;
    h->admm_max_rho_updates = max_rho_updates;
    return PDPLQR_OK;
}
int pdplqr_admm_stats(pdplqr_handle_t h, int* graph_launches, int* rho_updates) {
    if (!h) return PDPLQR_ERR_INVALID;
    if (graph_launches) *graph_launches = h->admm_graph_launches;
    if (rho_updates) *rho_updates = h->admm_rho_updates_last;
    return PDPLQR_OK;
}

namespace {

struct AdmmRun {   // kernel-layout device pointers of one conic solve
    const double* x0;
    double *w, *z, *y;
    double sigma, alpha;
};

// one outer iteration, enqueued on the handle's stream (captured into the graph, or issued directly by the fallback)
int admm_iteration(Solver& h, const AdmmRun& r, bool factorize, cudaGraphConditionalHandle handle, int use_handle) {
    h.cur_ws = r.w; h.cur_ys = r.y; h.cur_zs = r.z; h.cur_inv_rho = h.d_invrho_work; h.cur_rho = h.d_rho_work;
    h.sigma = r.sigma;
    h.updated = true;                                              // update_problem_data   lqr_solver_parallel.hpp:115-140
    int rc = factorize ? run_backward(h) : run_backward_nofact(h);  // backward[_without_factorization]   :142-154
    if (rc) return rc;
    rc = run_forward(h, r.x0, h.d_wtilde);                         // forward   :213-238
    if (rc) return rc;
    AdmmParams ap{};
    ap.nx = h.nx; ap.nu = h.nu; ap.N = h.N; ap.batch = h.batch; ap.ncmax = h.ncmax;
    ap.ncs = h.d_ncs; ap.coff = h.d_coff; ap.doff = h.d_doff; ap.Dm = h.d_D;
    ap.d_total = h.d_total_dev; ap.nc_total = h.nc_total;
    ap.sel_col = h.sel_mode ? h.d_sel_col : nullptr; ap.sel_val = h.sel_mode ? h.d_sel_val : nullptr;
    ap.cone_first = h.d_cone_first; ap.cone_type = h.d_cone_type; ap.cone_row = h.d_cone_row; ap.cone_dim = h.d_cone_dim; ap.row_box = h.d_row_box;
    ap.e_lb = h.d_elb; ap.e_ub = h.d_eub;
    ap.w_tilde = h.d_wtilde; ap.w = r.w; ap.z = r.z; ap.y = r.y; ap.rho = h.d_rho_work;
    ap.alpha = r.alpha; ap.sigma = r.sigma; ap.ctl = static_cast<AdmmCtl*>(h.d_ctl);
    const size_t smem = (size_t)(h.s + 3 * h.ncmax) * sizeof(double) * ADMM_WARPS;            // ordinary iterations
    const size_t smem_res = (size_t)(2 * h.s + 3 * h.ncmax) * sizeof(double) * ADMM_WARPS;    // check iterations
    // persistent warps: a few CTAs per SM loop over the (problem, stage) items
    const long long items = (long long)h.batch * (h.N + 1);
    const int ctas = (int)std::min<long long>((items + ADMM_WARPS - 1) / ADMM_WARPS, 148LL * 8);
    rc = set_smem(h, admm_update_kernel, smem);
    if (rc) return rc;
    rc = set_smem(h, admm_update_res_kernel, smem_res);
    if (rc) return rc;
    // both kernels are enqueued; the one whose turn it is not returns at once (decided on the device)
    admm_relax_kernel<<<148 * 8, 256, 0, h.stream>>>(ap);
    admm_update_kernel<<<ctas, ADMM_WARPS * 32, smem, h.stream>>>(ap);
    admm_update_res_kernel<<<ctas, ADMM_WARPS * 32, smem_res, h.stream>>>(ap);
    h.launches += 3;
    CU_TRY(&h, cudaGetLastError());
    admm_ctl_kernel<<<1, 1, 0, h.stream>>>(static_cast<AdmmCtl*>(h.d_ctl), handle, use_handle);
    h.launches++;
    CU_TRY(&h, cudaGetLastError());
    return PDPLQR_OK;
}

// Builds the graph of one conic solve:  [factorising iteration] -> WHILE (condition set by admm_ctl_kernel) { affine-only
// iteration }.  Returns PDPLQR_OK with h.admm_exec set, or an error (the caller falls back to the host loop).
int admm_build_graph(Solver& h, const AdmmRun& r) {
    if (h.admm_exec) { cudaGraphExecDestroy(h.admm_exec); h.admm_exec = nullptr; }
    if (h.admm_graph) { cudaGraphDestroy(h.admm_graph); h.admm_graph = nullptr; }
    const bool was_factorized = h.factorized;
    const long long l0 = h.launches;   // the capture counts kernels without running them: per-iteration counts are kept
    cudaGraph_t g = nullptr;
    CU_TRY(&h, cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle handle;
    cudaError_t e = cudaGraphConditionalHandleCreate(&handle, g, 0, cudaGraphCondAssignDefault);
    if (e != cudaSuccess) { cudaGraphDestroy(g); return fail(&h, PDPLQR_ERR_CUDA, std::string("conditional handle: ") + cudaGetErrorString(e)); }
    int rc = PDPLQR_OK;
    e = cudaStreamBeginCaptureToGraph(h.stream, g, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
    if (e != cudaSuccess) { cudaGraphDestroy(g); return fail(&h, PDPLQR_ERR_CUDA, std::string("begin capture: ") + cudaGetErrorString(e)); }
    rc = admm_iteration(h, r, true, handle, 1);
    h.admm_fact_kernels = (int)(h.launches - l0);
    cudaGraph_t body = nullptr;
    if (rc == PDPLQR_OK) {   // WHILE node after everything captured so far; later captured work would depend on it
        cudaStreamCaptureStatus st;
        unsigned long long id;
        cudaGraph_t cg;
        const cudaGraphNode_t* deps = nullptr;
        size_t nd = 0;
        e = cudaStreamGetCaptureInfo_v2(h.stream, &st, &id, &cg, &deps, &nd);
        cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
        cp.conditional.handle = handle;
        cp.conditional.type = cudaGraphCondTypeWhile;
        cp.conditional.size = 1;
        cudaGraphNode_t node;
        if (e == cudaSuccess) e = cudaGraphAddNode(&node, g, deps, nd, &cp);
        if (e == cudaSuccess) e = cudaStreamUpdateCaptureDependencies(h.stream, &node, 1, cudaStreamSetCaptureDependencies);
        if (e == cudaSuccess) body = cp.conditional.phGraph_out[0];
        else rc = fail(&h, PDPLQR_ERR_CUDA, std::string("conditional node: ") + cudaGetErrorString(e));
    }
    cudaGraph_t ended = nullptr;
    e = cudaStreamEndCapture(h.stream, &ended);
    if (rc == PDPLQR_OK && e != cudaSuccess) rc = fail(&h, PDPLQR_ERR_CUDA, std::string("end capture: ") + cudaGetErrorString(e));
    if (rc == PDPLQR_OK) {
        e = cudaStreamBeginCaptureToGraph(h.stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
        if (e != cudaSuccess) rc = fail(&h, PDPLQR_ERR_CUDA, std::string("begin body capture: ") + cudaGetErrorString(e));
        else {
            const long long l1 = h.launches;
            rc = admm_iteration(h, r, false, handle, 1);
            h.admm_aff_kernels = (int)(h.launches - l1);
            e = cudaStreamEndCapture(h.stream, &ended);
            if (rc == PDPLQR_OK && e != cudaSuccess) rc = fail(&h, PDPLQR_ERR_CUDA, std::string("end body capture: ") + cudaGetErrorString(e));
        }
    }
    if (rc == PDPLQR_OK) {
        e = cudaGraphInstantiate(&h.admm_exec, g, 0);
        if (e != cudaSuccess) rc = fail(&h, PDPLQR_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
    }
    // nothing ran during the capture: restore the host-side protocol state
    h.launches = l0;
    h.factorized = was_factorized; h.backward_done = false; h.updated = false;
    if (rc != PDPLQR_OK) {
        cudaGetLastError();
        cudaGraphDestroy(g);
        h.admm_exec = nullptr;
        return rc;
    }
    h.admm_graph = g;
    h.admm_key = {r.x0, r.w, r.z, r.y, r.sigma, r.alpha, h.stream};
    return PDPLQR_OK;
}

}  // namespace

// device-resident loop: w, z, y (in/out), rho, inv_rho, x0 are device arrays.  One CUDA graph launch per conic solve
// (plus one per rho rescale when the adaptation is on); the only host reads are the 120-byte control block at the end.
int pdplqr_admm_solve_device(pdplqr_handle_t h, const double* x0, double* w, double* z, double* y, const double* rho,
                             const double* inv_rho, double sigma, double alpha, int max_iter, double eps_abs,
                             double eps_rel, int check_every, int* iters_out, double* residuals_out) {
    if (!h || !x0 || !w || !z || !y || !rho || !inv_rho || max_iter < 1)
        return fail(h, PDPLQR_ERR_INVALID, "admm_solve: bad arguments");
    if (!h->cones_set) return fail(h, PDPLQR_ERR_ORDER, "admm_solve before admm_set_cones");
    if (!h->model_set) return fail(h, PDPLQR_ERR_ORDER, "admm_solve before set_model");
    cudaSetDevice(h->device);
    if (check_every < 1) check_every = 1;
    // the iterations live inside a WHILE conditional node: plain kernel dependencies there (no programmatic edges)
    struct PdlOff { Solver* s; int saved; PdlOff(Solver* q) : s(q), saved(q->use_pdl) { q->use_pdl = 0; } ~PdlOff() { s->use_pdl = saved; } } pdl_off(h);
    const size_t nb = (size_t)h->batch * h->nc_total * 8;
    // rho and 1/rho are copied: the adaptation rescales the library's copies, the caller's arrays stay as given
    CU_TRY(h, cudaMemcpyAsync(h->d_rho_work, rho, nb, cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(h->d_invrho_work, inv_rho, nb, cudaMemcpyDeviceToDevice, h->stream));
    AdmmRun r{x0, w, z, y, sigma, alpha};
    int rc = PDPLQR_OK;
    if (h->padded) {   // iterate and initial state in kernel layout for the whole loop
        rc = repack_ws(*h, w, h->d_wk, true);
        if (rc) return rc;
        rc = repack_vec(*h, x0, h->d_x0p, h->batch, h->nxu, h->nx);
        if (rc) return rc;
        r.w = h->d_wk; r.x0 = h->d_x0p;
    }
    AdmmCtl ctl{};
    ctl.max_iter = max_iter; ctl.check_every = check_every;
    ctl.adaptive = h->admm_adaptive; ctl.max_rho_updates = h->admm_max_rho_updates;
    ctl.eps_abs = eps_abs; ctl.eps_rel = eps_rel; ctl.rho_tau = h->admm_rho_tau; ctl.rho_scale = 1.0;
    AdmmCtl* d_ctl = static_cast<AdmmCtl*>(h->d_ctl);
    CU_TRY(h, cudaMemcpyAsync(d_ctl, &ctl, sizeof(ctl), cudaMemcpyHostToDevice, h->stream));
    bool use_graph = h->admm_use_graph != 0;
    if (use_graph) {
        const auto& k = h->admm_key;
        const bool same = h->admm_exec && k.x0 == r.x0 && k.w == r.w && k.z == r.z && k.y == r.y && k.sigma == sigma &&
                          k.alpha == alpha && k.stream == h->stream;
        if (!same && admm_build_graph(*h, r) != PDPLQR_OK) use_graph = false;   // (reason kept in last_error)
    }
    h->admm_rho_updates_last = 0;
    for (;;) {
        if (use_graph) {
            const int it0 = ctl.iter;
            CU_TRY(h, cudaGraphLaunch(h->admm_exec, h->stream));
            h->admm_graph_launches++;
            CU_TRY(h, cudaMemcpyAsync(&ctl, d_ctl, sizeof(ctl), cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(h, cudaStreamSynchronize(h->stream));
            // kernels that ran inside this launch: one factorising iteration, then affine-only ones
            h->launches += h->admm_fact_kernels + (long long)std::max(0, ctl.iter - it0 - 1) * h->admm_aff_kernels;
        } else {   // host loop (PDPLQR_ADMM_GRAPH=0 / pdplqr_admm_configure, or graph capture unavailable)
            bool first = true;
            do {
                rc = admm_iteration(*h, r, first, cudaGraphConditionalHandle{}, 0);
                if (rc) return rc;
                first = false;
                const int it = ctl.iter + 1;
                const bool check = (it % check_every == 0) || it >= max_iter;
                if (check) {
                    CU_TRY(h, cudaMemcpyAsync(&ctl, d_ctl, sizeof(ctl), cudaMemcpyDeviceToHost, h->stream));
                    CU_TRY(h, cudaStreamSynchronize(h->stream));
                } else {
                    ctl.iter = it;
                    ctl.cont = it < max_iter;
                }
            } while (ctl.cont);
        }
        if (!ctl.rho_update) break;
        // rho rescale requested by the check (OSQP rule): rescale, clear the request, run on (re-factorising first)
        const long long n = (long long)h->batch * h->nc_total;
        admm_rho_scale_kernel<<<ew_blocks(n), 256, 0, h->stream>>>(h->d_rho_work, h->d_invrho_work, n, ctl.rho_scale);
        h->launches++;
        CU_TRY(h, cudaGetLastError());
        ctl.rho_update = 0; ctl.n_rho_updates++; ctl.rho_scale = 1.0;
        h->admm_rho_updates_last = ctl.n_rho_updates;
        CU_TRY(h, cudaMemcpyAsync(d_ctl, &ctl, sizeof(ctl), cudaMemcpyHostToDevice, h->stream));
        CU_TRY(h, cudaStreamSynchronize(h->stream));   // (ctl is a stack object)
    }
    if (use_graph) {   // the graph ran backward / forward pairs: leave the protocol state as after a completed solve
        h->factorized = true; h->backward_done = false; h->updated = false;
        h->cur_ws = r.w; h->cur_ys = r.y; h->cur_zs = r.z; h->cur_inv_rho = h->d_invrho_work; h->cur_rho = h->d_rho_work;
        h->sigma = sigma;
        h->have_root = false;
    }
    if (h->padded) {
        rc = repack_ws(*h, h->d_wk, w, false);
        if (rc) return rc;
    }
    if (iters_out) *iters_out = ctl.iter;
    if (residuals_out) { residuals_out[0] = ctl.res[0]; residuals_out[1] = ctl.res[1]; }
    return PDPLQR_OK;
}

int pdplqr_admm_solve(pdplqr_handle_t h, const double* x0, double* ws, double* zs, double* ys, const double* rho,
                      double sigma, double alpha, int max_iter, double eps_abs, double eps_rel, int check_every,
                      int* iters_out, double* residuals_out) {
    if (!h || !x0 || !ws || !zs || !ys || !rho || max_iter < 1)
        return fail(h, PDPLQR_ERR_INVALID, "admm_solve: bad arguments");
    if (!h->cones_set) return fail(h, PDPLQR_ERR_ORDER, "admm_solve before admm_set_cones");
    cudaSetDevice(h->device);
    const size_t B = h->batch, nct = (size_t)h->nc_total, wsl = (size_t)h->N * h->su + h->nxu;   // caller's layout
    CU_TRY(h, cudaMemcpyAsync(h->d_w, ws, B * wsl * 8, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(h->d_z, zs, B * nct * 8, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(h->d_y, ys, B * nct * 8, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(h->d_rho_admm, rho, B * nct * 8, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(h->d_x0, x0, B * h->nxu * 8, cudaMemcpyHostToDevice, h->stream));
    // inv_rho = 1 / rho, as the caller of the reference protocol computes it (lqr_example.cpp:42-43): on the device (IEEE
    // division, the same bits; on the host it was a 370 MB temporary, 46 M divisions and a pageable upload per C4 solve:
    // 200 of the 590 ms of the host-buffer call)
    admm_inv_kernel<<<ew_blocks((long long)(B * nct)), 256, 0, h->stream>>>(h->d_rho_admm, h->d_invrho_admm, (long long)(B * nct));
    h->launches++;
    CU_TRY(h, cudaGetLastError());
    int rc = pdplqr_admm_solve_device(h, h->d_x0, h->d_w, h->d_z, h->d_y, h->d_rho_admm, h->d_invrho_admm, sigma, alpha,
                                      max_iter, eps_abs, eps_rel, check_every, iters_out, residuals_out);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(ws, h->d_w, B * wsl * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaMemcpyAsync(zs, h->d_z, B * nct * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaMemcpyAsync(ys, h->d_y, B * nct * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return PDPLQR_OK;
}

}  // extern "C"
