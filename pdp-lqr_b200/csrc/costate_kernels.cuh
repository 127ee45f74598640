// Costates (dual variables of the dynamics) of the last solve -- SURVEY.md section 8(f) item 2.
//
// The reference has this step written but commented out: `lambda+ = Lxx+ (Lxx+^T x+) + p+` in
// /root/reference include/clqr/lqr/lqr_kernel.hpp:205-211 and `... + F+^T uhat` inside non-last segments in
// lqr_kernel_parallel.hpp:207-216, i.e. lambda_k = P_k x_k + p_k (+ F_k^T uhat).  The stage kernels here do not keep
// P_k, F_k per stage (they are carried in shared memory), so the same quantity is recovered from the stationarity
// condition of the stage instead, segment-parallel, starting at the interface costates the tree already produced:
//        lambda_{N1} = uhat_seg                       (exit of a non-last segment, condensed_system.hpp:140-146)
//        lambda_N    = H~_N x_N + h~_N                (true terminal stage)
//        lambda_k    = [H~_k w_k + h~_k]_x + A_k^T lambda_{k+1}          k = N1-1 ... N0
// with the ADMM-augmented data of the last update_problem_data / backward (lqr_solver_parallel.hpp:129-137,
// lqr_kernel.hpp:106-112):  H~ w + h~ = H w + h + sigma (w - w_prev) + D^T (rho o (D w - z) + y).
// lambda_k is the multiplier of x_k = A x_{k-1} + B u_{k-1} + c_{k-1} in the sign convention
// L = cost + sum_k lambda_{k+1}^T (E_k w_k + c_k - x_{k+1});  output lam[b][k-1] = lambda_k, k = 1..N.
// One warp per (problem, segment); not a hot path (plain loads, no staging).
#pragma once
#include "batch_kernels.cuh"
#include "seg_kernels.cuh"

namespace pdplqr {

struct CostateParams {
    SegParams sp;            // model, partition, ADMM vectors of the last update, interface costates
    const double* traj;      // [batch][N*S+NX]  the solution returned by forward
    const double* lam_root;  // [batch][NX] exit costate of an interior horizon shard's slice, or nullptr
    double* lam;             // [batch][N][NX]
};

template <int NX, int NU>
__global__ void __launch_bounds__(32) seg_costate_kernel(CostateParams q) {
    using D = SegDims<NX, NU>;
    constexpr int S = D::S;
    const SegParams& p = q.sp;
    extern __shared__ double smem[];          // lam_next[NX] | w[S] | qrow[ncmax]
    double* ln = smem;
    double* w = smem + NX;
    double* qr = w + S;
    const int lane = threadIdx.x;
    const int b = blockIdx.x / p.S, seg = blockIdx.x % p.S;
    const int N0 = seg_first(p, seg), N1 = seg_first(p, seg + 1);
    const bool is_last = (seg == p.S - 1);
    const size_t ws_len = (size_t)p.N * S + NX;
    const double* tr = q.traj + (size_t)b * ws_len;
    const double* wprev = p.ws_prev ? p.ws_prev + (size_t)b * ws_len : nullptr;
    const size_t cb = (size_t)b * p.nc_total;
    const double* Db = p.Dm ? p.Dm + (size_t)b * p.d_total : nullptr;
    double* lam_b = q.lam + (size_t)b * p.N * NX;

    // constraint rows of stage k (dimension dim = S, or NX at the terminal stage): qr[r] = rho_r (D_r v - z_r) + y_r
    auto con_rows = [&](int k, int dim, const double* v) {
        const int nck = p.ncmax > 0 ? p.ncs[k] : 0;
        const size_t co = cb + (nck > 0 ? p.coff[k] : 0);
        const double* Dk = nck > 0 ? Db + p.doff[k] : nullptr;
        for (int r = lane; r < nck; r += 32) {
            double acc = 0.0;
            for (int j = 0; j < dim; ++j) acc = fma(Dk[r + (size_t)j * nck], v[j], acc);
            qr[r] = p.rho[co + r] * (acc - p.zs[co + r]) + p.ys[co + r];
        }
        __syncwarp();
        return nck;
    };

    // exit costate of the segment
    if (is_last && !p.interior) {             // lambda_N = H~_N x_N + h~_N
        const double* xN = tr + (size_t)p.N * S;
        for (int i = lane; i < NX; i += 32) w[i] = xN[i];
        __syncwarp();
        const int nck = con_rows(p.N, NX, w);
        const double* Dk = nck > 0 ? Db + p.doff[p.N] : nullptr;
        for (int i = lane; i < NX; i += 32) {
            double acc = p.hN[(size_t)b * NX + i] + p.sigma * (w[i] - (wprev ? wprev[(size_t)p.N * S + i] : 0.0));
            for (int j = 0; j < NX; ++j) acc = fma(p.HN[(size_t)b * NX * NX + i + (size_t)j * NX], w[j], acc);
            for (int r = 0; r < nck; ++r) acc = fma(Dk[r + (size_t)i * nck], qr[r], acc);
            ln[i] = acc;
            lam_b[(size_t)(p.N - 1) * NX + i] = acc;
        }
    } else {
        const double* src = (is_last && q.lam_root) ? q.lam_root + (size_t)b * NX : p.uhat + ((size_t)b * p.S + seg) * NX;
        for (int i = lane; i < NX; i += 32) {
            ln[i] = src[i];
            if (is_last) lam_b[(size_t)(p.N - 1) * NX + i] = src[i];   // interior shard: lambda at the slice's exit
        }
    }
    __syncwarp();
#pragma unroll 1
    for (int k = N1 - 1; k >= N0 && k >= 1; --k) {
        const double* R = p.model + ((size_t)b * p.N + k) * D::REC;
        for (int j = lane; j < S; j += 32) w[j] = tr[(size_t)k * S + j];
        __syncwarp();
        const int nck = con_rows(k, S, w);
        const double* Dk = nck > 0 ? Db + p.doff[k] : nullptr;
        double out[(NX + 31) / 32];
#pragma unroll
        for (int t = 0; t < (NX + 31) / 32; ++t) {
            const int i = lane + 32 * t;
            if (i < NX) {
                const int row = NU + i;
                double acc = R[D::h_at(row)] + p.sigma * (w[row] - (wprev ? wprev[(size_t)k * S + row] : 0.0));
                for (int j = 0; j < S; ++j) acc = fma(R[D::H_at(row, j)], w[j], acc);      // [H w]_x
                for (int j = 0; j < NX; ++j) acc = fma(R[D::er(j) + D::wi(row) * NX], ln[j], acc);           // A^T lambda+
                for (int r = 0; r < nck; ++r) acc = fma(Dk[r + (size_t)row * nck], qr[r], acc);        // [D^T q]_x
                out[t] = acc;
            }
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < (NX + 31) / 32; ++t) {
            const int i = lane + 32 * t;
            if (i < NX) {
                ln[i] = out[t];
                lam_b[(size_t)(k - 1) * NX + i] = out[t];
            }
        }
        __syncwarp();
    }
}

// Thread-per-problem version (batches of tiny systems, one segment, no constraints: batch_kernels.cuh): the same
// recursion with the whole state of a problem in one thread's registers; the stage records are read straight from the
// tile-interleaved blocks (adjacent lanes read adjacent 16-byte pairs: coalesced).  Not a hot path.
template <int NX, int NU>
__global__ void __launch_bounds__(128) batch_costate_kernel(CostateParams q) {
    using B = BatchDims<NX, NU>;
    constexpr int S = NX + NU;
    const SegParams& p = q.sp;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.batch) return;
    const size_t tile = (size_t)(b >> 5);
    const int lane = (int)(b & 31);
    const size_t ws_len = (size_t)p.N * S + NX;
    const double* tr = q.traj + (size_t)b * ws_len;
    const double* wprev = p.ws_prev ? p.ws_prev + (size_t)b * ws_len : nullptr;
    const double* model_t = p.model + tile * p.N * (B::TREC * 32);
    double* lam_b = q.lam + (size_t)b * p.N * NX;
    double ln[NX];
    {   // lambda_N = (H_N + sigma I) x_N + h_N - sigma w_prev_N
        const double* xN = tr + (size_t)p.N * S;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double acc = p.hN[(size_t)b * NX + i] + p.sigma * (xN[i] - (wprev ? wprev[(size_t)p.N * S + i] : 0.0));
#pragma unroll
            for (int j = 0; j < NX; ++j) acc = fma(p.HN[(size_t)b * NX * NX + i + j * NX], xN[j], acc);
            ln[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) lam_b[(size_t)(p.N - 1) * NX + i] = ln[i];
    }
#pragma unroll 1
    for (int k = p.N - 1; k >= 1; --k) {
        const double* blk = model_t + (size_t)k * (B::TREC * 32);
        auto ld = [&](int e) { return blk[tile_pos(e, lane)]; };
        double w[S];
#pragma unroll
        for (int j = 0; j < S; ++j) w[j] = tr[(size_t)k * S + j];
        double out[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const int row = NU + i;
            double acc = ld(B::TR_h + row) + p.sigma * (w[row] - (wprev ? wprev[(size_t)k * S + row] : 0.0));
#pragma unroll
            for (int j = 0; j < S; ++j) acc = fma(ld(B::TR_H + (row >= j ? B::hl(row, j) : B::hl(j, row))), w[j], acc);
#pragma unroll
            for (int j = 0; j < NX; ++j) acc = fma(ld(B::TR_E + j + row * NX), ln[j], acc);   // A^T lambda+
            out[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            ln[i] = out[i];
            lam_b[(size_t)(k - 1) * NX + i] = out[i];
        }
    }
}

}  // namespace pdplqr
