// Thread-per-problem kernels for batches of tiny LQ problems (nx + nu <= 8, one segment per problem):
// the whole Riccati state of a problem lives in one thread's registers; every lane streams its own problem's
// stage records from HBM with TMA 1-D bulk copies (cp.async.bulk) into a private shared-memory slot, double
// buffered on one mbarrier pair per warp.  Records stay in the reference's per-problem, per-stage layout
// ([batch][N][E|c|H|h]) -- no batch interleaving / repacking of the model is needed.
//
// Replaces, for the batched sequential case (BASELINE.json config 3):
//   LQRSolver::update_problem_data / backward / forward      /root/reference include/clqr/lqr/lqr_solver.hpp:41-77
//   LQRKernel::step_with_factorization / forward_step        lqr_kernel.hpp:103-147, :180-204
#pragma once
#include "common.cuh"
#include "seg_kernels.cuh"

namespace pdplqr {

constexpr int BATCH_WARPS = 4;  // warps per CTA

template <int NX, int NU>
struct BatchDims {
    static constexpr bool ENABLED = (NX + NU) <= 8;
    static constexpr int FRECT = even_up(NU * (NX + 1));  // compact factor record [K | d] (no Gt: single segment)
};

// per-lane slot: 16-byte multiples whose count is odd -> 128-bit shared loads of 32 lanes are conflict-free
constexpr int slot_doubles(int rec) { return ((rec / 2) % 2 == 1) ? rec : rec + 2; }

template <int NX, int NU>
struct BatchBwdSmem {
    using D = SegDims<NX, NU>;
    static constexpr int SLOT = slot_doubles(D::REC);
    static constexpr int WARP_DOUBLES = 2 * 32 * SLOT;
    static constexpr int o_bar = BATCH_WARPS * WARP_DOUBLES;
    static constexpr size_t BYTES = (size_t)(o_bar + 2 * BATCH_WARPS) * 8;
};

template <int NX, int NU>
__global__ void __launch_bounds__(BATCH_WARPS * 32) batch_backward_kernel(SegParams p) {
    using D = SegDims<NX, NU>;
    using L = BatchBwdSmem<NX, NU>;
    constexpr int S = D::S;
    constexpr int FRECT = BatchDims<NX, NU>::FRECT;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long bg = (long long)blockIdx.x * (BATCH_WARPS * 32) + threadIdx.x;
    const bool active = bg < p.batch;
    const size_t b = active ? (size_t)bg : (size_t)(p.batch - 1);  // idle lanes shadow the last problem

    double* slots = smem + warp * L::WARP_DOUBLES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::o_bar) + 2 * warp;
    const size_t ws_len = (size_t)p.N * S + NX;
    const double* model_b = p.model + b * p.N * D::REC;
    const double* ws_b = p.ws_prev ? p.ws_prev + b * ws_len : nullptr;
    double* fac_b = p.fac + b * p.N * FRECT;
    const double sigma = p.sigma;

    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncwarp();

    // terminal condition  (lqr_kernel.hpp:79-91):  P_N = H_N + sigma I,  p_N = h_N - sigma w_N
    double P[NX][NX], pv[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
        for (int i = 0; i < NX; ++i) P[i][j] = p.HN[b * NX * NX + i + j * NX] + ((i == j) ? sigma : 0.0);
#pragma unroll
    for (int i = 0; i < NX; ++i) pv[i] = p.hN[b * NX + i] - (ws_b ? sigma * ws_b[(size_t)p.N * S + i] : 0.0);

    const int N = p.N;
    if (lane == 0) mbar_expect_tx(&bar[0], 32 * D::REC * 8);
    __syncwarp();
    bulk_g2s(slots + lane * L::SLOT, model_b + (size_t)(N - 1) * D::REC, D::REC * 8, &bar[0]);

    int bad = 0;
#pragma unroll 1
    for (int it = 0; it < N; ++it) {
        const int k = N - 1 - it;
        const int buf = it & 1;
        if (it + 1 < N) {
            __syncwarp();  // all lanes are done reading buffer buf^1 (previous iteration)
            if (lane == 0) mbar_expect_tx(&bar[buf ^ 1], 32 * D::REC * 8);
            __syncwarp();
            bulk_g2s(slots + ((buf ^ 1) * 32 + lane) * L::SLOT, model_b + (size_t)(k - 1) * D::REC, D::REC * 8,
                     &bar[buf ^ 1]);
        }
        double wprev[S];
#pragma unroll
        for (int i = 0; i < S; ++i) wprev[i] = ws_b ? ws_b[(size_t)k * S + i] : 0.0;
        mbar_wait(&bar[buf], (it >> 1) & 1);
        const double2* r2 = reinterpret_cast<const double2*>(slots + (buf * 32 + lane) * L::SLOT);
        auto ld = [&](int e) {
            const double2 v = r2[e >> 1];
            return (e & 1) ? v.y : v.x;
        };
        // PEa = P [E c] + [0 p]      (NX x (S+1))
        double PE[NX][S + 1];
#pragma unroll
        for (int j = 0; j <= S; ++j) {
            double col[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) col[i] = ld(D::REC_E + i + j * NX);  // c follows E in the record
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double acc = (j == S) ? pv[i] : 0.0;
#pragma unroll
                for (int q = 0; q < NX; ++q) acc = fma(P[i][q], col[q], acc);
                PE[i][j] = acc;
            }
        }
        // [M | g] = [H + sigma I | h - sigma w] + E^T PEa   (lower triangle of M and the last column)
        double M[S][S + 1];
#pragma unroll
        for (int j = 0; j <= S; ++j) {
#pragma unroll
            for (int i = 0; i < S; ++i) {
                if (j < S && i < j) continue;
                double acc;
                if (j < S) acc = ld(D::REC_H + i + j * S) + ((i == j) ? sigma : 0.0);
                else acc = ld(D::REC_h + i) - sigma * wprev[i];
#pragma unroll
                for (int q = 0; q < NX; ++q) acc = fma(ld(D::REC_E + q + i * NX), PE[q][j], acc);
                M[i][j] = acc;
            }
        }
        // Luu = chol(Quu) in place, dinv = 1 / diag
        double dinv[NU];
#pragma unroll
        for (int c = 0; c < NU; ++c) {
            double a = M[c][c];
#pragma unroll
            for (int q = 0; q < c; ++q) a = fma(-M[c][q], M[c][q], a);
            if (!(a > 0.0)) { if (!bad) bad = k + 1; a = fabs(a) + 1e-300; }
            const double r = rsqrt(a);
            dinv[c] = r;
            M[c][c] = a * r;
#pragma unroll
            for (int i = c + 1; i < NU; ++i) {
                double v = M[i][c];
#pragma unroll
                for (int q = 0; q < c; ++q) v = fma(-M[i][q], M[c][q], v);
                M[i][c] = v * r;
            }
        }
        // Y = Luu^-1 [Qux | Qu]  (NU x (NX+1)),  Z = -Luu^-T Y = [K | d]
        double Y[NU][NX + 1], Z[NU][NX + 1];
#pragma unroll
        for (int c = 0; c <= NX; ++c) {
#pragma unroll
            for (int m = 0; m < NU; ++m) {
                double v = (c < NX) ? M[NU + c][m] : M[m][S];
#pragma unroll
                for (int q = 0; q < m; ++q) v = fma(-M[m][q], Y[q][c], v);
                Y[m][c] = v * dinv[m];
            }
#pragma unroll
            for (int m = NU - 1; m >= 0; --m) {
                double v = Y[m][c];
#pragma unroll
                for (int q = m + 1; q < NU; ++q) v = fma(-M[q][m], Z[q][c], v);
                Z[m][c] = v * dinv[m];
            }
#pragma unroll
            for (int m = 0; m < NU; ++m) Z[m][c] = -Z[m][c];
        }
        // P = Qxx - Yx^T Yx ,  p = Qx - Yx^T yu
#pragma unroll
        for (int j = 0; j < NX; ++j) {
#pragma unroll
            for (int i = j; i < NX; ++i) {
                double acc = M[NU + i][NU + j];
#pragma unroll
                for (int m = 0; m < NU; ++m) acc = fma(-Y[m][i], Y[m][j], acc);
                P[i][j] = acc;
                P[j][i] = acc;
            }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double acc = M[NU + i][S];
#pragma unroll
            for (int m = 0; m < NU; ++m) acc = fma(-Y[m][i], Y[m][NX], acc);
            pv[i] = acc;
        }
        if (active) {
            double* fk = fac_b + (size_t)k * FRECT;
#pragma unroll
            for (int c = 0; c <= NX; ++c)
#pragma unroll
                for (int m = 0; m < NU; ++m) fk[m + c * NU] = Z[m][c];
        }
    }
    // value function at the entry (P_0, p_0) -> summary slot, for the accessors
    if (active) {
        double* sm = p.sum + b * D::SREC;
#pragma unroll
        for (int j = 0; j < NX; ++j)
#pragma unroll
            for (int i = 0; i < NX; ++i) sm[D::SUM_P + i + j * NX] = P[i][j];
#pragma unroll
        for (int i = 0; i < NX; ++i) sm[D::SUM_p + i] = pv[i];
        if (bad) p.status[b] = bad;
    }
}

// ------------------------------------------------------------------------------------------------
template <int NX, int NU>
struct BatchFwdSmem {
    using D = SegDims<NX, NU>;
    static constexpr int FRECT = BatchDims<NX, NU>::FRECT;
    static constexpr int SLOT = slot_doubles(D::REC_EC + FRECT);  // [E c (pad) | K d]
    static constexpr int WARP_DOUBLES = 2 * 32 * SLOT;
    static constexpr int o_bar = BATCH_WARPS * WARP_DOUBLES;
    static constexpr size_t BYTES = (size_t)(o_bar + 2 * BATCH_WARPS) * 8;
};

template <int NX, int NU>
__global__ void __launch_bounds__(BATCH_WARPS * 32) batch_forward_kernel(SegParams p) {
    using D = SegDims<NX, NU>;
    using L = BatchFwdSmem<NX, NU>;
    constexpr int S = D::S;
    constexpr int FRECT = L::FRECT;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long bg = (long long)blockIdx.x * (BATCH_WARPS * 32) + threadIdx.x;
    const bool active = bg < p.batch;
    const size_t b = active ? (size_t)bg : (size_t)(p.batch - 1);

    double* slots = smem + warp * L::WARP_DOUBLES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::o_bar) + 2 * warp;
    const size_t ws_len = (size_t)p.N * S + NX;
    const double* model_b = p.model + b * p.N * D::REC;
    const double* fac_b = p.fac + b * p.N * FRECT;
    double* ws_b = p.ws_out + b * ws_len;
    constexpr uint32_t TX = (D::REC_EC + FRECT) * 8;

    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncwarp();
    double x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = p.xhat[b * NX + i];  // S == 1: xhat aliases x0

    const int N = p.N;
    if (lane == 0) mbar_expect_tx(&bar[0], 32 * TX);
    __syncwarp();
    bulk_g2s(slots + lane * L::SLOT, model_b, D::REC_EC * 8, &bar[0]);
    bulk_g2s(slots + lane * L::SLOT + D::REC_EC, fac_b, FRECT * 8, &bar[0]);
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
        const int buf = k & 1;
        if (k + 1 < N) {
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&bar[buf ^ 1], 32 * TX);
            __syncwarp();
            double* dst = slots + ((buf ^ 1) * 32 + lane) * L::SLOT;
            bulk_g2s(dst, model_b + (size_t)(k + 1) * D::REC, D::REC_EC * 8, &bar[buf ^ 1]);
            bulk_g2s(dst + D::REC_EC, fac_b + (size_t)(k + 1) * FRECT, FRECT * 8, &bar[buf ^ 1]);
        }
        mbar_wait(&bar[buf], (k >> 1) & 1);
        const double2* r2 = reinterpret_cast<const double2*>(slots + (buf * 32 + lane) * L::SLOT);
        auto ld = [&](int e) {
            const double2 v = r2[e >> 1];
            return (e & 1) ? v.y : v.x;
        };
        double u[NU];
#pragma unroll
        for (int m = 0; m < NU; ++m) {
            double acc = ld(D::REC_EC + NU * NX + m);  // d
#pragma unroll
            for (int j = 0; j < NX; ++j) acc = fma(ld(D::REC_EC + m + j * NU), x[j], acc);
            u[m] = acc;
        }
        double xn[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double acc = ld(D::REC_C + i);
#pragma unroll
            for (int m = 0; m < NU; ++m) acc = fma(ld(D::REC_E + i + m * NX), u[m], acc);
#pragma unroll
            for (int j = 0; j < NX; ++j) acc = fma(ld(D::REC_E + i + (NU + j) * NX), x[j], acc);
            xn[i] = acc;
        }
        if (active) {
            double* wk = ws_b + (size_t)k * S;
#pragma unroll
            for (int m = 0; m < NU; ++m) wk[m] = u[m];
#pragma unroll
            for (int i = 0; i < NX; ++i) wk[NU + i] = x[i];
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = xn[i];
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < NX; ++i) ws_b[(size_t)N * S + i] = x[i];
    }
}

}  // namespace pdplqr
