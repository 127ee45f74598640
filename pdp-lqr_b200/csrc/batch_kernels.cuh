// Thread-per-problem kernels for batches of tiny LQ problems (nx + nu <= 8, one segment per problem):
// the whole Riccati state of a problem lives in one thread's registers.  The records of the 32 problems of a warp
// ("tile") are interleaved in HBM, so ONE TMA 1-D bulk copy (cp.async.bulk) per warp and stage streams a fully
// contiguous 11 KB block into shared memory, in a DEPTH-deep ring completed on one mbarrier per ring slot.
//
// Replaces, for the batched sequential case (BASELINE.json config 3):
//   LQRSolver::update_problem_data / backward / forward      /root/reference include/clqr/lqr/lqr_solver.hpp:41-77
//   LQRKernel::step_with_factorization / forward_step        lqr_kernel.hpp:103-147, :180-204
//
// Device record of one stage on this path ("thread record", written by pack_model_kernel):
//   [E (NX x S) | c (NX) | lower(H) packed by columns (S(S+1)/2) | h (S)]   -- H is symmetric (lqr_model.hpp:18),
// so only its lower triangle is kept in HBM: 44 instead of 54 doubles per stage at nx=4, nu=1.
// Tile layout: element e of problem (tile*32 + lane) at stage k lives at
//   ((tile*N + k)*TREC + (e & ~1))*32 + lane*2 + (e & 1)
// i.e. element PAIRS are interleaved across the 32 lanes: a lane reads its pair with one conflict-free 128-bit
// shared load, the forward kernel copies only the leading [E | c] part of every block (also contiguous), and the
// factor records [K | d] use the same tiling, so their stores and loads are fully coalesced.
#pragma once
#include "common.cuh"
#include "seg_kernels.cuh"

namespace pdplqr {

template <int NX, int NU>
struct BatchDims {
    static constexpr int S = NX + NU;
    static constexpr bool ENABLED = S <= 8;
    static constexpr int FRECT = even_up(NU * (NX + 1));  // compact factor record [K | d] (no Gt: single segment)
    static constexpr int TR_E = 0;
    static constexpr int TR_C = NX * S;
    static constexpr int TR_H = TR_C + NX;                 // packed lower triangle, column by column
    static constexpr int TR_h = TR_H + S * (S + 1) / 2;
    static constexpr int TREC = even_up(TR_h + S);
    static constexpr int TREC_EC = even_up(NX * S + NX);
    // index of H(i,j), i >= j, in the packed lower triangle
    static constexpr int hl(int i, int j) { return j * S - j * (j - 1) / 2 + (i - j); }
};

// position of element e of lane `lane` inside a tile block (in doubles)
PDPLQR_DEVINL constexpr int tile_pos(int e, int lane) { return (e & ~1) * 32 + lane * 2 + (e & 1); }

template <int NX, int NU, int WARPS, int DEPTH>
struct BatchBwdSmem {
    static constexpr int BLOCK = BatchDims<NX, NU>::TREC * 32;   // doubles of one tile block (one warp-stage)
    static constexpr int WARP_DOUBLES = DEPTH * BLOCK;
    static constexpr int o_bar = WARPS * WARP_DOUBLES;
    static constexpr size_t BYTES = (size_t)(o_bar + DEPTH * WARPS) * 8;
};

template <int NX, int NU, int WARPS, int DEPTH, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) batch_backward_kernel(SegParams p) {
    using D = SegDims<NX, NU>;
    using B = BatchDims<NX, NU>;
    using L = BatchBwdSmem<NX, NU, WARPS, DEPTH>;
    constexpr int S = D::S;
    constexpr int FRECT = B::FRECT;
    constexpr uint32_t REC_BYTES = L::BLOCK * 8;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long bg = (long long)blockIdx.x * (WARPS * 32) + threadIdx.x;
    const bool active = bg < p.batch;
    const size_t b = active ? (size_t)bg : (size_t)(p.batch - 1);  // idle lanes shadow the last problem

    double* slots = smem + warp * L::WARP_DOUBLES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::o_bar) + DEPTH * warp;
    const size_t ws_len = (size_t)p.N * S + NX;
    const size_t tile = (size_t)(blockIdx.x * WARPS + warp);
    const double* model_t = p.model + tile * p.N * L::BLOCK;
    const double* ws_b = p.ws_prev ? p.ws_prev + b * ws_len : nullptr;
    double* fac_t = p.fac + tile * p.N * (FRECT * 32);
    const double sigma = p.sigma;
    const int N = p.N;
    const bool tile_live = (long long)tile * 32 < p.batch;   // warps past the batch issue no copies and exit early
    if (!tile_live) return;
    auto fetch = [&](int stage, int ring_slot) {   // lane 0: one bulk copy of the whole tile block
        if (lane == 0) {
            mbar_expect_tx(&bar[ring_slot], REC_BYTES);
            bulk_g2s(slots + ring_slot * L::BLOCK, model_t + (size_t)stage * L::BLOCK, REC_BYTES, &bar[ring_slot]);
        }
    };

    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) mbar_init(&bar[d], 1);
        mbar_fence_init();
    }
    __syncwarp();
    // prologue: fill DEPTH-1 ring slots (stages N-1, N-2, ...)
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) {
        if (d < N) fetch(N - 1 - d, d);
    }

    // terminal condition  (lqr_kernel.hpp:79-91):  P_N = H_N + sigma I,  p_N = h_N - sigma w_N
    double P[NX][NX], pv[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
        for (int i = 0; i < NX; ++i) P[i][j] = p.HN[b * NX * NX + i + j * NX] + ((i == j) ? sigma : 0.0);
#pragma unroll
    for (int i = 0; i < NX; ++i) pv[i] = p.hN[b * NX + i] - (ws_b ? sigma * ws_b[(size_t)N * S + i] : 0.0);
    // w_prev of the first stage; later stages are fetched one stage ahead (register prefetch)
    double wnext[S];
#pragma unroll
    for (int i = 0; i < S; ++i) wnext[i] = ws_b ? ws_b[(size_t)(N - 1) * S + i] : 0.0;

    int bad = 0;
    int slot = 0, phase = 0;          // ring position of stage `it`
    int pslot = DEPTH - 1;            // ring position the next prefetch goes to
#pragma unroll 1
    for (int it = 0; it < N; ++it) {
        const int k = N - 1 - it;
        if (it + DEPTH - 1 < N) {
            __syncwarp();  // every lane has finished with ring slot `pslot` (consumed in the previous iteration)
            fetch(k - (DEPTH - 1), pslot);
        }
        pslot = (pslot + 1 == DEPTH) ? 0 : pslot + 1;
        double wprev[S];
#pragma unroll
        for (int i = 0; i < S; ++i) wprev[i] = wnext[i];
        if (k > 0) {
#pragma unroll
            for (int i = 0; i < S; ++i) wnext[i] = ws_b ? ws_b[(size_t)(k - 1) * S + i] : 0.0;
        }
        mbar_wait(&bar[slot], phase);
        const double2* r2 = reinterpret_cast<const double2*>(slots + slot * L::BLOCK) + lane;
        slot = (slot + 1 == DEPTH) ? 0 : slot + 1;
        phase ^= (slot == 0);
        auto ld = [&](int e) {   // pair (e>>1) of this lane: one 128-bit load, 32 lanes x 16 B contiguous
            const double2 v = r2[(e >> 1) * 32];
            return (e & 1) ? v.y : v.x;
        };
        // E (and c as column S) into registers
        double Ea[NX][S + 1];
#pragma unroll
        for (int j = 0; j <= S; ++j)
#pragma unroll
            for (int i = 0; i < NX; ++i) Ea[i][j] = ld(B::TR_E + i + j * NX);
        // [M | g] = [H + sigma I | h - sigma w] + E^T (P [E c] + [0 p]) : lower triangle of M and the last column,
        // one column of P [E c] at a time
        double M[S][S + 1];
#pragma unroll
        for (int j = 0; j <= S; ++j) {
            double col[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double acc = (j == S) ? pv[i] : 0.0;
#pragma unroll
                for (int q = 0; q < NX; ++q) acc = fma(P[i][q], Ea[q][j], acc);
                col[i] = acc;
            }
#pragma unroll
            for (int i = 0; i < S; ++i) {
                if (j < S && i < j) continue;
                double acc;
                if (j < S) acc = ld(B::TR_H + B::hl(i, j)) + ((i == j) ? sigma : 0.0);
                else acc = fma(-sigma, wprev[i], ld(B::TR_h + i));
#pragma unroll
                for (int q = 0; q < NX; ++q) acc = fma(Ea[q][i], col[q], acc);
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
        {   // factor record [K | d], tile-interleaved: coalesced stores
            double* fk = fac_t + (size_t)k * (FRECT * 32);
#pragma unroll
            for (int c = 0; c <= NX; ++c)
#pragma unroll
                for (int m = 0; m < NU; ++m) fk[tile_pos(m + c * NU, lane)] = Z[m][c];
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
        p.status[b] = bad;   // (one slot per (problem, segment): plain store, nothing to clear between solves)
    }
}

// ------------------------------------------------------------------------------------------------
template <int NX, int NU, int WARPS, int DEPTH>
struct BatchFwdSmem {
    using B = BatchDims<NX, NU>;
    static constexpr int BLOCK = (B::TREC_EC + B::FRECT) * 32;   // [E c (pad)] tile block followed by the [K d] tile block
    // output staging: every lane collects OUT_CH stages of its trajectory (OUT_CH * S contiguous doubles of ws) and
    // writes them with ONE bulk S2G copy instead of OUT_CH * S scattered 8-byte stores
    static constexpr int OUT_CH = 4;
    static constexpr int OUT_STRIDE = even_up(OUT_CH * B::S) + 2;   // per-lane slot (16-byte multiple)
    static constexpr int WARP_DOUBLES = DEPTH * BLOCK + 32 * OUT_STRIDE;
    static constexpr int o_bar = WARPS * WARP_DOUBLES;
    static constexpr size_t BYTES = (size_t)(o_bar + DEPTH * WARPS) * 8;
};

template <int NX, int NU, int WARPS, int DEPTH, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) batch_forward_kernel(SegParams p) {
    using D = SegDims<NX, NU>;
    using B = BatchDims<NX, NU>;
    using L = BatchFwdSmem<NX, NU, WARPS, DEPTH>;
    constexpr int S = D::S;
    constexpr int FRECT = B::FRECT;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long bg = (long long)blockIdx.x * (WARPS * 32) + threadIdx.x;
    const bool active = bg < p.batch;
    const size_t b = active ? (size_t)bg : (size_t)(p.batch - 1);

    double* slots = smem + warp * L::WARP_DOUBLES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::o_bar) + DEPTH * warp;
    const size_t ws_len = (size_t)p.N * S + NX;
    const size_t tile = (size_t)(blockIdx.x * WARPS + warp);
    const double* model_t = p.model + tile * p.N * (B::TREC * 32);
    const double* fac_t = p.fac + tile * p.N * (FRECT * 32);
    double* ws_b = p.ws_out + b * ws_len;
    double* ostage = slots + DEPTH * L::BLOCK + lane * L::OUT_STRIDE;
    // bulk stores need 16-byte aligned global chunks: even ws_len and an aligned base (else plain stores)
    const bool out_bulk = (ws_len % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.ws_out) & 15) == 0);
    constexpr uint32_t TX = L::BLOCK * 8;
    const int N = p.N;
    if ((long long)tile * 32 >= p.batch) return;
    auto fetch = [&](int stage, int ring_slot) {   // lane 0: leading [E | c] part of the record block + factor block
        if (lane == 0) {
            double* dst = slots + ring_slot * L::BLOCK;
            mbar_expect_tx(&bar[ring_slot], TX);
            bulk_g2s(dst, model_t + (size_t)stage * (B::TREC * 32), B::TREC_EC * 32 * 8, &bar[ring_slot]);
            bulk_g2s(dst + B::TREC_EC * 32, fac_t + (size_t)stage * (FRECT * 32), FRECT * 32 * 8, &bar[ring_slot]);
        }
    };

    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) mbar_init(&bar[d], 1);
        mbar_fence_init();
    }
    __syncwarp();
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) {
        if (d < N) fetch(d, d);
    }
    double x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = p.xhat[b * NX + i];  // S == 1: xhat aliases x0

    int slot = 0, phase = 0, pslot = DEPTH - 1;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
        if (k + DEPTH - 1 < N) {
            __syncwarp();
            fetch(k + DEPTH - 1, pslot);
        }
        pslot = (pslot + 1 == DEPTH) ? 0 : pslot + 1;
        mbar_wait(&bar[slot], phase);
        const double2* r2 = reinterpret_cast<const double2*>(slots + slot * L::BLOCK) + lane;
        slot = (slot + 1 == DEPTH) ? 0 : slot + 1;
        phase ^= (slot == 0);
        auto ld = [&](int e) {
            const double2 v = r2[(e >> 1) * 32];
            return (e & 1) ? v.y : v.x;
        };
        double u[NU];
#pragma unroll
        for (int m = 0; m < NU; ++m) {
            double acc = ld(B::TREC_EC + NU * NX + m);  // d
#pragma unroll
            for (int j = 0; j < NX; ++j) acc = fma(ld(B::TREC_EC + m + j * NU), x[j], acc);
            u[m] = acc;
        }
        double xn[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double acc = ld(B::TR_C + i);
#pragma unroll
            for (int m = 0; m < NU; ++m) acc = fma(ld(B::TR_E + i + m * NX), u[m], acc);
#pragma unroll
            for (int j = 0; j < NX; ++j) acc = fma(ld(B::TR_E + i + (NU + j) * NX), x[j], acc);
            xn[i] = acc;
        }
        if (out_bulk && k - (k % L::OUT_CH) + L::OUT_CH <= N) {   // whole chunk inside the horizon: stage + bulk store
            const int q = k % L::OUT_CH;
            if (q == 0) bulk_wait_read<0>();   // the previous chunk's bulk store has finished reading the staging slot
#pragma unroll
            for (int m = 0; m < NU; ++m) ostage[q * S + m] = u[m];
#pragma unroll
            for (int i = 0; i < NX; ++i) ostage[q * S + NU + i] = x[i];
            if (q == L::OUT_CH - 1) {
                fence_proxy_async();
                if (active) bulk_s2g(ws_b + (size_t)(k - q) * S, ostage, L::OUT_CH * S * 8);
                bulk_commit();
            }
        } else if (active) {
            double* wk = ws_b + (size_t)k * S;
#pragma unroll
            for (int m = 0; m < NU; ++m) wk[m] = u[m];
#pragma unroll
            for (int i = 0; i < NX; ++i) wk[NU + i] = x[i];
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = xn[i];
    }
    bulk_wait<0>();   // all bulk stores of this thread are complete before the CTA may exit
    if (active) {
#pragma unroll
        for (int i = 0; i < NX; ++i) ws_b[(size_t)N * S + i] = x[i];
    }
}

}  // namespace pdplqr
