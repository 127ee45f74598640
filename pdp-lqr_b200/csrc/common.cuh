// Shared device helpers for the pdplqr sm_100a kernels: group-cooperative small dense FP64 products with
// per-thread register tiles (operands staged in shared memory), TMA 1-D bulk copies (cp.async.bulk) completed
// on mbarriers, and small utilities.  A "group" is the set of T threads that owns one (problem, segment):
// one warp (T = 32, __syncwarp) or a whole CTA (T = blockDim.x > 32, __syncthreads).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace pdplqr {

#define PDPLQR_DEVINL __device__ __forceinline__

template <int T>
PDPLQR_DEVINL void group_sync() {
    if constexpr (T == 32) __syncwarp();
    else __syncthreads();
}

PDPLQR_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier + TMA 1-D bulk copy (G2S / S2G)
PDPLQR_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
PDPLQR_DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
PDPLQR_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
PDPLQR_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.  SASS: UBLKCP.
PDPLQR_DEVINL void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (bulk async-group completion).
PDPLQR_DEVINL void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
PDPLQR_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
PDPLQR_DEVINL void bulk_wait_read() {  // source smem of all but the N newest groups may be reused
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
PDPLQR_DEVINL void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// order generic-proxy smem accesses before subsequent async-proxy (TMA) accesses of the same locations
PDPLQR_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- group GEMM with register tiles
// C(i,j) = epi(i, j, sum_k A(i,k) * B(k,j))   for i < M, j < N.
// Tiles of TM x TN outputs are dealt round-robin to the T threads of the group (tile index fastest along i, so
// neighbouring threads read neighbouring rows of A: conflict-free for column-major A, broadcast for B).
// LA(i,k), LB(k,j) are element loaders (any layout / transposition / scaling), EPI stores.
template <int M, int N, int K, int TM, int TN, int T, class LA, class LB, class EPI>
PDPLQR_DEVINL void group_mm(int tid, LA la, LB lb, EPI epi) {
    constexpr int MT = (M + TM - 1) / TM;
    constexpr int NT = (N + TN - 1) / TN;
    constexpr int TILES = MT * NT;
    constexpr bool FULL_M = (M % TM) == 0, FULL_N = (N % TN) == 0;
#pragma unroll 1
    for (int t = tid; t < TILES; t += T) {
        const int ti = t % MT, tj = t / MT;
        const int i0 = ti * TM, j0 = tj * TN;
        double acc[TM][TN];
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c) acc[r][c] = 0.0;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            double av[TM], bv[TN];
#pragma unroll
            for (int r = 0; r < TM; ++r) av[r] = la(FULL_M ? i0 + r : min(i0 + r, M - 1), k);
#pragma unroll
            for (int c = 0; c < TN; ++c) bv[c] = lb(k, FULL_N ? j0 + c : min(j0 + c, N - 1));
#pragma unroll
            for (int r = 0; r < TM; ++r)
#pragma unroll
                for (int c = 0; c < TN; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
        }
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c)
                if ((FULL_M || i0 + r < M) && (FULL_N || j0 + c < N)) epi(i0 + r, j0 + c, acc[r][c]);
    }
}

// Same as group_mm with a run-time contraction length K (constraint fold-in: K = nc of the stage).
template <int M, int N, int TM, int TN, int T, class LA, class LB, class EPI>
PDPLQR_DEVINL void group_mm_rt(int tid, int K, LA la, LB lb, EPI epi) {
    constexpr int MT = (M + TM - 1) / TM;
    constexpr int NT = (N + TN - 1) / TN;
    constexpr int TILES = MT * NT;
    constexpr bool FULL_M = (M % TM) == 0, FULL_N = (N % TN) == 0;
#pragma unroll 1
    for (int t = tid; t < TILES; t += T) {
        const int ti = t % MT, tj = t / MT;
        const int i0 = ti * TM, j0 = tj * TN;
        double acc[TM][TN];
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c) acc[r][c] = 0.0;
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            double av[TM], bv[TN];
#pragma unroll
            for (int r = 0; r < TM; ++r) av[r] = la(FULL_M ? i0 + r : min(i0 + r, M - 1), k);
#pragma unroll
            for (int c = 0; c < TN; ++c) bv[c] = lb(k, FULL_N ? j0 + c : min(j0 + c, N - 1));
#pragma unroll
            for (int r = 0; r < TM; ++r)
#pragma unroll
                for (int c = 0; c < TN; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
        }
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c)
                if ((FULL_M || i0 + r < M) && (FULL_N || j0 + c < N)) epi(i0 + r, j0 + c, acc[r][c]);
    }
}

// G products of identical shape in ONE pass (tiles of all G products dealt to the group together, so that no
// lane idles between small products):  C_g(i,j) = epi(g, i, j, sum_k A_g(i,k) B_g(k,j)).
template <int G, int M, int N, int K, int TM, int TN, int T, class LA, class LB, class EPI>
PDPLQR_DEVINL void group_mm_multi(int tid, LA la, LB lb, EPI epi) {
    constexpr int MT = (M + TM - 1) / TM;
    constexpr int NT = (N + TN - 1) / TN;
    constexpr int TILES = MT * NT;
    constexpr bool FULL_M = (M % TM) == 0, FULL_N = (N % TN) == 0;
#pragma unroll 1
    for (int t = tid; t < G * TILES; t += T) {
        const int g = t / TILES, tl = t - g * TILES;
        const int ti = tl % MT, tj = tl / MT;
        const int i0 = ti * TM, j0 = tj * TN;
        double acc[TM][TN];
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c) acc[r][c] = 0.0;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            double av[TM], bv[TN];
#pragma unroll
            for (int r = 0; r < TM; ++r) av[r] = la(g, FULL_M ? i0 + r : min(i0 + r, M - 1), k);
#pragma unroll
            for (int c = 0; c < TN; ++c) bv[c] = lb(g, k, FULL_N ? j0 + c : min(j0 + c, N - 1));
#pragma unroll
            for (int r = 0; r < TM; ++r)
#pragma unroll
                for (int c = 0; c < TN; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
        }
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c)
                if ((FULL_M || i0 + r < M) && (FULL_N || j0 + c < N)) epi(g, i0 + r, j0 + c, acc[r][c]);
    }
}

// Cholesky of the leading NU x NU block of a column-major matrix in shared memory (leading dimension ld),
// right-looking, cooperative over the group.  On exit the lower triangle holds L and dinv[k] = 1 / L(k,k).
// Returns (to every thread) 0 or the 1-based index of the first non-positive pivot (factorisation continues
// with |pivot| so that the kernel stays finite; the status is reported through the C ABI).
template <int NU, int T>
PDPLQR_DEVINL int group_chol(int tid, double* A, int ld, double* dinv) {
    int bad = 0;
#pragma unroll 1
    for (int k = 0; k < NU; ++k) {
        double akk = A[k + k * ld];
        if (!(akk > 0.0)) {
            if (!bad) bad = k + 1;
            akk = fabs(akk) + 1e-300;
        }
        const double r = rsqrt(akk);
        group_sync<T>();  // everyone has read A(k,k)
        for (int i = k + tid; i < NU; i += T) {
            if (i == k) { A[k + k * ld] = akk * r; dinv[k] = r; }
            else A[i + k * ld] *= r;
        }
        group_sync<T>();
        // trailing update, lower part: A(i,j) -= A(i,k) A(j,k), k < j <= i < NU
        constexpr int NN = NU * NU;
        for (int e = tid; e < NN; e += T) {
            const int i = e % NU, j = e / NU;
            if (j > k && i >= j) A[i + j * ld] -= A[i + k * ld] * A[j + k * ld];
        }
        group_sync<T>();
    }
    return bad;
}

}  // namespace pdplqr
