// Shared device helpers for the pdplqr sm_100a kernels: group-cooperative small dense FP64 products with
// per-thread register tiles (operands staged in shared memory), TMA 1-D bulk copies (cp.async.bulk) completed
// on mbarriers, and small utilities.  A "group" is the set of T threads that owns one (problem, segment):
// one warp (T = 32, __syncwarp) or a whole CTA (T = blockDim.x > 32, __syncthreads).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>

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
// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may be
// scheduled while its predecessor in the stream (or graph) is still running; pdl_wait() returns when the predecessor grid
// has COMPLETED and its memory operations are visible -- placed before the first global access, the kernel is as safe as
// an ordinary launch and only its launch latency / CTA scheduling overlaps the predecessor's tail.  pdl_trigger() lets the
// NEXT kernel of the chain be scheduled from this point on.  Both are no-ops in a kernel launched without the attribute.
PDPLQR_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
PDPLQR_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// order generic-proxy smem accesses before subsequent async-proxy (TMA) accesses of the same locations
PDPLQR_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// compile-time loop: f(std::integral_constant<int, I>) for I = BEGIN .. END-1.  Indices of register arrays written this way
// are constants by construction (a `#pragma unroll` loop that the compiler declines to unroll completely makes them
// dynamic, and the array goes to local memory: the 10 x 10 factor of the nx30/nu10 stage did).
template <int BEGIN, int END, class F>
PDPLQR_DEVINL void static_for(F&& f) {
    if constexpr (BEGIN < END) {
        f(std::integral_constant<int, BEGIN>{});
        static_for<BEGIN + 1, END>(f);
    }
}

// 1/a to within an ulp or two: hardware seed (2^-23) + two Newton steps (~48 cycles against ~72 for an IEEE division
// and ~130 for rsqrt, scripts/micro/lat_bench.cu); a must be a normal, non-zero number (a pivot)
PDPLQR_DEVINL double rcp_newton(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// ---------------------------------------------------------------- stage-record index permutations (seg_warp_kernel.cuh)
// For the sizes the register-resident warp kernel supports, the stage record keeps [E c], H, h in the order that
// kernel wants them (the record layout is ours -- pack_model_kernel writes it, every reader goes through these):
//   * w-indices x FIRST, then u  ([A B c] instead of [B A c]; H, h likewise): Qxx and F+A then sit at the tile origin
//     of the tensor-core accumulators and become the next P / F without any data movement;
//   * the NX rows of [E c] in "accumulator order": position 4 s + q holds the row that lane q of a quad owns in
//     contraction step s when an accumulator fragment (columns 2q, 2q+1 of an 8-wide tile) is fed back as an operand
//     (rows 0,2,4,6 | 1,3,5,7 of every full group of 8; rows 0,2,1,3 of a trailing group of 4).
// Both are the identity for every other size.
#define PDPLQR_HD __host__ __device__ __forceinline__
// rows of column j of H are stored rotated by 4 (j / 2) when s is a multiple of 16: H is only ever read as the initial value
// of a tensor-core accumulator tile (lane -> row r, columns 2 (lane % 4) + {0, 1}); with such a leading dimension the 16 lanes
// of a half-warp would otherwise hit 4 banks 4 times each.  (Tried for every multiple of 8, i.e. also s = 40: the index
// arithmetic of a rotation modulo 40 cost more than the conflicts, S3 of the nx30/nu10 stage 5,000 -> 7,000 cycles.)
PDPLQR_HD constexpr bool h_rotated(int s) { return s % 16 == 0; }
PDPLQR_HD constexpr bool warp_layout(int nx, int nu) { return nx % 4 == 0 && nx <= 16 && nu <= 8 && nx + nu + 1 <= 24; }
PDPLQR_HD constexpr int erow_pos(int i, int nx, bool on) {       // storage position of row i
    if (!on) return i;
    const int b = i / 8, j = i % 8;
    if (nx - 8 * b == 4) return 8 * b + (((j & 1) << 1) | (j >> 1));
    return 8 * b + 4 * (j & 1) + (j >> 1);
}
PDPLQR_HD constexpr int erow_inv(int p, int nx, bool on) {       // row stored at position p
    if (!on) return p;
    const int b = p / 8, j = p % 8;
    if (nx - 8 * b == 4) return 8 * b + (((j & 1) << 1) | (j >> 1));   // (0,2,1,3) is its own inverse
    return 8 * b + 2 * (j & 3) + (j >> 2);
}
PDPLQR_HD constexpr int widx_pos(int i, int nx, int nu, bool on) { return on ? (i < nu ? nx + i : i - nu) : i; }
PDPLQR_HD constexpr int widx_inv(int p, int nx, int nu, bool on) { return on ? (p < nx ? nu + p : p - nx) : p; }

// smallest leading dimension >= n that is 4 (mod 8)  (see BwdSmem)
constexpr int ld4mod8(int n) { return n + ((4 - n % 8) + 8) % 8; }

// ---------------------------------------------------------------- group GEMM with register tiles
// C(i,j) = epi(i, j, sum_k A(i,k) * B(k,j))   for i < M, j < N.
// Tiles of TM x TN outputs are dealt round-robin to the T threads of the group, tile index fastest along i.  The TM
// rows of a tile are STRIDED (ti, ti + MT, ti + 2 MT, ...): lane l and lane l+1 then read and write adjacent rows, so
// every warp-wide access to a column-major operand or result is one contiguous run (one shared-memory wavefront per
// 128 bytes, no bank conflicts) and B is a broadcast.  Blocked rows (ti TM + r) cost 1.6-2.6x the wavefronts in the
// stage kernels, which are shared-memory-bandwidth bound (profiles/r1_ncu_c5small_seg_backward_t64.txt).
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
        const int j0 = tj * TN;
        double acc[TM][TN];
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c) acc[r][c] = 0.0;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            double av[TM], bv[TN];
#pragma unroll
            for (int r = 0; r < TM; ++r) av[r] = la(FULL_M ? ti + r * MT : min(ti + r * MT, M - 1), k);
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
                if ((FULL_M || ti + r * MT < M) && (FULL_N || j0 + c < N)) epi(ti + r * MT, j0 + c, acc[r][c]);
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
        const int j0 = tj * TN;
        double acc[TM][TN];
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c) acc[r][c] = 0.0;
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            double av[TM], bv[TN];
#pragma unroll
            for (int r = 0; r < TM; ++r) av[r] = la(FULL_M ? ti + r * MT : min(ti + r * MT, M - 1), k);
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
                if ((FULL_M || ti + r * MT < M) && (FULL_N || j0 + c < N)) epi(ti + r * MT, j0 + c, acc[r][c]);
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
        const int j0 = tj * TN;
        double acc[TM][TN];
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c) acc[r][c] = 0.0;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            double av[TM], bv[TN];
#pragma unroll
            for (int r = 0; r < TM; ++r) av[r] = la(g, FULL_M ? ti + r * MT : min(ti + r * MT, M - 1), k);
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
                if ((FULL_M || ti + r * MT < M) && (FULL_N || j0 + c < N)) epi(g, ti + r * MT, j0 + c, acc[r][c]);
    }
}

// ---------------------------------------------------------------- FP64 tensor-core versions (DMMA, mma.sync m8n8k4)
// Same interface as group_mm / group_mm_rt / group_mm_multi.  Measured on B200 (scripts/micro/dmma_bench.cu): one
// m8n8k4 per 16 cycles per SM sub-partition (same 37 TFLOP/s peak as the FP64 FMA pipe), latency 26 cycles -- the
// gain is 8x fewer issued instructions per MAC and two operand loads per 256 MACs, which is what the issue-bound
// segment kernels need.  Fragment layout: A(row) lane l -> A[l/4][l%4]; B(col) lane l -> B[l%4][l/4];
// C lane l -> C[l/4][2(l%4) + {0,1}].  Out-of-range rows / columns / k of the padded 8x8x4 tiles are fed zeros.
PDPLQR_DEVINL void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    // (not volatile: a pure function of its operands -- the compiler may interleave it with shuffle / FMA chains)
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// G products of M x N x K over W warps (W may be a run-time value: the latency-mode tree combines)
template <int G, int M, int N, class LA, class LB, class EPI>
PDPLQR_DEVINL void dmma_tiles(int warp, int W, int lane, int K, LA la, LB lb, EPI epi) {
    constexpr int MT = (M + 7) / 8, NT = (N + 7) / 8, TILES = MT * NT;
    const int r = lane >> 2, q = lane & 3;
    const int KT = (K + 3) >> 2;
#pragma unroll 1
    for (int t = warp; t < G * TILES; t += 2 * W) {   // two independent tiles per pass cover the DMMA latency
        const int t2 = t + W;
        const bool has2 = t2 < G * TILES;             // warp-uniform
        const int g0 = t / TILES, l0 = t - g0 * TILES, g1 = has2 ? t2 / TILES : g0, l1 = has2 ? t2 - g1 * TILES : l0;
        const int i0 = (l0 % MT) * 8, j0 = (l0 / MT) * 8, i1 = (l1 % MT) * 8, j1 = (l1 / MT) * 8;
        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll 3
        for (int kt = 0; kt < KT; ++kt) {
            const int k = kt * 4 + q;
            const bool kin = k < K;
            const double a0 = (kin && i0 + r < M) ? la(g0, i0 + r, k) : 0.0;
            const double b0 = (kin && j0 + r < N) ? lb(g0, k, j0 + r) : 0.0;
            dmma_m8n8k4(c00, c01, a0, b0);
            if (has2) {
                const double a1 = (kin && i1 + r < M) ? la(g1, i1 + r, k) : 0.0;
                const double b1 = (kin && j1 + r < N) ? lb(g1, k, j1 + r) : 0.0;
                dmma_m8n8k4(c10, c11, a1, b1);
            }
        }
        if (i0 + r < M) {
            if (j0 + 2 * q < N) epi(g0, i0 + r, j0 + 2 * q, c00);
            if (j0 + 2 * q + 1 < N) epi(g0, i0 + r, j0 + 2 * q + 1, c01);
        }
        if (has2 && i1 + r < M) {
            if (j1 + 2 * q < N) epi(g1, i1 + r, j1 + 2 * q, c10);
            if (j1 + 2 * q + 1 < N) epi(g1, i1 + r, j1 + 2 * q + 1, c11);
        }
    }
}
// Register-blocked version: a warp owns a block of MB x NB output tiles and loads MB A-fragments + NB B-fragments per
// k-step for MB*NB DMMAs (dmma_tiles: 2 loads per DMMA).  The stage kernels are bound by shared-memory wavefronts
// (an LDS.64 with distinct lane addresses costs 2, scripts/micro/lds_bench.cu), so operand reuse in registers is what
// pays.  Block shape: the cheapest of the candidates under a simple issue + load + latency cost with W warps.
struct DmmaBlock { int mb, nb; };
constexpr DmmaBlock pick_dmma_block(int G, int MT, int NT, int W) {
    DmmaBlock best{1, 1};
    long best_cost = 1L << 60;
    for (int mb = 1; mb <= (MT < 4 ? MT : 4); ++mb)
        for (int nb = 1; nb <= (NT < 4 ? NT : 4); ++nb) {
            if (mb * nb > 9) continue;   // 2 accumulator registers (doubles) per tile and lane
            const int blocks = G * ((MT + mb - 1) / mb) * ((NT + nb - 1) / nb);
            const int rounds = (blocks + W - 1) / W;
            const long cost = (long)rounds * (mb * nb * 16 + (mb + nb) * 10 + (mb * nb == 1 ? 26 : 8));
            if (cost < best_cost) { best_cost = cost; best = DmmaBlock{mb, nb}; }
        }
    return best;
}
template <int G, int M, int N, int MB, int NB, class LA, class LB, class EPI>
PDPLQR_DEVINL void dmma_blocks(int warp, int W, int lane, int K, LA la, LB lb, EPI epi) {
    constexpr int MT = (M + 7) / 8, NT = (N + 7) / 8;
    constexpr int MBT = (MT + MB - 1) / MB, NBT = (NT + NB - 1) / NB, BLOCKS = MBT * NBT;
    const int r = lane >> 2, q = lane & 3;
    const int KT = (K + 3) >> 2;
#pragma unroll 1
    for (int t = warp; t < G * BLOCKS; t += W) {
        const int g = t / BLOCKS, l = t - g * BLOCKS;
        const int i0 = (l % MBT) * (8 * MB), j0 = (l / MBT) * (8 * NB);
        double c[MB][NB][2];
#pragma unroll
        for (int a = 0; a < MB; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) c[a][b][0] = c[a][b][1] = 0.0;
#pragma unroll 2
        for (int kt = 0; kt < KT; ++kt) {
            const int k = kt * 4 + q;
            const bool kin = k < K;
            double af[MB], bf[NB];
#pragma unroll
            for (int a = 0; a < MB; ++a) af[a] = (kin && i0 + 8 * a + r < M) ? la(g, i0 + 8 * a + r, k) : 0.0;
#pragma unroll
            for (int b = 0; b < NB; ++b) bf[b] = (kin && j0 + 8 * b + r < N) ? lb(g, k, j0 + 8 * b + r) : 0.0;
#pragma unroll
            for (int a = 0; a < MB; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) dmma_m8n8k4(c[a][b][0], c[a][b][1], af[a], bf[b]);
        }
#pragma unroll
        for (int a = 0; a < MB; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int i = i0 + 8 * a + r, j = j0 + 8 * b + 2 * q;
                if (i < M) {
                    if (j < N) epi(g, i, j, c[a][b][0]);
                    if (j + 1 < N) epi(g, i, j + 1, c[a][b][1]);
                }
            }
    }
}
template <int G, int M, int N, int T, class LA, class LB, class EPI>
PDPLQR_DEVINL void group_mm_dmma_impl(int tid, int K, LA la, LB lb, EPI epi) {
    constexpr DmmaBlock blk = pick_dmma_block(G, (M + 7) / 8, (N + 7) / 8, T / 32);
    dmma_blocks<G, M, N, blk.mb, blk.nb>(tid >> 5, T / 32, tid & 31, K, la, lb, epi);
}
template <int M, int N, int K, int T, class LA, class LB, class EPI>
PDPLQR_DEVINL void group_mm_dmma(int tid, LA la, LB lb, EPI epi) {
    group_mm_dmma_impl<1, M, N, T>(
        tid, K, [&](int, int i, int k) { return la(i, k); }, [&](int, int k, int j) { return lb(k, j); },
        [&](int, int i, int j, double v) { epi(i, j, v); });
}
template <int M, int N, int T, class LA, class LB, class EPI>
PDPLQR_DEVINL void group_mm_dmma_rt(int tid, int K, LA la, LB lb, EPI epi) {
    group_mm_dmma_impl<1, M, N, T>(
        tid, K, [&](int, int i, int k) { return la(i, k); }, [&](int, int k, int j) { return lb(k, j); },
        [&](int, int i, int j, double v) { epi(i, j, v); });
}
template <int G, int M, int N, int K, int T, class LA, class LB, class EPI>
PDPLQR_DEVINL void group_mm_dmma_multi(int tid, LA la, LB lb, EPI epi) {
    group_mm_dmma_impl<G, M, N, T>(tid, K, la, lb, epi);
}

// [M | g] = S x (S+1) product of which only the tiles on or below the diagonal and the tile column that holds column S are
// formed: M is symmetric and every reader of the stage kernel (L D L^T of Quu, the Qxu rows, the lower triangle of Qxx, the
// g column) stays in that part.  The needed tiles are dealt out in equal contiguous runs, longest rows first, so that the
// warps finish together (the rectangular 3 x 3 blocking left 9 / 9 / 6 / 6 tiles on the four warps at S = 40: 30 tiles formed,
// 20 needed) and a run reuses the A fragment of its row.  Same k order per tile as dmma_blocks: identical results.
// The epilogue gets the two accumulator elements of a lane at once: epi(i, j, v(i,j), v(i,j+1)), j even.
constexpr int sym_lower_tiles(int S) {
    const int MT = (S + 7) / 8, TG = S / 8;
    int n = 0;
    for (int ti = 0; ti < MT; ++ti) n += ti + 1 + (ti < TG ? 1 : 0);
    return n;
}
// tile n of the enumeration "rows from the bottom, row i holds columns 0 .. i and TG": (row, column)
constexpr int sym_tile_row(int S, int n) {
    const int MT = (S + 7) / 8, TG = S / 8;
    for (int i = MT - 1; i >= 0; --i) {
        const int cnt = i + 1 + (i < TG ? 1 : 0);
        if (n < cnt) return i;
        n -= cnt;
    }
    return -1;
}
constexpr int sym_tile_col(int S, int n) {
    const int MT = (S + 7) / 8, TG = S / 8;
    for (int i = MT - 1; i >= 0; --i) {
        const int cnt = i + 1 + (i < TG ? 1 : 0);
        if (n < cnt) return n <= i ? n : TG;
        n -= cnt;
    }
    return -1;
}
// the run of warp WARP: every tile index is a compile-time constant (with run-time tile indices the operand loads sat behind
// warp-uniform branches and were no longer hoisted over the tensor ops: S3 at S = 40 went from 5,100 to 8,700 cycles)
template <int S, int W, int WARP, class LA, class LB, class EPI>
PDPLQR_DEVINL void dmma_sym_lower_run(int lane, int K, LA la, LB lb, EPI epi) {
    constexpr int NEED = sym_lower_tiles(S), CH = (NEED + W - 1) / W;
    constexpr int N0 = WARP * CH;
    constexpr int CNT = NEED - N0 < CH ? (NEED - N0 > 0 ? NEED - N0 : 0) : CH;   // tiles of this warp
    if constexpr (CNT > 0) {
        const int r = lane >> 2, q = lane & 3;
        const int KT = (K + 3) >> 2;
        double c[CNT][2];
        static_for<0, CNT>([&](auto tc) { c[tc][0] = c[tc][1] = 0.0; });
#pragma unroll 2
        for (int kt = 0; kt < KT; ++kt) {
            const int k = kt * 4 + q;
            const bool kin = k < K;
            double af[CNT], bf[CNT];
            static_for<0, CNT>([&](auto tc) {
                constexpr int t = tc;
                constexpr int ti = sym_tile_row(S, N0 + t), tj = sym_tile_col(S, N0 + t);
                constexpr bool fresh = (t == 0) || sym_tile_row(S, N0 + t - 1) != ti;
                if constexpr (fresh) {
                    if constexpr (8 * ti + 7 < S) af[t] = kin ? la(8 * ti + r, k) : 0.0;
                    else af[t] = (kin && 8 * ti + r < S) ? la(8 * ti + r, k) : 0.0;
                } else
                    af[t] = af[t > 0 ? t - 1 : 0];
                if constexpr (8 * tj + 7 < S + 1) bf[t] = kin ? lb(k, 8 * tj + r) : 0.0;
                else bf[t] = (kin && 8 * tj + r < S + 1) ? lb(k, 8 * tj + r) : 0.0;
            });
            static_for<0, CNT>([&](auto tc) { dmma_m8n8k4(c[tc][0], c[tc][1], af[tc], bf[tc]); });
        }
        static_for<0, CNT>([&](auto tc) {
            constexpr int t = tc;
            constexpr int ti = sym_tile_row(S, N0 + t), tj = sym_tile_col(S, N0 + t);
            const int i = 8 * ti + r, j = 8 * tj + 2 * q;
            if (i < S && j <= S) epi(i, j, c[t][0], c[t][1]);   // elements (i, j) and (i, j + 1), j even; (i, j + 1) may be
                                                                // out of range (j + 1 > S): the epilogue checks
        });
    }
}
template <int S, int W, class LA, class LB, class EPI>
PDPLQR_DEVINL void dmma_sym_lower(int warp, int lane, int K, LA la, LB lb, EPI epi) {
    static_for<0, W>([&](auto wc) {
        if (warp == wc) dmma_sym_lower_run<S, W, wc>(lane, K, la, lb, epi);
    });
}

#ifndef PDPLQR_DMMA_MIN_MACS
#define PDPLQR_DMMA_MIN_MACS 500     // products smaller than this stay on the register-tile FMA path.  Measured (C5,
#endif                               // nx12/nu4, same box): 20000 -> 4.22 ms, 3000 -> 3.90 ms, 500 and 1 -> 3.58 ms
// dispatchers used by the kernels
template <int M, int N, int K, int TM, int TN, int T, class LA, class LB, class EPI>
PDPLQR_DEVINL void gmm(int tid, LA la, LB lb, EPI epi) {
    if constexpr ((long long)M * N * K >= PDPLQR_DMMA_MIN_MACS) group_mm_dmma<M, N, K, T>(tid, la, lb, epi);
    else group_mm<M, N, K, TM, TN, T>(tid, la, lb, epi);
}
template <int M, int N, int TM, int TN, int T, class LA, class LB, class EPI>
PDPLQR_DEVINL void gmm_rt(int tid, int K, LA la, LB lb, EPI epi) {
    if constexpr ((long long)M * N * 16 >= PDPLQR_DMMA_MIN_MACS) group_mm_dmma_rt<M, N, T>(tid, K, la, lb, epi);
    else group_mm_rt<M, N, TM, TN, T>(tid, K, la, lb, epi);
}
template <int G, int M, int N, int K, int TM, int TN, int T, class LA, class LB, class EPI>
PDPLQR_DEVINL void gmm_multi(int tid, LA la, LB lb, EPI epi) {
    if constexpr ((long long)M * N * K >= PDPLQR_DMMA_MIN_MACS) group_mm_dmma_multi<G, M, N, K, T>(tid, la, lb, epi);
    else group_mm_multi<G, M, N, K, TM, TN, T>(tid, la, lb, epi);
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
