// Latency-mode versions of the interface-tree kernels (tree_kernels.cuh), used when there are fewer combine groups
// than SMs (one long-horizon problem, horizon shards, the shard coupler): the interface solve is then a chain of
// dependent small dense steps and what matters is the length of that chain, not throughput.
//
// Same mathematics and the same summary / down-sweep records as tree_kernels.cuh (reference: the serial block recursion
// of /root/reference include/clqr/lqr/condensed_system.hpp:82-146), different mapping:
//   * a combine is spread element-parallel over tt = 64 .. 256 threads (run-time tt: the upper tree gives every
//     surviving combine more threads as the level shrinks: 8 x 64, 4 x 128, 2 x 256, 1 x 256);
//   * Gauss-Jordan: every thread owns a fixed set of matrix elements and keeps them in registers over all pivot steps;
//     rows are pivoted implicitly (no exchange, one un-permuting store at the end); the pivot of step k+1 is found by
//     the owners of column k+1 with a shared-memory atomicMax on (magnitude bits | row) keys while they write step k;
//     the two buffers ping-pong (one barrier per step); the reciprocal is a Newton-refined hardware seed;
//   * the down-sweep records are pulled into shared memory by TMA bulk copies before the level loop starts, so the
//     dependent mat-vec chain never waits for L2.
#pragma once
#include "tree_kernels.cuh"

namespace pdplqr {

PDPLQR_DEVINL void rt_sync(int tt, int bar_id) {
    if (tt == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(tt) : "memory");
}

// 1/a to within an ulp or two: hardware seed (2^-23) + two Newton steps; a is a Gauss-Jordan pivot (normal, non-zero)
PDPLQR_DEVINL double rcp_newton(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    return r;
}

template <int NX>
struct LatSmem {
    using D = TreeDims<NX>;
    static_assert(NX <= 32, "pivot keys carry the row index in 5 bits");
    // Gauss-Jordan work items: 4 consecutive rows of one column (one 32-byte shared-memory chunk), columns 1 .. NCOL-1
    static constexpr int NRCH = (NX + 3) / 4;            // 4-row chunks per column
    static constexpr int LDG = 4 * NRCH;                 // leading dimension of the augmented matrix (multiple of 4)
    static constexpr int AUG = LDG * D::NCOL;
    static constexpr int ITEMS = NRCH * (D::NCOL - 1);
    static constexpr int MINTT = NX <= 16 ? 64 : 128;    // smallest group a combine may run on
    static constexpr int MAXQ = (ITEMS + MINTT - 1) / MINTT;   // items (x 4 registers) per thread at most
    static constexpr int WORK = 2 * AUG + NX;            // two Gauss-Jordan buffers (the idle one is the scratch of the
                                                         // later phases: [T1 | T2 | x_f | lv]) + pivot keys / row map
    static_assert(AUG >= 2 * D::N2 + 2 * NX, "phase scratch must fit in one Gauss-Jordan buffer");
    // upper tree: one slot = [a | b | work]
    static constexpr int o_WK = even_up(2 * D::SREC);   // the workspace is accessed with 16-byte vectors
    static constexpr int SLOT = o_WK + even_up(WORK);
    static constexpr int TOP_THREADS = 512;
    static constexpr int TT_CAP = 256;                   // more threads than work items per phase do not help
    static constexpr int fit = (200 * 1024 / 8) / SLOT;
    static constexpr int by_threads = TOP_THREADS / MINTT;
    static constexpr int lim = fit < by_threads ? fit : by_threads;
    static constexpr int TOP_SLOTS = lim >= 8 ? 8 : (lim >= 4 ? 4 : (lim >= 2 ? 2 : (lim >= 1 ? 1 : 0)));
    static constexpr int TOP_NODES = 2 * TOP_SLOTS;      // widest level the one-launch upper tree accepts
    static constexpr size_t TOP_BYTES = (size_t)(TOP_SLOTS > 0 ? TOP_SLOTS : 1) * SLOT * 8;
    // lower levels: [s0 | s1 | s2 | work], one group of UP_TT threads per CTA
    static constexpr int UP_TT = ITEMS > 128 ? 256 : 128;
    static constexpr int o_UPWK = even_up(3 * D::SREC);
    static constexpr size_t UP_BYTES = (size_t)(o_UPWK + even_up(WORK)) * 8;
    static constexpr bool UP_OK = UP_BYTES <= 220 * 1024;
    // down-sweeps: the records of every pair of the upper tree (<= TOP_NODES - 1) + per-level vectors
    static constexpr int DOWN_MAXREC = TOP_NODES > 1 ? TOP_NODES - 1 : 1;
    static constexpr int DOWN_VEC = 2 * NX * (2 * (TOP_NODES > 0 ? TOP_NODES : 1)) + 4 * NX * 16;
    static constexpr size_t DOWN_BYTES = (size_t)(DOWN_MAXREC * D::DREC + DOWN_VEC) * 8;
    static constexpr bool DOWN_OK = TOP_NODES >= 2 && DOWN_BYTES <= 220 * 1024;
};

// pivot key: magnitude bits of v (sign and the 5 lowest bits of the high word dropped) | (31 - row): the largest
// key is the largest |v| up to 2^-15 relative, ties to the smaller row
PDPLQR_DEVINL int piv_key(double v, int row) { return (__double2hiint(v) & 0x7fffffe0) | (31 - row); }

// clears the pivot keys of a combine workspace; once per kernel, followed by a group barrier before combine_lat
template <int NX>
PDPLQR_DEVINL void combine_lat_init(int t, int tt, double* wk) {
    int* pivkey = reinterpret_cast<int*>(wk + 2 * LatSmem<NX>::AUG);
    for (int r = t; r < NX; r += tt) pivkey[r] = 0;
}

// One combine by a group of tt >= LatSmem::MINTT threads (t = index in the group).  sa, sb: member / suffix summaries
// (shared); wk: LatSmem::WORK doubles of shared workspace (16-byte aligned) whose pivot keys are zero (they are again on
// exit); out: combined summary (shared or global, distinct from sa, sb); ddi: down-sweep record of member a (global).
// Ends with a group barrier.
template <int NX>
PDPLQR_DEVINL void combine_lat(int t, int tt, int bar_id, const double* sa, const double* sb, double* wk, double* out,
                               double* ddi) {
    using D = TreeDims<NX>;
    using L = LatSmem<NX>;
    constexpr int N2 = D::N2, LDG = L::LDG, NCOL = D::NCOL, NRCH = L::NRCH, ITEMS = L::ITEMS, MAXQ = L::MAXQ;
    const double *Pa = sa + D::SUM_P, *Fa = sa + D::SUM_F, *Ca = sa + D::SUM_C, *pa = sa + D::SUM_p, *fa = sa + D::SUM_f;
    const double *Pb = sb + D::SUM_P, *Fb = sb + D::SUM_F, *Cb = sb + D::SUM_C, *pb = sb + D::SUM_p, *fb = sb + D::SUM_f;
    const int warp = t >> 5, lane = t & 31, nw = tt >> 5;
    double* cur = wk;
    double* nxt = wk + L::AUG;
    int* pivkey = reinterpret_cast<int*>(wk + 2 * L::AUG);   // [NX] atomicMax targets, one per pivot step
    int* rowof = pivkey + NX;                                // [NX] rowof[i] = k : row i was the pivot row of step k
    // ---- phase 1: Aug = [I + C_a P_b | F_a | C_a | f_a] ; the parts of the record that do not need the inverse
    {
        auto la = [&](int, int r, int k) { return Ca[r + k * NX]; };
        auto lb = [&](int, int k, int c) { return Pb[k + c * NX]; };
        auto epi = [&](int, int r, int c, double v) {
            if (r == c) v += 1.0;
            cur[r + c * LDG] = v;
            if (c == 0) atomicMax(&pivkey[0], piv_key(v, r));    // pivot of step 0
        };
        dmma_tiles<1, NX, NX>(warp, nw, lane, NX, la, lb, epi);
    }
    for (int e = t; e < N2; e += tt) {
        const int r = e % NX, c = e / NX;
        cur[r + (NX + c) * LDG] = Fa[e];
        cur[r + (2 * NX + c) * LDG] = Ca[e];
        ddi[D::DD_PB + e] = Pb[e];
        ddi[D::DD_FB + e] = Fb[e];
    }
    for (int r = t; r < NX; r += tt) {
        cur[r + 3 * NX * LDG] = fa[r];
        ddi[D::DD_pb + r] = pb[r];
    }
    if constexpr (LDG > NX)                                      // padding rows: keep them finite
        for (int e = t; e < NCOL * (LDG - NX); e += tt) cur[NX + e % (LDG - NX) + (e / (LDG - NX)) * LDG] = 0.0;
    rt_sync(tt, bar_id);
    // ---- phase 2: Gauss-Jordan, implicit partial (row) pivoting.  Step k: pivot row piv (largest |a(i,k)| among the
    //      rows not used yet), a(piv, :) /= a(piv,k), every other row i: a(i, :) -= a(i,k) a(piv, :).  Row piv then
    //      holds component k of the solution.  Work item id = (column j = 1 + id / NRCH, rows 4c .. 4c+3, c = id % NRCH)
    //      is the 32-byte chunk NRCH + id of the buffer; its owner keeps the four values in registers for all steps.
    double val[MAXQ][4];
    int ij[MAXQ], ic4[MAXQ];
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) {
        const int id = t + q * tt;
        ij[q] = (id < ITEMS) ? 1 + id / NRCH : 0;            // 0 marks "no item"
        ic4[q] = 4 * (id % NRCH);
        if (ij[q]) {
            const double2* src = reinterpret_cast<const double2*>(cur + ij[q] * LDG + ic4[q]);
            const double2 v01 = src[0], v23 = src[1];
            val[q][0] = v01.x; val[q][1] = v01.y; val[q][2] = v23.x; val[q][3] = v23.y;
        }
    }
    int piv = 31 - (pivkey[0] & 31);
    unsigned used = 0;
#pragma unroll 1
    for (int k = 0; k < NX; ++k) {
        used |= 1u << piv;
        if (t == 0) rowof[piv] = k;
        const double pinv = rcp_newton(cur[piv + k * LDG]);
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            const int j = ij[q], c4 = ic4[q];
            if (j > k) {
                const double2* mp = reinterpret_cast<const double2*>(cur + k * LDG + c4);
                const double2 m01 = mp[0], m23 = mp[1];
                const double tv = cur[piv + j * LDG] * pinv;
                double v0 = fma(-m01.x, tv, val[q][0]), v1 = fma(-m01.y, tv, val[q][1]);
                double v2 = fma(-m23.x, tv, val[q][2]), v3 = fma(-m23.y, tv, val[q][3]);
                const int pr = piv - c4;
                if (pr == 0) v0 = tv;
                if (pr == 1) v1 = tv;
                if (pr == 2) v2 = tv;
                if (pr == 3) v3 = tv;
                val[q][0] = v0; val[q][1] = v1; val[q][2] = v2; val[q][3] = v3;
                double2* dst = reinterpret_cast<double2*>(nxt + j * LDG + c4);
                dst[0] = make_double2(v0, v1);
                dst[1] = make_double2(v2, v3);
                if (j == k + 1 && j < NX) {                  // candidate pivots of the next step
                    int key = -1;
                    if (c4 + 0 < NX && !((used >> (c4 + 0)) & 1u)) key = max(key, piv_key(v0, c4 + 0));
                    if (c4 + 1 < NX && !((used >> (c4 + 1)) & 1u)) key = max(key, piv_key(v1, c4 + 1));
                    if (c4 + 2 < NX && !((used >> (c4 + 2)) & 1u)) key = max(key, piv_key(v2, c4 + 2));
                    if (c4 + 3 < NX && !((used >> (c4 + 3)) & 1u)) key = max(key, piv_key(v3, c4 + 3));
                    if (key >= 0) atomicMax(&pivkey[j], key);
                }
            }
        }
        rt_sync(tt, bar_id);
        double* sw = cur; cur = nxt; nxt = sw;
        if (k + 1 < NX) piv = 31 - (pivkey[k + 1] & 31);
    }
    // un-permute: solution row k is row i with rowof[i] == k; the result goes to the idle buffer
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) {
        const int j = ij[q], c4 = ic4[q];
        if (j >= NX) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
                if (c4 + rr < NX) nxt[rowof[c4 + rr] + j * LDG] = val[q][rr];
        }
    }
    for (int r = t; r < NX; r += tt) pivkey[r] = 0;          // ready for the next combine on this workspace
    rt_sync(tt, bar_id);
    // nxt holds [ . | X_F | X_C | w_f ] ; cur is scratch
    const double* XF = nxt + NX * LDG;
    const double* XC = nxt + 2 * NX * LDG;
    const double* wf = nxt + 3 * NX * LDG;
    double* T1 = cur;              // P_b X_F
    double* T2 = cur + N2;         // F_b X_C
    double* xf = cur + 2 * N2;
    double* lv = xf + NX;
    // ---- phase 3: T1 = P_b X_F ; T2 = F_b X_C ; F = F_b X_F ; x_f = w_f - X_C p_b ; rest of the record
    {
        auto la = [&](int g, int r, int k) { return (g == 0 ? Pb : Fb)[r + k * NX]; };
        auto lb = [&](int g, int k, int c) { return (g == 1 ? XC : XF)[k + c * LDG]; };
        auto epi = [&](int g, int r, int c, double v) {
            if (g == 0) T1[r + c * NX] = v;
            else if (g == 1) T2[r + c * NX] = v;
            else out[D::SUM_F + r + c * NX] = v;
        };
        dmma_tiles<3, NX, NX>(warp, nw, lane, NX, la, lb, epi);
    }
    for (int r = tt - 1 - t; r < NX; r += tt) {   // vector work goes to the group's last threads
        double acc0 = wf[r], acc1 = 0.0;
#pragma unroll
        for (int k = 0; k + 1 < NX; k += 2) {
            acc0 = fma(-XC[r + k * LDG], pb[k], acc0);
            acc1 = fma(-XC[r + (k + 1) * LDG], pb[k + 1], acc1);
        }
        if (NX & 1) acc0 = fma(-XC[r + (NX - 1) * LDG], pb[NX - 1], acc0);
        xf[r] = acc0 + acc1;
        ddi[D::DD_wf + r] = wf[r];
    }
    for (int e = t; e < N2; e += tt) {
        const int r = e % NX, c = e / NX;
        ddi[D::DD_XF + e] = XF[r + c * LDG];
        ddi[D::DD_XC + e] = XC[r + c * LDG];
    }
    rt_sync(tt, bar_id);
    // ---- phase 4: P = P_a + F_a^T T1 ; C = C_b + T2 F_b^T ; lv = P_b x_f + p_b ; f = F_b x_f + f_b
    {
        auto la = [&](int g, int r, int k) { return g == 0 ? Fa[k + r * NX] : T2[r + k * NX]; };
        auto lb = [&](int g, int k, int c) { return g == 0 ? T1[k + c * NX] : Fb[c + k * NX]; };
        auto epi = [&](int g, int r, int c, double v) {
            if (g == 0) out[D::SUM_P + r + c * NX] = Pa[r + c * NX] + v;
            else out[D::SUM_C + r + c * NX] = Cb[r + c * NX] + v;
        };
        dmma_tiles<2, NX, NX>(warp, nw, lane, NX, la, lb, epi);
    }
    for (int r = tt - 1 - t; r < NX; r += tt) {
        double al = pb[r], af = fb[r];
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            al = fma(Pb[r + k * NX], xf[k], al);
            af = fma(Fb[r + k * NX], xf[k], af);
        }
        lv[r] = al;
        out[D::SUM_f + r] = af;
    }
    rt_sync(tt, bar_id);
    // ---- phase 5: p = p_a + F_a^T lv
    for (int r = t; r < NX; r += tt) {
        double ap0 = pa[r], ap1 = 0.0;
#pragma unroll
        for (int k = 0; k + 1 < NX; k += 2) {
            ap0 = fma(Fa[k + r * NX], lv[k], ap0);
            ap1 = fma(Fa[k + 1 + r * NX], lv[k + 1], ap1);
        }
        if (NX & 1) ap0 = fma(Fa[NX - 1 + r * NX], lv[NX - 1], ap0);
        out[D::SUM_p + r] = ap0 + ap1;
    }
    rt_sync(tt, bar_id);
}

// ---------------------------------------------------------------- lower level: one CTA of TT threads per group
template <int NX, int TT>
__global__ void __launch_bounds__(TT) tree_up_lat_kernel(TreeParams p) {
    using D = TreeDims<NX>;
    using L = LatSmem<NX>;
    constexpr int NPRE = (D::SREC + TT - 1) / TT;
    extern __shared__ __align__(16) double smem[];
    const int t = threadIdx.x;
    const int b = blockIdx.x / p.groups, g = blockIdx.x % p.groups;
    const int first = g * p.R, last = min(first + p.R, p.count) - 1;
    double* sa = smem;
    double* sb = smem + D::SREC;
    double* sn = smem + 2 * D::SREC;
    double* wk = smem + L::o_UPWK;
    const double* in_b = p.sum_in + (size_t)b * p.count * D::SREC;
    double* dd_b = p.dd + (size_t)b * p.count * D::DREC;
    double pre[NPRE];
    auto prefetch = [&](int node) {   // into registers: the L2 latency hides behind the running combine
#pragma unroll
        for (int q = 0; q < NPRE; ++q) {
            const int e = t + q * TT;
            if (e < D::SREC) pre[q] = in_b[(size_t)node * D::SREC + e];
        }
    };
    if (last > first) prefetch(last - 1);
    combine_lat_init<NX>(t, TT, wk);
    for (int e = t; e < D::SREC; e += TT) sb[e] = in_b[(size_t)last * D::SREC + e];
#pragma unroll 1
    for (int i = last - 1; i >= first; --i) {
#pragma unroll
        for (int q = 0; q < NPRE; ++q) {
            const int e = t + q * TT;
            if (e < D::SREC) sa[e] = pre[q];
        }
        __syncthreads();
        if (i > first) prefetch(i - 1);
        combine_lat<NX>(t, TT, 1, sa, sb, wk, sn, dd_b + (size_t)i * D::DREC);
        double* sw = sb; sb = sn; sn = sw;
    }
    __syncthreads();
    if (p.sum_out) {
        double* out = p.sum_out + ((size_t)b * p.groups + g) * D::SREC;
        for (int e = t; e < D::SREC; e += TT) out[e] = sb[e];
    }
}

// ---------------------------------------------------------------- upper levels: one CTA per problem, one launch
template <int NX>
__global__ void __launch_bounds__(LatSmem<NX>::TOP_THREADS) tree_top_up_lat_kernel(TreeTopParams p) {
    using D = TreeDims<NX>;
    using L = LatSmem<NX>;
    constexpr int THREADS = L::TOP_THREADS;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
#pragma unroll 1
    for (int l = 0; l + 1 < p.nlevels; ++l) {   // level l (count[l] nodes) -> level l+1 (pairs)
        const int cnt = p.count[l], groups = (cnt + 1) / 2;
        int ng = 1;
        while (ng < groups) ng <<= 1;           // <= TOP_SLOTS (the plan caps count[0] at TOP_NODES)
        int tt = THREADS / ng;
        if (tt > L::TT_CAP) tt = L::TT_CAP;     // a combine has at most ~3 NX^2 independent elements per phase
        const int grp = tid / tt, t = tid % tt;
        if (grp < groups) {
            const double* in_b = p.sum[l] + (size_t)b * cnt * D::SREC;
            double* out = p.sum[l + 1] + ((size_t)b * groups + grp) * D::SREC;
            const int ia = 2 * grp, ib = 2 * grp + 1;
            if (ib >= cnt) {                    // odd node out: passes through unchanged
                for (int e = t; e < D::SREC; e += tt) out[e] = in_b[(size_t)ia * D::SREC + e];
            } else {
                double* ws = smem + (size_t)grp * L::SLOT;
                double* sa = ws;
                double* sb = ws + D::SREC;
                for (int e = t; e < D::SREC; e += tt) {
                    sa[e] = in_b[(size_t)ia * D::SREC + e];
                    sb[e] = in_b[(size_t)ib * D::SREC + e];
                }
                combine_lat_init<NX>(t, tt, ws + L::o_WK);   // slots change owners between levels
                rt_sync(tt, 1 + grp);
                combine_lat<NX>(t, tt, 1 + grp, sa, sb, ws + L::o_WK, out, p.dd[l] + ((size_t)b * cnt + ia) * D::DREC);
            }
        }
        __threadfence_block();
        __syncthreads();
    }
}

// one down-sweep step from a record in shared memory; x, le: entry state / exit costate of the pair (shared);
// writes lam of the first member and x of the second (shared), pt: NX doubles of scratch
template <int NX>
PDPLQR_DEVINL void warp_down_step_smem(int lane, const double* rec, const double* x, const double* le, double* pt,
                                       double* lam_first, double* x_second) {
    using D = TreeDims<NX>;
    for (int r = lane; r < NX; r += 32) {
        double acc = rec[D::DD_pb + r];
#pragma unroll
        for (int k = 0; k < NX; ++k) acc = fma(rec[D::DD_FB + k + r * NX], le[k], acc);   // F_b^T lam_e
        pt[r] = acc;
    }
    __syncwarp();
    for (int r = lane; r < NX; r += 32) {
        double acc0 = rec[D::DD_wf + r], acc1 = 0.0;
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            acc0 = fma(rec[D::DD_XF + r + k * NX], x[k], acc0);
            acc1 = fma(-rec[D::DD_XC + r + k * NX], pt[k], acc1);
        }
        x_second[r] = acc0 + acc1;
    }
    __syncwarp();
    for (int r = lane; r < NX; r += 32) {
        double acc = pt[r];
#pragma unroll
        for (int k = 0; k < NX; ++k) acc = fma(rec[D::DD_PB + r + k * NX], x_second[k], acc);
        lam_first[r] = acc;
    }
    __syncwarp();
}

template <int NX>
__global__ void __launch_bounds__(LatSmem<NX>::TOP_THREADS) tree_top_down_lat_kernel(TreeTopParams p) {
    using D = TreeDims<NX>;
    using L = LatSmem<NX>;
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar[TREE_TOP_MAX_LEVELS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    double* recs = smem;                                   // pair records, upper levels first
    double* vec = smem + L::DOWN_MAXREC * D::DREC;         // per level: x[count][NX] then lam[count][NX]
    double* scratch = vec + 2 * NX * 2 * L::TOP_NODES + warp * 4 * NX;
    const int top = p.nlevels - 1;
    if (tid == 0) {
        for (int l = 0; l < top; ++l) mbar_init(&bar[l], 1);
        mbar_fence_init();
        int off = 0;
        for (int l = top - 1; l >= 0; --l) {               // the order the sweep needs them in
            const int cnt = p.count[l], pairs = cnt / 2;
            if (pairs > 0) mbar_expect_tx(&bar[l], (uint32_t)(pairs * D::DREC * 8));
            for (int g = 0; g < pairs; ++g) {
                bulk_g2s(recs + (size_t)off * D::DREC, p.dd[l] + ((size_t)b * cnt + 2 * g) * D::DREC, D::DREC * 8, &bar[l]);
                ++off;
            }
        }
    }
    // vec offsets: level l starts after all upper levels
    auto vec_off = [&](int l) {
        int o = 0;
        for (int m = top; m > l; --m) o += p.count[m];
        return o * 2 * NX;
    };
    if (warp == 1)
        for (int r = lane; r < NX; r += 32) {
            vec[r] = p.x0[(size_t)b * NX + r];
            vec[NX + r] = p.lam0 ? p.lam0[(size_t)b * NX + r] : 0.0;
        }
    __syncthreads();
    if (top == 0 && warp == 1)   // a one-node upper tree: the root itself is what the lower levels read
        for (int r = lane; r < NX; r += 32) {
            p.x[0][(size_t)b * NX + r] = vec[r];
            p.lam[0][(size_t)b * NX + r] = vec[NX + r];
        }
    int rec_off = 0;
#pragma unroll 1
    for (int l = top - 1; l >= 0; --l) {
        const int cnt = p.count[l], groups = (cnt + 1) / 2, pairs = cnt / 2;
        const int pcnt = p.count[l + 1];
        const double* xp = vec + vec_off(l + 1);
        const double* lp = xp + pcnt * NX;
        double* xo = vec + vec_off(l);
        double* lo = xo + cnt * NX;
        if (pairs > 0) mbar_wait(&bar[l], 0);
        if (warp < groups) {
            const int g = warp, first = 2 * g, last = min(2 * g + 1, cnt - 1);
            for (int r = lane; r < NX; r += 32) {
                xo[first * NX + r] = xp[g * NX + r];
                lo[last * NX + r] = lp[g * NX + r];
            }
            __syncwarp();
            if (last > first)
                warp_down_step_smem<NX>(lane, recs + (size_t)(rec_off + g) * D::DREC, xp + g * NX, lp + g * NX, scratch,
                                        lo + first * NX, xo + last * NX);
            if (l == 0) {   // only the widest level is read by later launches
                double* gx = p.x[0] + (size_t)b * cnt * NX;
                double* gl = p.lam[0] + (size_t)b * cnt * NX;
                for (int r = lane; r < NX; r += 32) {
                    gx[first * NX + r] = xo[first * NX + r];
                    gl[last * NX + r] = lo[last * NX + r];
                    if (last > first) {
                        gx[last * NX + r] = xo[last * NX + r];
                        gl[first * NX + r] = lo[first * NX + r];
                    }
                }
            }
        }
        rec_off += pairs;
        __syncthreads();
    }
}

// lower level down-sweep, one warp per group: the group's records arrive by TMA while the parent vectors are read
template <int NX>
__global__ void __launch_bounds__(32) tree_down_lat_kernel(TreeParams p) {
    using D = TreeDims<NX>;
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x;
    const int b = blockIdx.x / p.groups, g = blockIdx.x % p.groups;
    const int first = g * p.R, last = min(first + p.R, p.count) - 1;
    const int nrec = last - first;
    double* recs = smem;                            // (R-1) records
    double* x = smem + (size_t)(p.R - 1) * D::DREC; // NX
    double* le = x + NX;
    double* pt = x + 2 * NX;
    double* xn = x + 3 * NX;
    const double* dd_b = p.dd + (size_t)b * p.count * D::DREC;
    double* xo = p.x_node + (size_t)b * p.count * NX;
    double* lo = p.lam_node + (size_t)b * p.count * NX;
    if (lane == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        if (nrec > 0) {
            mbar_expect_tx(&bar, (uint32_t)(nrec * D::DREC * 8));
            bulk_g2s(recs, dd_b + (size_t)first * D::DREC, (uint32_t)(nrec * D::DREC * 8), &bar);   // contiguous members
        }
    }
    for (int r = lane; r < NX; r += 32) {
        x[r] = p.x_parent[((size_t)b * p.groups + g) * NX + r];
        le[r] = p.lam_parent ? p.lam_parent[((size_t)b * p.groups + g) * NX + r] : 0.0;
        xo[(size_t)first * NX + r] = x[r];
        lo[(size_t)last * NX + r] = le[r];
    }
    __syncwarp();
    if (nrec > 0) mbar_wait(&bar, 0);
    double* xc = x;
    double* xs = xn;
#pragma unroll 1
    for (int i = first; i < last; ++i) {
        warp_down_step_smem<NX>(lane, recs + (size_t)(i - first) * D::DREC, xc, le, pt, lo + (size_t)i * NX, xs);
        for (int r = lane; r < NX; r += 32) xo[(size_t)(i + 1) * NX + r] = xs[r];
        double* sw = xc; xc = xs; xs = sw;
    }
}

}  // namespace pdplqr
