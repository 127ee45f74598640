// Latency-mode versions of the interface-tree kernels (tree_kernels.cuh), used when there are few problems (one
// long-horizon problem, horizon shards, the shard coupler): the interface solve is then a chain of dependent small
// dense steps and what matters is the length of that chain, not throughput.
//
// Same mathematics and the same summary / down-sweep records as tree_kernels.cuh (reference: the serial block recursion
// of /root/reference include/clqr/lqr/condensed_system.hpp:82-146), different mapping:
//   * the tree is binary; one CTA reduces a block of 8 consecutive nodes through all its levels inside one launch
//     (4 x 256, 2 x 256, 1 x 256 threads), handing each result to its consumer through shared memory; the next launch
//     does the same one storey up (S = 128: two launches of three and four levels);
//   * a combine is spread element-parallel over tt = 64 .. 256 threads; its five 12x12x12-class products run on the
//     FP64 tensor cores (DMMA), one 8x8 output tile per warp at a time;
//   * Gauss-Jordan: every thread owns a fixed 4-row chunk of one column and keeps it in registers over all pivot steps;
//     rows are pivoted implicitly (no exchange, one un-permuting store at the end); the pivot of step k+1 is found by
//     the owners of column k+1 with a shared-memory atomicMax on (magnitude bits | row) keys while they write step k;
//     the two buffers ping-pong (one barrier per step); the reciprocal is a Newton-refined hardware seed;
//   * the down-sweep records are pulled into shared memory by TMA bulk copies before the level loop starts, so the
//     dependent mat-vec chain never waits for L2.
// Round 2 tried a register-resident alternative for the elimination (one warp, one column of [M | I] per lane, column k
// fetched by shuffles, redundant per-lane pivot search, W = M^-1 followed by tensor-core products W F_a, W C_a): its pivot
// step measured 625 cycles against the 610 of the scheme below (the serial argmax / select chains of a single in-order warp
// cost what the barrier and the shared atomics cost here), and with four of them on one SM sub-partition 2.3x more -- not
// kept (git history: "combine_fast").
// A second round-2 experiment, "gj_blocked": two pivot steps per barrier, blocked like an LU panel factorisation -- a group of
// 3 lanes owns a pair of columns in registers (rows in 4-row chunks), the warp holding the panel finds both pivots with
// shuffles, publishes the two multiplier columns + reciprocals once, and every group applies the rank-2 update with the
// pivot-row entries shuffled inside the group (57 threads instead of 256, one barrier per two steps).  Parity-identical, and
// again no faster: per pair  step A 455 + step B / publish 424 + barrier 14 + rank-2 update 366 = 1,259 cycles = 630 per pivot
// step (clock64, C2), C2 94.3 against 94.1 us.  (A first version whose run-time register selects `x[piv & 3]` the compiler
// turned into branches with reconvergence barriers measured 950 per step; selp.f64 in inline PTX fixed that.)  Three
// formulations, 610 - 630 cycles per step each: the step is a chain of ~100 dependent instructions -- key, reduce, select,
// fetch, reciprocal, scale, fetch, update -- on an in-order warp, and none of them shortens that chain.
// Measured latencies that shape this (scripts/micro/lat_bench.cu, B200): LDS 30, DFMA 8, bar.sync (128 thr) 20,
// STS->bar->LDS 55, ATOMS.MAX+LDS 51, reciprocal 48 (IEEE division 72), 64-bit shuffle 28 cycles.
#pragma once
#include "tree_kernels.cuh"

namespace pdplqr {

PDPLQR_DEVINL void rt_sync(int tt, int bar_id) {
    if (tt == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(tt) : "memory");
}

constexpr int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

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
    // one slot per concurrent combine = [a | b | a' | b' | work]: the inputs of even and odd levels alternate between
    // the two pairs, so that a combine can drop its result straight into the slot of the combine that consumes it
    static constexpr int o_WK = even_up(4 * D::SREC);   // the workspace is accessed with 16-byte vectors
    static constexpr int SLOT = o_WK + even_up(WORK);
    static constexpr int TOP_THREADS = NX <= 12 ? 1024 : 512;
    static constexpr int TT_CAP = 256;                   // threads per combine at most (measured: 256 beats 128 and 64
                                                         // at nx = 12 although only ITEMS = 108 threads own elements)
    static constexpr int fit = (200 * 1024 / 8) / SLOT;
    static constexpr int by_threads = TOP_THREADS / MINTT;
    static constexpr int lim = fit < by_threads ? fit : by_threads;
    // measured at nx = 12 (C2): blocks of 8 nodes (4 concurrent combines per SM) beat 16 and 4
    static constexpr int TOP_SLOTS = lim >= 4 ? 4 : (lim >= 2 ? 2 : (lim >= 1 ? 1 : 0));
    static constexpr int TOP_NODES = 2 * TOP_SLOTS;      // widest block of nodes one CTA reduces
    static constexpr size_t TOP_BYTES = (size_t)(TOP_SLOTS > 0 ? TOP_SLOTS : 1) * SLOT * 8;
    // down-sweeps: the records of every pair of the block's sub-tree (<= TOP_NODES - 1) + per-level vectors
    static constexpr int DOWN_MAXREC = TOP_NODES > 1 ? TOP_NODES - 1 : 1;
    static constexpr int DOWN_VEC = 2 * NX * (2 * (TOP_NODES > 0 ? TOP_NODES : 1)) + 4 * NX * (TOP_THREADS / 32);
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
#ifdef PDPLQR_PHASE_CLOCKS
#define TPH_DECL long long tph[8]; const bool tph_on = (t == 0 && bar_id == 1 && blockIdx.x == 0); int tph_n = 0; if (tph_on) tph[tph_n++] = clock64();
#define TPH() do { if (tph_on) tph[tph_n++] = clock64(); } while (0)
#define TPH_PRINT() do { if (tph_on) printf("combine_lat tt %d | form %lld gj %lld unperm %lld prod3 %lld prod4 %lld p %lld | total %lld\n", tt, tph[1]-tph[0], tph[2]-tph[1], tph[3]-tph[2], tph[4]-tph[3], tph[5]-tph[4], tph[6]-tph[5], tph[6]-tph[0]); } while (0)
#else
#define TPH_DECL
#define TPH() do {} while (0)
#define TPH_PRINT() do {} while (0)
#endif
template <int NX>
PDPLQR_DEVINL void combine_lat(int t, int tt, int bar_id, const double* sa, const double* sb, double* wk, double* out,
                               double* ddi) {
    using D = TreeDims<NX>;
    using L = LatSmem<NX>;
    TPH_DECL
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
    TPH();
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
    TPH();
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
    TPH();
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
    TPH();
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
    TPH();
    // ---- phase 5: p = p_a + F_a^T lv   (then: `out` may be picked up by a bulk copy -> async-proxy fence)
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
    fence_proxy_async();
    rt_sync(tt, bar_id);
    TPH();
    TPH_PRINT();
}

// ---------------------------------------------------------------- binary sub-trees: one CTA per (problem, block of
// `width` consecutive nodes of the launch's lowest level), all levels of the block inside one launch.  The launch of
// the uppermost levels has one block per problem (is_root).  Level l of the launch holds count[l] nodes per problem,
// count[l+1] = ceil(count[l] / 2); block g owns nodes [g (width >> l), (g+1) (width >> l)) of level l.
template <int NX>
__global__ void __launch_bounds__(LatSmem<NX>::TOP_THREADS) tree_sub_up_lat_kernel(TreeTopParams p) {
    using D = TreeDims<NX>;
    using L = LatSmem<NX>;
    constexpr int THREADS = L::TOP_THREADS;
    constexpr bool BULK_OUT = (D::SREC % 2 == 0);   // 16-byte granularity of cp.async.bulk
    extern __shared__ __align__(16) double smem[];
    pdl_wait();      // the predecessor kernel of the solve chain has completed (no-op without the launch attribute)
    pdl_trigger();   // the next kernel of the chain may be scheduled from here on
    const int tid = threadIdx.x;
    const int b = blockIdx.x / p.ngroups, g0 = blockIdx.x % p.ngroups;
#pragma unroll 1
    for (int l = 0; l + 1 < p.nlevels; ++l) {   // level l -> level l+1 (pairs)
        const int cnt = p.count[l], cnt_up = p.count[l + 1];
        const int w = p.width >> l, n0 = g0 * w;
        const int mine = min(w, cnt - n0);      // nodes of this block at this level (>= 1)
        const int groups = (mine + 1) / 2;
        int ng = 1;
        while (ng < groups) ng <<= 1;           // <= TOP_SLOTS (width <= TOP_NODES)
        int tt = THREADS / ng;
        if (tt > p.tt_cap) tt = p.tt_cap;
        const int grp = tid / tt, t = tid % tt;
        bool issued = false;                    // this thread committed a bulk store at this level
        if (grp < groups) {
            const double* in_b = p.sum[l] + ((size_t)b * cnt + n0) * D::SREC;
            double* out = p.sum[l + 1] + ((size_t)b * cnt_up + (n0 >> 1) + grp) * D::SREC;
            double* dd_a = p.dd[l] + ((size_t)b * cnt + n0 + 2 * grp) * D::DREC;
            const int ia = 2 * grp, ib = 2 * grp + 1;
            double* ws = smem + (size_t)grp * L::SLOT;
            if (p.affine_only) {                // members' p, f changed only: mat-vec work, through global memory
                if (ib >= mine) {
                    for (int r = t; r < NX; r += tt) {
                        out[D::SUM_p + r] = in_b[(size_t)ia * D::SREC + D::SUM_p + r];
                        out[D::SUM_f + r] = in_b[(size_t)ia * D::SREC + D::SUM_f + r];
                    }
                } else if (t < 32) {            // the first warp of the group (warp_combine_affine, tree_kernels.cuh)
                    double* pb = ws;
                    double* fb = ws + NX;
                    for (int r = t; r < NX; r += 32) {
                        pb[r] = in_b[(size_t)ib * D::SREC + D::SUM_p + r];
                        fb[r] = in_b[(size_t)ib * D::SREC + D::SUM_f + r];
                    }
                    __syncwarp();
                    warp_combine_affine<NX>(t, in_b + (size_t)ia * D::SREC, dd_a, pb, fb, ws + 2 * NX);
                    for (int r = t; r < NX; r += 32) {
                        out[D::SUM_p + r] = pb[r];
                        out[D::SUM_f + r] = fb[r];
                    }
                }
            } else {
                // inputs of this level: pair (l & 1) of the slot; the result goes to the other pair of the slot of the
                // combine that consumes it at level l+1 (as its a if grp is even, its b if odd) and to global memory
                double* sa = ws + (l & 1) * 2 * D::SREC;
                double* sb = sa + D::SREC;
                double* outs = smem + (size_t)(grp >> 1) * L::SLOT + ((l & 1) ^ 1) * 2 * D::SREC + (grp & 1) * D::SREC;
                if (l == 0) {                   // the block's lowest level comes from global memory
                    for (int e = t; e < D::SREC; e += tt) {
                        sa[e] = in_b[(size_t)ia * D::SREC + e];
                        if (ib < mine) sb[e] = in_b[(size_t)ib * D::SREC + e];
                    }
                }
                combine_lat_init<NX>(t, tt, ws + L::o_WK);   // slots change owners between levels
                rt_sync(tt, 1 + grp);
                if (ib >= mine) {               // odd node out: passes through unchanged
                    for (int e = t; e < D::SREC; e += tt) outs[e] = sa[e];
                    fence_proxy_async();
                    rt_sync(tt, 1 + grp);
                } else
                    combine_lat<NX>(t, tt, 1 + grp, sa, sb, ws + L::o_WK, outs, dd_a);
                if constexpr (BULK_OUT) {
                    if (t == 0) {
                        bulk_s2g(out, outs, D::SREC * 8);
                        bulk_commit();
                        issued = true;
                    }
                } else
                    for (int e = t; e < D::SREC; e += tt) out[e] = outs[e];
            }
        }
        // `outs` is an input of the next level and is overwritten two levels on: what has to be over before the barrier below
        // is the read of the PREVIOUS level's store, not of the one just issued (waiting for that one held the whole CTA for
        // the drain time of the copy at every level of the chain)
        if constexpr (BULK_OUT) {
            if (issued) bulk_wait_read<1>();
            else bulk_wait_read<0>();
        }
        __threadfence_block();
        __syncthreads();
    }
    if constexpr (BULK_OUT) bulk_wait_read<0>();   // shared memory must outlive the last store's read
}

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
__global__ void __launch_bounds__(LatSmem<NX>::TOP_THREADS) tree_sub_down_lat_kernel(TreeTopParams p) {
    using D = TreeDims<NX>;
    using L = LatSmem<NX>;
    extern __shared__ __align__(16) double smem[];
    pdl_wait();      // the predecessor kernel of the solve chain has completed (no-op without the launch attribute)
    pdl_trigger();   // the next kernel of the chain may be scheduled from here on
    __shared__ __align__(8) uint64_t bar[TREE_TOP_MAX_LEVELS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / p.ngroups, g0 = blockIdx.x % p.ngroups;
    double* recs = smem;                                   // pair records, upper levels first
    double* vec = smem + L::DOWN_MAXREC * D::DREC;         // per level: x[width >> l][NX] then lam[width >> l][NX]
    double* scratch = vec + 2 * NX * 2 * L::TOP_NODES + warp * 4 * NX;
    const int top = p.nlevels - 1;
    auto mine_at = [&](int l) { return min(p.width >> l, p.count[l] - g0 * (p.width >> l)); };
    if (tid == 0) {
        for (int l = 0; l < top; ++l) mbar_init(&bar[l], 1);
        mbar_fence_init();
        int off = 0;
        for (int l = top - 1; l >= 0; --l) {               // the order the sweep needs them in
            const int cnt = p.count[l], n0 = g0 * (p.width >> l), pairs = mine_at(l) / 2;
            if (pairs > 0) mbar_expect_tx(&bar[l], (uint32_t)(pairs * D::DREC * 8));
            for (int g = 0; g < pairs; ++g) {
                bulk_g2s(recs + (size_t)off * D::DREC, p.dd[l] + ((size_t)b * cnt + n0 + 2 * g) * D::DREC, D::DREC * 8, &bar[l]);
                ++off;
            }
        }
    }
    auto vec_off = [&](int l) {                            // level l starts after all upper levels
        int o = 0;
        for (int m = top; m > l; --m) o += p.width >> m;
        return o * 2 * NX;
    };
    if (warp == 1) {   // the block's node of the launch's top level: the root boundary, or what the launch above left
        const size_t at = ((size_t)b * p.count[top] + g0) * NX;
        for (int r = lane; r < NX; r += 32) {
            vec[r] = p.is_root ? p.x0[(size_t)b * NX + r] : p.x[top][at + r];
            vec[NX + r] = p.is_root ? (p.lam0 ? p.lam0[(size_t)b * NX + r] : 0.0) : p.lam[top][at + r];
        }
    }
    __syncthreads();
    if (top == 0 && warp == 1 && p.is_root)   // a one-node upper tree: the root itself is what the lower levels read
        for (int r = lane; r < NX; r += 32) {
            p.x[0][(size_t)b * NX + r] = vec[r];
            p.lam[0][(size_t)b * NX + r] = vec[NX + r];
        }
    int rec_off = 0;
#pragma unroll 1
    for (int l = top - 1; l >= 0; --l) {
        const int cnt = p.count[l], w = p.width >> l, n0 = g0 * w;
        const int mine = mine_at(l), groups = (mine + 1) / 2, pairs = mine / 2;
        const int wup = p.width >> (l + 1);
        const double* xp = vec + vec_off(l + 1);
        const double* lp = xp + wup * NX;
        double* xo = vec + vec_off(l);
        double* lo = xo + w * NX;
        if (pairs > 0) mbar_wait(&bar[l], 0);
        if (warp < groups) {
            const int g = warp, first = 2 * g, last = min(2 * g + 1, mine - 1);
            for (int r = lane; r < NX; r += 32) {
                xo[first * NX + r] = xp[g * NX + r];
                lo[last * NX + r] = lp[g * NX + r];
            }
            __syncwarp();
            if (last > first)
                warp_down_step_smem<NX>(lane, recs + (size_t)(rec_off + g) * D::DREC, xp + g * NX, lp + g * NX, scratch,
                                        lo + first * NX, xo + last * NX);
            if (l == 0) {   // only the launch's lowest level is read by later launches
                double* gx = p.x[0] + ((size_t)b * cnt + n0) * NX;
                double* gl = p.lam[0] + ((size_t)b * cnt + n0) * NX;
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

}  // namespace pdplqr
