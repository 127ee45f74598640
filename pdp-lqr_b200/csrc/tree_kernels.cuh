// Condensed interface-state system, solved on the device by a tree of associative segment-summary combines
// instead of the reference's serial block recursion.
//
// Replaces  CondensedSystemLUSolver / CondensedSystemCholeskySolver  (/root/reference
// include/clqr/lqr/condensed_system.hpp:82-138 and :203-290) and the serial calls at
// lqr_solver_parallel.hpp:145,215.  Outputs are the same quantities: xhat_i (entry state of segment i) and
// uhat_i (costate at the exit of segment i), condensed_system.hpp:140-146.
//
// A segment (or any contiguous run of segments) is a "two-port"
//        x_out = F x_in + f - C lam_out ,     lam_in = P x_in + p + F^T lam_out            
// and two adjacent two-ports a (first) and b (second) compose associatively: with
// W = (I + C_a P_b)^-1,  X_F = W F_a,  X_C = W C_a,  w_f = W f_a,  x_f = w_f - X_C p_b :
//        P = P_a + F_a^T P_b X_F      p = p_a + F_a^T (P_b x_f + p_b)
//        F = F_b X_F                  f = F_b x_f + f_b               C = C_b + F_b X_C F_b^T
// (the true-terminal segment has F = 0, f = 0, C = 0, i.e. it is a pure value function).
//
// Up-sweep: groups of R consecutive nodes of a level are folded right-to-left, keeping for every member i the
// data the down-sweep needs: X_F, X_C, w_f of the fold step and (P_b, F_b, p_b) of the suffix b = (i+1 .. end of
// group).  The group total becomes a node of the next level.
// Down-sweep: given the group's entry state x and exit costate lam_e, left-to-right
//        pt = p_b + F_b^T lam_e ;  x' = X_F x + w_f - X_C pt ;  lam' = P_b x' + pt .
// These are the THROUGHPUT kernels (many problems: one warp per combine): lower levels (many nodes) run one kernel
// launch per level, one warp per group of 4; the upper levels (<= 32 nodes) run inside ONE launch of a 16-warp CTA per
// problem (binary tree, __syncthreads between levels).  With few problems the interface solve is pure latency and
// tree_lat_kernels.cuh takes over (same records, wide element-parallel combines); the plan is made in pdplqr_create.
#pragma once
#include "common.cuh"
#include "seg_kernels.cuh"

namespace pdplqr {

template <int NX>
struct TreeDims {
    static constexpr int N2 = NX * NX;
    static constexpr int SREC = 3 * N2 + 2 * NX;  // [P | F | C | p | f]  (same as SegDims::SREC)
    static constexpr int SUM_P = 0, SUM_F = N2, SUM_C = 2 * N2, SUM_p = 3 * N2, SUM_f = 3 * N2 + NX;
    // down-sweep record per member: [X_F | X_C | P_b | F_b | w_f | p_b]
    static constexpr int DD_XF = 0, DD_XC = N2, DD_PB = 2 * N2, DD_FB = 3 * N2, DD_wf = 4 * N2, DD_pb = 4 * N2 + NX;
    static constexpr int DREC = 4 * N2 + 2 * NX;
    static constexpr int NCOL = 3 * NX + 1;       // augmented [I + C_a P_b | F_a | C_a | f_a]
    static constexpr int LDA = odd_ld(NX);
};

struct TreeParams {
    int batch;
    int count;        // nodes at this level (per problem)
    int R;            // group size
    int groups;       // ceil(count / R)
    const double* sum_in;   // [batch][count][SREC]
    double* sum_out;        // [batch][groups][SREC]   (nullptr for the top level)
    double* dd;             // [batch][count][DREC]
    // down-sweep
    const double* x_parent;    // [batch][groups][NX]  entry state of each group  (x0 for the top level)
    const double* lam_parent;  // [batch][groups][NX]  exit costate of each group (nullptr -> zeros, top level)
    double* x_node;            // [batch][count][NX]
    double* lam_node;          // [batch][count][NX]
};

constexpr int TREE_TOP_MAX_LEVELS = 8;
constexpr int TREE_TOP_MAX_NODES = 32;
struct TreeTopParams {   // the upper, binary part of the tree: levels[0] is the widest (<= 32 nodes)
    int batch, nlevels;
    int count[TREE_TOP_MAX_LEVELS];
    double* sum[TREE_TOP_MAX_LEVELS];
    double* dd[TREE_TOP_MAX_LEVELS];
    double* x[TREE_TOP_MAX_LEVELS];
    double* lam[TREE_TOP_MAX_LEVELS];
    const double* x0;      // [batch][NX]   entry state of the root
    const double* lam0;    // [batch][NX]   exit costate of the root (nullptr -> zeros: the root ends at the terminal)
    int affine_only;
    // latency-mode sub-tree launches (tree_lat_kernels.cuh): `ngroups` blocks of `width` level-0 nodes per problem
    int width, ngroups, is_root, tt_cap;
};

// Gauss-Jordan with partial (row) pivoting on an NX x NCOL augmented matrix in shared memory, one warp.
// Columns are owned by lanes (column j -> lane j % 32).  On exit columns NX.. hold A^-1 * RHS.
// The pivot search is redundant per lane (broadcast reads of column k) instead of a shuffle reduction.
template <int NX, int NCOL, int LDA>
PDPLQR_DEVINL void warp_gauss_jordan(int lane, double* Aug) {
#pragma unroll 1
    for (int k = 0; k < NX; ++k) {
        // pivot search, redundant per lane (broadcast shared loads of column k, rows k..NX-1)
        double best = -1.0;
        int piv = k;
        for (int i = k; i < NX; ++i) {
            const double v = fabs(Aug[i + k * LDA]);
            if (v > best) { best = v; piv = i; }
        }
        if (piv != k) {                       // physical swap of column k's two entries (one lane), so that the
            __syncwarp();                     // multipliers below are simply column k after the row exchange
            if (lane == 0) {
                const double t = Aug[k + k * LDA];
                Aug[k + k * LDA] = Aug[piv + k * LDA];
                Aug[piv + k * LDA] = t;
            }
            __syncwarp();
        }
        double colk[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) colk[i] = Aug[i + k * LDA];
        const double pinv = 1.0 / Aug[k + k * LDA];
        for (int j = k + 1 + lane; j < NCOL; j += 32) {
            double* cj = Aug + j * LDA;
            const double akj = cj[piv];
            if (piv != k) cj[piv] = cj[k];
            const double t = akj * pinv;
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                if (i == k) cj[i] = t;
                else cj[i] = fma(-colk[i], t, cj[i]);
            }
        }
        __syncwarp();
    }
}

// barrier among the TT threads of one combine group: a warp, or TT threads on named barrier `bar_id` (1..15)
template <int TT>
PDPLQR_DEVINL void sub_sync(int bar_id) {
    if constexpr (TT == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(TT) : "memory");
}

template <int NX>
struct CombSmem {   // per-warp workspace of one combine
    using D = TreeDims<NX>;
    static constexpr int o_a = 0;                            // member summary a (SREC)
    static constexpr int o_b = o_a + D::SREC;                // suffix summary b (SREC)
    static constexpr int o_aug = o_b + D::SREC;              // LDA x NCOL
    static constexpr int o_t1 = o_aug + D::LDA * D::NCOL;    // P_b X_F  (N2)
    static constexpr int o_t2 = o_t1 + D::N2;                // F_b X_C  (N2)
    static constexpr int o_v = o_t2 + D::N2;                 // x_f, P_b x_f + p_b   (2 NX)
    static constexpr int DOUBLES = even_up(o_v + 2 * NX);
};

// One combine step by one group of TT threads (a warp, or 4 warps on a named barrier).  a, b: summaries in shared memory (ws + o_a / o_b).  Writes the combined summary to
// `out` (shared or global) and the down-sweep record of member a to `ddi` (global).
template <int NX, int TT>
PDPLQR_DEVINL void group_combine(int lane, int bar_id, double* ws, double* out, double* ddi) {
    using D = TreeDims<NX>;
    using L = CombSmem<NX>;
    constexpr int N2 = D::N2, LDA = D::LDA;
    const double* sa = ws + L::o_a;
    const double* sb = ws + L::o_b;
    double* Aug = ws + L::o_aug;
    double* T1 = ws + L::o_t1;
    double* T2 = ws + L::o_t2;
    double* xf = ws + L::o_v;
    double* lv = xf + NX;
    const double *Pa = sa + D::SUM_P, *Fa = sa + D::SUM_F, *Ca = sa + D::SUM_C, *pa = sa + D::SUM_p, *fa = sa + D::SUM_f;
    const double *Pb = sb + D::SUM_P, *Fb = sb + D::SUM_F, *Cb = sb + D::SUM_C, *pb = sb + D::SUM_p, *fb = sb + D::SUM_f;
    constexpr Tile t1 = pick_tile(NX, NX, TT);
    // Aug = [I + C_a P_b | F_a | C_a | f_a]
    {
        auto la = [&](int r, int k) { return Ca[r + k * NX]; };
        auto lb = [&](int k, int c) { return Pb[k + c * NX]; };
        auto epi = [&](int r, int c, double v) { Aug[r + c * LDA] = v + ((r == c) ? 1.0 : 0.0); };
        gmm<NX, NX, NX, t1.tm, t1.tn, TT>(lane, la, lb, epi);
        for (int e = lane; e < N2; e += TT) {
            const int r = e % NX, c = e / NX;
            Aug[r + (NX + c) * LDA] = Fa[e];
            Aug[r + (2 * NX + c) * LDA] = Ca[e];
        }
        for (int r = lane; r < NX; r += TT) Aug[r + 3 * NX * LDA] = fa[r];
    }
    sub_sync<TT>(bar_id);
    static_assert(TT == 32, "one warp per combine (the wide combines live in tree_lat_kernels.cuh)");
    warp_gauss_jordan<NX, D::NCOL, LDA>(lane, Aug);
    const double* XF = Aug + NX * LDA;
    const double* XC = Aug + 2 * NX * LDA;
    const double* wf = Aug + 3 * NX * LDA;
    // x_f = w_f - X_C p_b ; down-sweep record of member a
    for (int r = lane; r < NX; r += TT) {
        double acc = wf[r];
#pragma unroll 4
        for (int k = 0; k < NX; ++k) acc = fma(-XC[r + k * LDA], pb[k], acc);
        xf[r] = acc;
        ddi[D::DD_wf + r] = wf[r];
        ddi[D::DD_pb + r] = pb[r];
    }
    for (int e = lane; e < N2; e += TT) {
        const int r = e % NX, c = e / NX;
        ddi[D::DD_XF + e] = XF[r + c * LDA];
        ddi[D::DD_XC + e] = XC[r + c * LDA];
        ddi[D::DD_PB + e] = Pb[e];
        ddi[D::DD_FB + e] = Fb[e];
    }
    // one pass: T1 = P_b X_F ; T2 = F_b X_C ; F = F_b X_F
    {
        constexpr Tile t3 = pick_tile(3 * NX, NX, TT);
        auto la = [&](int g, int r, int k) { return g == 0 ? Pb[r + k * NX] : Fb[r + k * NX]; };
        auto lb = [&](int g, int k, int c) { return g == 1 ? XC[k + c * LDA] : XF[k + c * LDA]; };
        auto ep = [&](int g, int r, int c, double v) {
            if (g == 0) T1[r + c * NX] = v;
            else if (g == 1) T2[r + c * NX] = v;
            else out[D::SUM_F + r + c * NX] = v;
        };
        gmm_multi<3, NX, NX, NX, t3.tm, t3.tn, TT>(lane, la, lb, ep);
    }
    sub_sync<TT>(bar_id);
    // lv = P_b x_f + p_b  (needs xf)
    for (int r = lane; r < NX; r += TT) {
        double acc = pb[r];
#pragma unroll 4
        for (int k = 0; k < NX; ++k) acc = fma(Pb[r + k * NX], xf[k], acc);
        lv[r] = acc;
    }
    // one pass: P = P_a + F_a^T T1 ; C = C_b + T2 F_b^T
    {
        constexpr Tile t2 = pick_tile(2 * NX, NX, TT);
        auto la = [&](int g, int r, int k) { return g == 0 ? Fa[k + r * NX] : T2[r + k * NX]; };
        auto lb = [&](int g, int k, int c) { return g == 0 ? T1[k + c * NX] : Fb[c + k * NX]; };
        auto ep = [&](int g, int r, int c, double v) {
            if (g == 0) out[D::SUM_P + r + c * NX] = Pa[r + c * NX] + v;
            else out[D::SUM_C + r + c * NX] = Cb[r + c * NX] + v;
        };
        gmm_multi<2, NX, NX, NX, t2.tm, t2.tn, TT>(lane, la, lb, ep);
    }
    sub_sync<TT>(bar_id);
    // p = p_a + F_a^T lv ; f = F_b x_f + f_b
    for (int r = lane; r < NX; r += TT) {
        double ap = pa[r], af = fb[r];
#pragma unroll 4
        for (int k = 0; k < NX; ++k) {
            ap = fma(Fa[k + r * NX], lv[k], ap);
            af = fma(Fb[r + k * NX], xf[k], af);
        }
        out[D::SUM_p + r] = ap;
        out[D::SUM_f + r] = af;
    }
    sub_sync<TT>(bar_id);
}

// affine-only version of one combine step: members' p, f changed, P, F, C (hence X_F, X_C, P_b, F_b) did not.
// With W = I - X_C P_b:  x_f = f_a - X_C (P_b f_a + p_b) ; w_f = x_f + X_C p_b ; p = p_a + F_a^T (P_b x_f + p_b) ;
// f = F_b x_f + f_b.   sa: summary of a (global), pb/fb: suffix vectors (shared, NX each, updated in place),
// tmp: 3*NX doubles of shared scratch.  NX <= 32.
template <int NX>
PDPLQR_DEVINL void warp_combine_affine(int lane, const double* sa, double* ddi, double* pb, double* fb, double* tmp) {
    using D = TreeDims<NX>;
    double* t0 = tmp;
    double* xf = tmp + NX;
    double* lv = tmp + 2 * NX;
    if (lane < NX) {
        double acc = pb[lane];
#pragma unroll 4
        for (int k = 0; k < NX; ++k) acc = fma(ddi[D::DD_PB + lane + k * NX], sa[D::SUM_f + k], acc);
        t0[lane] = acc;
    }
    __syncwarp();
    if (lane < NX) {
        double acc = sa[D::SUM_f + lane], accw = 0.0;
#pragma unroll 4
        for (int k = 0; k < NX; ++k) {
            acc = fma(-ddi[D::DD_XC + lane + k * NX], t0[k], acc);
            accw = fma(ddi[D::DD_XC + lane + k * NX], pb[k], accw);
        }
        xf[lane] = acc;
        ddi[D::DD_wf + lane] = acc + accw;
        ddi[D::DD_pb + lane] = pb[lane];
    }
    __syncwarp();
    if (lane < NX) {
        double acc = pb[lane];
#pragma unroll 4
        for (int k = 0; k < NX; ++k) acc = fma(ddi[D::DD_PB + lane + k * NX], xf[k], acc);
        lv[lane] = acc;
    }
    __syncwarp();
    double pnew = 0.0, fnew = 0.0;
    if (lane < NX) {
        double ap = sa[D::SUM_p + lane], af = fb[lane];
#pragma unroll 4
        for (int k = 0; k < NX; ++k) {
            ap = fma(sa[D::SUM_F + k + lane * NX], lv[k], ap);
            af = fma(ddi[D::DD_FB + lane + k * NX], xf[k], af);
        }
        pnew = ap; fnew = af;
    }
    __syncwarp();
    if (lane < NX) { pb[lane] = pnew; fb[lane] = fnew; }
    __syncwarp();
}

// one down-sweep step of member i (shared x, le, pt, xn: NX each)
template <int NX>
PDPLQR_DEVINL void warp_down_step(int lane, const double* ddi, double* x, const double* le, double* pt, double* xn,
                                  double* lam_i, double* x_next) {
    using D = TreeDims<NX>;
    for (int r = lane; r < NX; r += 32) {
        double acc = ddi[D::DD_pb + r];
#pragma unroll 4
        for (int k = 0; k < NX; ++k) acc = fma(ddi[D::DD_FB + k + r * NX], le[k], acc);   // F_b^T lam_e
        pt[r] = acc;
    }
    __syncwarp();
    for (int r = lane; r < NX; r += 32) {
        double acc = ddi[D::DD_wf + r];
#pragma unroll 4
        for (int k = 0; k < NX; ++k) {
            acc = fma(ddi[D::DD_XF + r + k * NX], x[k], acc);
            acc = fma(-ddi[D::DD_XC + r + k * NX], pt[k], acc);
        }
        xn[r] = acc;
    }
    __syncwarp();
    for (int r = lane; r < NX; r += 32) {
        double acc = pt[r];
#pragma unroll 4
        for (int k = 0; k < NX; ++k) acc = fma(ddi[D::DD_PB + r + k * NX], xn[k], acc);
        lam_i[r] = acc;
        x_next[r] = xn[r];
        x[r] = xn[r];
    }
    __syncwarp();
}

// ---------------------------------------------------------------- one launch per level, one warp per group
template <int NX>
struct TreeSmem {
    static constexpr int o_n = CombSmem<NX>::DOUBLES;        // new suffix (SREC)
    static constexpr size_t BYTES = (size_t)(o_n + TreeDims<NX>::SREC) * 8;
};

template <int NX, int TT>
__global__ void __launch_bounds__(TT) tree_up_kernel(TreeParams p) {
    using D = TreeDims<NX>;
    using L = CombSmem<NX>;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x / p.groups, g = blockIdx.x % p.groups;
    const int first = g * p.R, last = min(first + p.R, p.count) - 1;
    double* sa = smem + L::o_a;
    double* sb = smem + L::o_b;
    double* sn = smem + TreeSmem<NX>::o_n;
    const double* in_b = p.sum_in + (size_t)b * p.count * D::SREC;
    double* dd_b = p.dd + (size_t)b * p.count * D::DREC;
    for (int e = lane; e < D::SREC; e += TT) sb[e] = in_b[(size_t)last * D::SREC + e];
    sub_sync<TT>(1);
#pragma unroll 1
    for (int i = last - 1; i >= first; --i) {
        for (int e = lane; e < D::SREC; e += TT) sa[e] = in_b[(size_t)i * D::SREC + e];
        sub_sync<TT>(1);
        group_combine<NX, TT>(lane, 1, smem, sn, dd_b + (size_t)i * D::DREC);
        for (int e = lane; e < D::SREC; e += TT) sb[e] = sn[e];
        sub_sync<TT>(1);
    }
    if (p.sum_out) {
        double* out = p.sum_out + ((size_t)b * p.groups + g) * D::SREC;
        for (int e = lane; e < D::SREC; e += TT) out[e] = sb[e];
    }
}

template <int NX>
__global__ void __launch_bounds__(32) tree_up_affine_kernel(TreeParams p) {
    using D = TreeDims<NX>;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x / p.groups, g = blockIdx.x % p.groups;
    const int first = g * p.R, last = min(first + p.R, p.count) - 1;
    double* pb = smem;
    double* fb = smem + NX;
    const double* in_b = p.sum_in + (size_t)b * p.count * D::SREC;
    double* dd_b = p.dd + (size_t)b * p.count * D::DREC;
    for (int r = lane; r < NX; r += 32) {
        pb[r] = in_b[(size_t)last * D::SREC + D::SUM_p + r];
        fb[r] = in_b[(size_t)last * D::SREC + D::SUM_f + r];
    }
    __syncwarp();
#pragma unroll 1
    for (int i = last - 1; i >= first; --i)
        warp_combine_affine<NX>(lane, in_b + (size_t)i * D::SREC, dd_b + (size_t)i * D::DREC, pb, fb, smem + 2 * NX);
    if (p.sum_out) {
        double* out = p.sum_out + ((size_t)b * p.groups + g) * D::SREC;
        for (int r = lane; r < NX; r += 32) {
            out[D::SUM_p + r] = pb[r];
            out[D::SUM_f + r] = fb[r];
        }
    }
}

template <int NX>
__global__ void __launch_bounds__(32) tree_down_kernel(TreeParams p) {
    using D = TreeDims<NX>;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x / p.groups, g = blockIdx.x % p.groups;
    const int first = g * p.R, last = min(first + p.R, p.count) - 1;
    double* x = smem;            // NX
    double* le = smem + NX;      // NX  exit costate of the group
    double* pt = smem + 2 * NX;  // NX
    double* xn = smem + 3 * NX;  // NX
    const double* dd_b = p.dd + (size_t)b * p.count * D::DREC;
    double* xo = p.x_node + (size_t)b * p.count * NX;
    double* lo = p.lam_node + (size_t)b * p.count * NX;
    for (int r = lane; r < NX; r += 32) {
        x[r] = p.x_parent[((size_t)b * p.groups + g) * NX + r];
        le[r] = p.lam_parent ? p.lam_parent[((size_t)b * p.groups + g) * NX + r] : 0.0;
        xo[(size_t)first * NX + r] = x[r];
        lo[(size_t)last * NX + r] = le[r];
    }
    __syncwarp();
#pragma unroll 1
    for (int i = first; i < last; ++i)
        warp_down_step<NX>(lane, dd_b + (size_t)i * D::DREC, x, le, pt, xn, lo + (size_t)i * NX, xo + (size_t)(i + 1) * NX);
}

// ---------------------------------------------------------------- upper levels: one CTA per problem, one launch
template <int NX>
struct TreeTopSmem {
    static constexpr int WARP_DOUBLES = CombSmem<NX>::DOUBLES;
    // as many warps as fit in ~220 KB of shared memory (16 at nx = 12)
    static constexpr int WARPS = (220 * 1024 / 8 / WARP_DOUBLES) >= 16 ? 16 : ((220 * 1024 / 8 / WARP_DOUBLES) < 1 ? 1 : (220 * 1024 / 8 / WARP_DOUBLES));
    static constexpr size_t BYTES = (size_t)WARPS * WARP_DOUBLES * 8;
};

template <int NX, int TT>
__global__ void __launch_bounds__(TreeTopSmem<NX>::WARPS * 32) tree_top_up_kernel(TreeTopParams p) {
    using D = TreeDims<NX>;
    using L = CombSmem<NX>;
    constexpr int NGROUPS = TreeTopSmem<NX>::WARPS * 32 / TT;    // combine groups per CTA
    static_assert(NGROUPS >= 1 && (TT == 32 || NGROUPS <= 15), "one named barrier per group");
    extern __shared__ __align__(16) double smem[];
    const int grp = threadIdx.x / TT, lane = threadIdx.x % TT;
    const int bar_id = 1 + grp;
    const int b = blockIdx.x;
    double* ws = smem + grp * TreeTopSmem<NX>::WARP_DOUBLES;   // each group uses one warp-slot of the layout
#pragma unroll 1
    for (int l = 0; l + 1 < p.nlevels; ++l) {   // level l (count[l] nodes) -> level l+1 (pairs)
        const int cnt = p.count[l], groups = (cnt + 1) / 2;
        const double* in_b = p.sum[l] + (size_t)b * cnt * D::SREC;
        double* out_b = p.sum[l + 1] + (size_t)b * groups * D::SREC;
        double* dd_b = p.dd[l] + (size_t)b * cnt * D::DREC;
        for (int g = grp; g < groups; g += NGROUPS) {
            const int ia = 2 * g, ib = 2 * g + 1;
            double* out = out_b + (size_t)g * D::SREC;
            if (ib >= cnt) {                       // odd node out: passes through unchanged
                if (p.affine_only) {
                    for (int r = lane; r < NX; r += TT) {
                        out[D::SUM_p + r] = in_b[(size_t)ia * D::SREC + D::SUM_p + r];
                        out[D::SUM_f + r] = in_b[(size_t)ia * D::SREC + D::SUM_f + r];
                    }
                } else
                    for (int e = lane; e < D::SREC; e += TT) out[e] = in_b[(size_t)ia * D::SREC + e];
                continue;
            }
            if (p.affine_only) {
                if (lane < 32) {   // matvec-only: the first warp of the group
                    double* pb = ws;
                    double* fb = ws + NX;
                    for (int r = lane; r < NX; r += 32) {
                        pb[r] = in_b[(size_t)ib * D::SREC + D::SUM_p + r];
                        fb[r] = in_b[(size_t)ib * D::SREC + D::SUM_f + r];
                    }
                    __syncwarp();
                    warp_combine_affine<NX>(lane, in_b + (size_t)ia * D::SREC, dd_b + (size_t)ia * D::DREC, pb, fb, ws + 2 * NX);
                    for (int r = lane; r < NX; r += 32) {
                        out[D::SUM_p + r] = pb[r];
                        out[D::SUM_f + r] = fb[r];
                    }
                }
            } else {
                sub_sync<TT>(bar_id);   // previous round's readers of ws are done
                for (int e = lane; e < D::SREC; e += TT) {
                    ws[L::o_a + e] = in_b[(size_t)ia * D::SREC + e];
                    ws[L::o_b + e] = in_b[(size_t)ib * D::SREC + e];
                }
                sub_sync<TT>(bar_id);
                group_combine<NX, TT>(lane, bar_id, ws, out, dd_b + (size_t)ia * D::DREC);
            }
        }
        __threadfence_block();
        __syncthreads();
    }
}

template <int NX>
__global__ void __launch_bounds__(TreeTopSmem<NX>::WARPS * 32) tree_top_down_kernel(TreeTopParams p) {
    using D = TreeDims<NX>;
    constexpr int TREE_TOP_WARPS = TreeTopSmem<NX>::WARPS;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x;
    double* x = smem + warp * 4 * NX;
    double* le = x + NX;
    double* pt = x + 2 * NX;
    double* xn = x + 3 * NX;
    // root: the single node of the last level gets x0 and a zero exit costate
    {
        const int top = p.nlevels - 1;
        if (warp == 0)
            for (int r = lane; r < NX; r += 32) {
                p.x[top][(size_t)b * NX + r] = p.x0[(size_t)b * NX + r];
                p.lam[top][(size_t)b * NX + r] = p.lam0 ? p.lam0[(size_t)b * NX + r] : 0.0;
            }
        __threadfence_block();
        __syncthreads();
    }
#pragma unroll 1
    for (int l = p.nlevels - 2; l >= 0; --l) {
        const int cnt = p.count[l], groups = (cnt + 1) / 2;
        const double* dd_b = p.dd[l] + (size_t)b * cnt * D::DREC;
        const double* xp = p.x[l + 1] + (size_t)b * groups * NX;
        const double* lp = p.lam[l + 1] + (size_t)b * groups * NX;
        double* xo = p.x[l] + (size_t)b * cnt * NX;
        double* lo = p.lam[l] + (size_t)b * cnt * NX;
        for (int g = warp; g < groups; g += TREE_TOP_WARPS) {
            const int first = 2 * g, last = min(2 * g + 1, cnt - 1);
            for (int r = lane; r < NX; r += 32) {
                x[r] = xp[(size_t)g * NX + r];
                le[r] = lp[(size_t)g * NX + r];
                xo[(size_t)first * NX + r] = x[r];
                lo[(size_t)last * NX + r] = le[r];
            }
            __syncwarp();
            if (last > first)
                warp_down_step<NX>(lane, dd_b + (size_t)first * D::DREC, x, le, pt, xn, lo + (size_t)first * NX,
                                   xo + (size_t)last * NX);
        }
        __threadfence_block();
        __syncthreads();
    }
}

}  // namespace pdplqr
