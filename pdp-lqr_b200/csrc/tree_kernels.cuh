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
// Up-sweep (one warp per group of R consecutive nodes of a level): fold the group right-to-left, keeping for
// every member i the data the down-sweep needs: X_F, X_C, w_f of the fold step and (P_b, F_b, p_b) of the
// suffix b = (i+1 .. end of group).  The group total becomes a node of the next level.
// Down-sweep: given the group's entry state x and exit costate lam_e, left-to-right
//        pt = p_b + F_b^T lam_e ;  x' = X_F x + w_f - X_C pt ;  lam' = P_b x' + pt .
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

// Gauss-Jordan with partial (row) pivoting on an NX x NCOL augmented matrix in shared memory, one warp.
// Columns are owned by lanes (column j -> lane j % 32).  On exit columns NX.. hold A^-1 * RHS.
template <int NX, int NCOL, int LDA>
PDPLQR_DEVINL void warp_gauss_jordan(int lane, double* Aug) {
#pragma unroll 1
    for (int k = 0; k < NX; ++k) {
        // pivot search in column k, rows k..NX-1
        double best = -1.0;
        int piv = k;
        for (int i = k + lane; i < NX; i += 32) {
            const double v = fabs(Aug[i + k * LDA]);
            if (v > best) { best = v; piv = i; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int op = __shfl_xor_sync(0xffffffffu, piv, off);
            if (ob > best || (ob == best && op < piv)) { best = ob; piv = op; }
        }
        // row swap (each lane in its own columns)
        if (piv != k) {
            for (int j = k + lane; j < NCOL; j += 32) {
                const double t = Aug[k + j * LDA];
                Aug[k + j * LDA] = Aug[piv + j * LDA];
                Aug[piv + j * LDA] = t;
            }
        }
        __syncwarp();
        const double pinv = 1.0 / Aug[k + k * LDA];
        for (int j = k + 1 + lane; j < NCOL; j += 32) {
            const double t = Aug[k + j * LDA] * pinv;
            Aug[k + j * LDA] = t;
#pragma unroll 4
            for (int i = 0; i < NX; ++i)
                if (i != k) Aug[i + j * LDA] = fma(-Aug[i + k * LDA], t, Aug[i + j * LDA]);
        }
        __syncwarp();
    }
}

template <int NX>
struct TreeSmem {
    using D = TreeDims<NX>;
    static constexpr int o_a = 0;                     // member summary a (SREC)
    static constexpr int o_b = o_a + D::SREC;         // suffix summary b (SREC)
    static constexpr int o_n = o_b + D::SREC;         // new suffix (SREC)
    static constexpr int o_aug = o_n + D::SREC;       // LDA x NCOL
    static constexpr int o_t1 = o_aug + D::LDA * D::NCOL;   // P_b X_F  (N2)
    static constexpr int o_t2 = o_t1 + D::N2;               // F_b X_C  (N2)
    static constexpr int o_v = o_t2 + D::N2;                // x_f, P_b x_f + p_b   (2 NX)
    static constexpr int DOUBLES = o_v + 2 * NX;
    static constexpr size_t BYTES = (size_t)DOUBLES * 8;
};

template <int NX>
__global__ void __launch_bounds__(32) tree_up_kernel(TreeParams p) {
    using D = TreeDims<NX>;
    using L = TreeSmem<NX>;
    constexpr int N2 = D::N2, LDA = D::LDA;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x / p.groups, g = blockIdx.x % p.groups;
    const int first = g * p.R, last = min(first + p.R, p.count) - 1;
    double* sa = smem + L::o_a;
    double* sb = smem + L::o_b;
    double* sn = smem + L::o_n;
    double* Aug = smem + L::o_aug;
    double* T1 = smem + L::o_t1;
    double* T2 = smem + L::o_t2;
    double* xf = smem + L::o_v;
    double* lv = xf + NX;
    const double* in_b = p.sum_in + (size_t)b * p.count * D::SREC;
    double* dd_b = p.dd + (size_t)b * p.count * D::DREC;

    for (int e = lane; e < D::SREC; e += 32) sb[e] = in_b[(size_t)last * D::SREC + e];
    __syncwarp();
#pragma unroll 1
    for (int i = last - 1; i >= first; --i) {
        for (int e = lane; e < D::SREC; e += 32) sa[e] = in_b[(size_t)i * D::SREC + e];
        __syncwarp();
        const double *Pa = sa + D::SUM_P, *Fa = sa + D::SUM_F, *Ca = sa + D::SUM_C, *pa = sa + D::SUM_p, *fa = sa + D::SUM_f;
        const double *Pb = sb + D::SUM_P, *Fb = sb + D::SUM_F, *Cb = sb + D::SUM_C, *pb = sb + D::SUM_p, *fb = sb + D::SUM_f;
        // Aug = [I + C_a P_b | F_a | C_a | f_a]
        {
            constexpr Tile tl = pick_tile(NX, NX, 32);
            auto la = [&](int r, int k) { return Ca[r + k * NX]; };
            auto lb = [&](int k, int c) { return Pb[k + c * NX]; };
            auto epi = [&](int r, int c, double v) { Aug[r + c * LDA] = v + ((r == c) ? 1.0 : 0.0); };
            group_mm<NX, NX, NX, tl.tm, tl.tn, 32>(lane, la, lb, epi);
            for (int e = lane; e < N2; e += 32) {
                const int r = e % NX, c = e / NX;
                Aug[r + (NX + c) * LDA] = Fa[e];
                Aug[r + (2 * NX + c) * LDA] = Ca[e];
            }
            for (int r = lane; r < NX; r += 32) Aug[r + 3 * NX * LDA] = fa[r];
        }
        __syncwarp();
        warp_gauss_jordan<NX, D::NCOL, LDA>(lane, Aug);
        const double* XF = Aug + NX * LDA;
        const double* XC = Aug + 2 * NX * LDA;
        const double* wf = Aug + 3 * NX * LDA;
        // x_f = w_f - X_C p_b ; down-sweep record of member i
        double* ddi = dd_b + (size_t)i * D::DREC;
        for (int r = lane; r < NX; r += 32) {
            double acc = wf[r];
#pragma unroll 4
            for (int k = 0; k < NX; ++k) acc = fma(-XC[r + k * LDA], pb[k], acc);
            xf[r] = acc;
            ddi[D::DD_wf + r] = wf[r];
            ddi[D::DD_pb + r] = pb[r];
        }
        for (int e = lane; e < N2; e += 32) {
            const int r = e % NX, c = e / NX;
            ddi[D::DD_XF + e] = XF[r + c * LDA];
            ddi[D::DD_XC + e] = XC[r + c * LDA];
            ddi[D::DD_PB + e] = Pb[e];
            ddi[D::DD_FB + e] = Fb[e];
        }
        // T1 = P_b X_F ; T2 = F_b X_C
        {
            constexpr Tile t1 = pick_tile(NX, NX, 32);
            auto lb = [&](int k, int c) { return XF[k + c * LDA]; };
            auto lb2 = [&](int k, int c) { return XC[k + c * LDA]; };
            auto e1 = [&](int r, int c, double v) { T1[r + c * NX] = v; };
            auto e2 = [&](int r, int c, double v) { T2[r + c * NX] = v; };
            auto laP = [&](int r, int k) { return Pb[r + k * NX]; };
            auto laF = [&](int r, int k) { return Fb[r + k * NX]; };
            group_mm<NX, NX, NX, t1.tm, t1.tn, 32>(lane, laP, lb, e1);
            group_mm<NX, NX, NX, t1.tm, t1.tn, 32>(lane, laF, lb2, e2);
        }
        __syncwarp();
        // lv = P_b x_f + p_b
        for (int r = lane; r < NX; r += 32) {
            double acc = pb[r];
#pragma unroll 4
            for (int k = 0; k < NX; ++k) acc = fma(Pb[r + k * NX], xf[k], acc);
            lv[r] = acc;
        }
        __syncwarp();
        // new suffix
        {
            constexpr Tile t1 = pick_tile(NX, NX, 32);
            // P = P_a + F_a^T T1
            auto laFt = [&](int r, int k) { return Fa[k + r * NX]; };
            auto lbT1 = [&](int k, int c) { return T1[k + c * NX]; };
            auto eP = [&](int r, int c, double v) { sn[D::SUM_P + r + c * NX] = Pa[r + c * NX] + v; };
            group_mm<NX, NX, NX, t1.tm, t1.tn, 32>(lane, laFt, lbT1, eP);
            // F = F_b X_F
            auto laF = [&](int r, int k) { return Fb[r + k * NX]; };
            auto lbXF = [&](int k, int c) { return XF[k + c * LDA]; };
            auto eF = [&](int r, int c, double v) { sn[D::SUM_F + r + c * NX] = v; };
            group_mm<NX, NX, NX, t1.tm, t1.tn, 32>(lane, laF, lbXF, eF);
            // C = C_b + T2 F_b^T
            auto laT2 = [&](int r, int k) { return T2[r + k * NX]; };
            auto lbFt = [&](int k, int c) { return Fb[c + k * NX]; };
            auto eC = [&](int r, int c, double v) { sn[D::SUM_C + r + c * NX] = Cb[r + c * NX] + v; };
            group_mm<NX, NX, NX, t1.tm, t1.tn, 32>(lane, laT2, lbFt, eC);
            // p = p_a + F_a^T lv ; f = F_b x_f + f_b
            for (int r = lane; r < NX; r += 32) {
                double ap = pa[r], af = fb[r];
#pragma unroll 4
                for (int k = 0; k < NX; ++k) {
                    ap = fma(Fa[k + r * NX], lv[k], ap);
                    af = fma(Fb[r + k * NX], xf[k], af);
                }
                sn[D::SUM_p + r] = ap;
                sn[D::SUM_f + r] = af;
            }
        }
        __syncwarp();
        for (int e = lane; e < D::SREC; e += 32) sb[e] = sn[e];
        __syncwarp();
    }
    if (p.sum_out) {
        double* out = p.sum_out + ((size_t)b * p.groups + g) * D::SREC;
        for (int e = lane; e < D::SREC; e += 32) out[e] = sb[e];
    }
}

// Affine-only up-sweep (after backward_without_factorization): the members' p, f changed, their P, F, C and
// hence X_F, X_C, P_b, F_b did not.  With W = I - X_C P_b:
//     x_f = f_a - X_C (P_b f_a + p_b) ;  w_f = x_f + X_C p_b ;  p = p_a + F_a^T (P_b x_f + p_b) ;  f = F_b x_f + f_b
// Replaces the (p, c)-only update_segment_data + condensed forward's first loop (condensed_system.hpp:76-80,
// :106-118 / :197-201, :253-271).
template <int NX>
__global__ void __launch_bounds__(32) tree_up_affine_kernel(TreeParams p) {
    using D = TreeDims<NX>;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x / p.groups, g = blockIdx.x % p.groups;
    const int first = g * p.R, last = min(first + p.R, p.count) - 1;
    double* pb = smem;            // suffix p
    double* fb = smem + NX;       // suffix f
    double* t0 = smem + 2 * NX;   // P_b f_a + p_b
    double* xf = smem + 3 * NX;
    double* lv = smem + 4 * NX;
    const double* in_b = p.sum_in + (size_t)b * p.count * D::SREC;
    double* dd_b = p.dd + (size_t)b * p.count * D::DREC;
    for (int r = lane; r < NX; r += 32) {
        pb[r] = in_b[(size_t)last * D::SREC + D::SUM_p + r];
        fb[r] = in_b[(size_t)last * D::SREC + D::SUM_f + r];
    }
    __syncwarp();
#pragma unroll 1
    for (int i = last - 1; i >= first; --i) {
        const double* sa = in_b + (size_t)i * D::SREC;
        double* ddi = dd_b + (size_t)i * D::DREC;
        for (int r = lane; r < NX; r += 32) {
            double acc = pb[r];
#pragma unroll 4
            for (int k = 0; k < NX; ++k) acc = fma(ddi[D::DD_PB + r + k * NX], sa[D::SUM_f + k], acc);
            t0[r] = acc;
        }
        __syncwarp();
        for (int r = lane; r < NX; r += 32) {
            double acc = sa[D::SUM_f + r], accw = 0.0;
#pragma unroll 4
            for (int k = 0; k < NX; ++k) {
                acc = fma(-ddi[D::DD_XC + r + k * NX], t0[k], acc);
                accw = fma(ddi[D::DD_XC + r + k * NX], pb[k], accw);
            }
            xf[r] = acc;
            ddi[D::DD_wf + r] = acc + accw;
            ddi[D::DD_pb + r] = pb[r];
        }
        __syncwarp();
        for (int r = lane; r < NX; r += 32) {
            double acc = pb[r];
#pragma unroll 4
            for (int k = 0; k < NX; ++k) acc = fma(ddi[D::DD_PB + r + k * NX], xf[k], acc);
            lv[r] = acc;
        }
        __syncwarp();
        double pnew = 0.0, fnew = 0.0;
        for (int r = lane; r < NX; r += 32) {   // NX <= 32 rows per pass; results kept per lane
            double ap = sa[D::SUM_p + r], af = fb[r];
#pragma unroll 4
            for (int k = 0; k < NX; ++k) {
                ap = fma(sa[D::SUM_F + k + r * NX], lv[k], ap);
                af = fma(ddi[D::DD_FB + r + k * NX], xf[k], af);
            }
            pnew = ap; fnew = af;
        }
        __syncwarp();
        for (int r = lane; r < NX; r += 32) { pb[r] = pnew; fb[r] = fnew; }
        __syncwarp();
    }
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
    for (int i = first; i < last; ++i) {
        const double* ddi = dd_b + (size_t)i * D::DREC;
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
            lo[(size_t)i * NX + r] = acc;
            xo[(size_t)(i + 1) * NX + r] = xn[r];
            x[r] = xn[r];
        }
        __syncwarp();
    }
}

}  // namespace pdplqr
