// Register-resident stage kernel: ONE WARP per (problem, segment), the value function P, the sensitivities F, C and every
// intermediate product live in FP64 tensor-core accumulator fragments for the whole segment; shared memory holds only
// the TMA landing buffer of [E c] and a few small exchange arrays.  Unconstrained stages, nx % 4 == 0, nx <= 16, nu <= 8
// (SegDims::WLAY); everything else runs on seg_backward_kernel (seg_kernels.cuh) -- same mathematics, same records.
//
// Replaces (like seg_backward_kernel):
//   LQRParallelSolver::update_problem_data / reduction_per_thread   /root/reference include/clqr/lqr/lqr_solver_parallel.hpp:115-188
//   LQRKernel::step_with_factorization                              lqr_kernel.hpp:103-147
//   ParallelLQRKernel::step_with_factorization                      lqr_kernel_parallel.hpp:87-136
//
// Why: the round-2 clock64() breakdown of seg_backward_kernel (profiles/r2_phase_clocks.txt) showed a stage spending its
// time in shared-memory round trips between the products (accumulator -> shared -> operand, ~325 wavefronts per stage,
// the SM's LSU pipe ~70 % busy at 14 warps per SM), not in the tensor pipe.  An m8n8k4 accumulator fragment (lane (r, q)
// of a quad-row holds columns 2q, 2q+1 of row r) IS an operand fragment of the next product if the contraction index is
// enumerated in "accumulator order" (step (b, e): lane q contracts index 8b + 2q + e) -- a free choice, as long as both
// operands use it.  With the stage record stored in that row order and with x before u (common.cuh, pack_model_kernel):
//     T   = [A B c]^T P+          B operand = the accumulator fragments of P+ (P+ symmetric)            no data movement
//     FE  = F+ [A B c]            A operand = the accumulator fragments of F+                           no data movement
//     M'  = H~' + [A B]^T (PE)    B operand = the accumulator fragments of T (+ p+ on the affine row)   no data movement
//     P   = Qxx + Qxu K           accumulators initialised with M' in place (x first: Qxx sits at the tile origin)
//     F   = F+A + (F+B) K         accumulators initialised with FE in place
//     C  += (F+B)(-Gt)
// and the [E c] fragments are loaded from shared memory ONCE per stage (they are the A operand of T and M' and the B operand
// of FE).  Only the nu columns Qxu / Quu / F+B, the affine column and Z = [K | d | Gt] pass through shared memory (the
// per-lane L D L^T solve of seg_backward_kernel needs them by column).  H~, h~ - sigma w_prev come straight from global
// memory into the accumulators of M' (prefetched one stage ahead), so the TMA copy is the 1.6 KB [E c] prefix only.
#pragma once
#include "seg_kernels.cuh"

namespace pdplqr {

template <int NX, int NU>
struct WarpSmem {
    using D = SegDims<NX, NU>;
    static constexpr int S = D::S;
    static constexpr int XT = (NX + 7) / 8;        // 8-wide tiles over the state index
    static constexpr int ST = (S + 7) / 8;         // ... over the stage variable w' = [x; u]
    static constexpr int S1T = (S + 8) / 8;        // ... over [x; u; 1] (affine column / row at index S)
    static constexpr int NS = NX / 4;              // contraction steps over the state index (accumulator order)
    static constexpr int UK = (NU + 3) / 4;        // contraction steps over the input index (natural order)
    static constexpr int LDMS = S <= 20 ? 20 : 36; // Ms: M'[:, NX..S) (S x NU), leading dimension = 4 (mod 16):
    static constexpr int LDFB = 20;                // conflict-free operand-fragment reads (BwdSmem); FBs: (F+B) (NX x NU)
    static constexpr int o_rec = 0;                                  // [E c] (TMA destination)
    static constexpr int o_Z = o_rec + D::REC_EC;                    // Z = [K | d | Gt]
    static constexpr int o_Ms = o_Z + D::FREC;
    static constexpr int o_FB = o_Ms + LDMS * NU;
    static constexpr int o_gs = o_FB + LDFB * NU;                    // g' = M'[:, S]   (S)
    static constexpr int o_fc = o_gs + even_up(S);                   // F+ c            (NX)
    static constexpr int o_pn = o_fc + NX;
    static constexpr int o_fn = o_pn + NX;
    static constexpr int o_ts = o_fn + NX;                           // t = P+ c + p+   (NX)
    static constexpr int o_bar = even_up(o_ts + NX);
    static constexpr int DOUBLES = o_bar + 2;
    static constexpr size_t BYTES = (size_t)DOUBLES * 8;
};

// operand fragment of contraction step (b, e) taken from an accumulator fragment pair x[0..1] (columns 2q, 2q+1 of tile b):
// plain step: the lane's own slot e; merged step of a trailing group of 4 (lane q contracts 8b + 2 (q & 1) + (q >> 1)):
// lanes 0, 1 their slot 0, lanes 2, 3 the slot 1 of lanes 0, 1
template <bool MERGED>
PDPLQR_DEVINL double frag_from_acc(const double (&x)[2], int e, int q) {
    if constexpr (MERGED) {
        const double o = __shfl_xor_sync(0xffffffffu, x[1], 2);
        return q < 2 ? x[0] : o;
    } else {
        return x[e];
    }
}

#ifndef PDPLQR_WARP_MINB
#define PDPLQR_WARP_MINB 16    // minimum resident CTAs per SM asked of the compiler (register cap 65536 / (32 x 16) = 128:
                               // measured at C5, kernel ms: no cap (190 regs) 1.77, 16 -> 1.62, 20 (96 regs, spills) 1.80, 24 -> 2.02)
#endif
template <int NX, int NU>
__global__ void __launch_bounds__(32, PDPLQR_WARP_MINB) seg_backward_warp_kernel(SegParams p) {
    using D = SegDims<NX, NU>;
    using L = WarpSmem<NX, NU>;
    static_assert(D::WLAY, "record layout of the warp kernel");
    constexpr int S = D::S, XT = L::XT, ST = L::ST, S1T = L::S1T, NS = L::NS, UK = L::UK;
    constexpr int LDMS = L::LDMS, LDFB = L::LDFB;
    // When S is a multiple of 8 the affine index S would open a tile row / column of its own in T, FE and M' (18 of the 66
    // tensor ops of a stage at nx12/nu4 for one useful row or column each, on a kernel that is tensor-pipe bound):
    // the affine parts  t = P+ c + p+,  F+ c  and  g' = h~' + [A B]^T t  are then formed with FMAs on the accumulator /
    // operand fragments already in registers + quad reductions, and exchanged through shared memory.
    constexpr bool AFX = (S % 8 == 0);
    constexpr int TT = AFX ? ST : S1T;      // tiles over the record columns that ride on the tensor pipe
    extern __shared__ __align__(16) double smem[];
    pdl_wait();      // the predecessor kernel of the solve chain has completed (no-op without the launch attribute)
    pdl_trigger();   // the next kernel of the chain may be scheduled from here on
    const int lane = threadIdx.x, r = lane >> 2, q = lane & 3;
    const int g = blockIdx.x;
    const int b = g / p.S, seg = g % p.S;
    const int N0 = seg_first(p, seg), LEN = seg_first(p, seg + 1) - N0, N1 = N0 + LEN;
    const bool is_last = (seg == p.S - 1) && !p.interior;
    const bool pdp = !is_last;

    double* rec = smem + L::o_rec;
    double* Z = smem + L::o_Z;
    double* Ms = smem + L::o_Ms;
    double* FBs = smem + L::o_FB;
    double* gs = smem + L::o_gs;
    double* fcs = smem + L::o_fc;
    double* pn = smem + L::o_pn;
    double* fn = smem + L::o_fn;
    double* ts = smem + L::o_ts;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::o_bar);

    const size_t ws_len = (size_t)p.N * S + NX;
    const double* model_b = p.model + (size_t)b * p.N * D::REC;
    const double* ws_b = p.ws_prev ? p.ws_prev + (size_t)b * ws_len : nullptr;
    double* fac_b = p.fac + (size_t)b * p.N * D::FREC;
    const double sigma = p.sigma;

    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_fence_init();
    }
    // segment terminal condition (lqr_kernel_parallel.hpp:51-67; lqr_kernel.hpp:79-91 for the last segment), as accumulator
    // fragments: lane (r, q) holds elements (8a + r, 8c + 2q + e)
    double Pc[XT][XT][2], Fc[XT][XT][2], Cc[XT][XT][2];
#pragma unroll
    for (int a = 0; a < XT; ++a)
#pragma unroll
        for (int c = 0; c < XT; ++c)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = 8 * a + r, j = 8 * c + 2 * q + e;
                const bool in = i < NX && j < NX;
                Pc[a][c][e] = (is_last && in) ? p.HN[(size_t)b * NX * NX + i + j * NX] + ((i == j) ? sigma : 0.0) : 0.0;
                Fc[a][c][e] = (!is_last && in && i == j) ? 1.0 : 0.0;
                Cc[a][c][e] = 0.0;
            }
    for (int i = lane; i < NX; i += 32) {
        pn[i] = is_last ? p.hN[(size_t)b * NX + i] - (ws_b ? sigma * ws_b[(size_t)p.N * S + i] : 0.0) : 0.0;
        fn[i] = 0.0;
    }
    __syncwarp();
    auto issue_stage = [&](int kk) {   // lane 0: the [E c] prefix of the stage record
        mbar_expect_tx(&bar[0], D::REC_EC * 8);
        bulk_g2s(rec, model_b + (size_t)kk * D::REC, D::REC_EC * 8, &bar[0]);
    };
    // [H~' | h~' - sigma w_prev] of a stage straight from global memory, as the initial accumulator fragments of M':
    // element (m', n') = H'(n', m') (symmetric, lqr_model.hpp:18): the pair e = 0, 1 is contiguous in column m'
    // The loaded values are only MOVED into registers here; sigma w_prev is folded in when the accumulators are initialised a
    // stage later (arithmetic on a loaded value at this point made the warp wait for the whole DRAM round trip: 2,000 of
    // the stage's 5,800 cycles in the first version, profiles/r2_phase_clocks.txt).
    double Hc[ST][TT][2], Wc[ST], hv[ST];
    auto load_H = [&](int kk) {
        const double* Rg = model_b + (size_t)kk * D::REC;
#pragma unroll
        for (int mt = 0; mt < ST; ++mt) {
            const int m = 8 * mt + r;
            Wc[mt] = (ws_b && m < S) ? ws_b[(size_t)kk * S + widx_inv(m, NX, NU, true)] : 0.0;
            hv[mt] = (AFX && m < S) ? Rg[D::REC_h + m] : 0.0;
#pragma unroll
            for (int nt = 0; nt < TT; ++nt) {
                if constexpr (S % 2 == 0) {
                    if (8 * nt + 7 < S) {   // compile-time: a tile of H proper
                        double2 v = make_double2(0.0, 0.0);
                        if (m < S) v = *reinterpret_cast<const double2*>(Rg + D::REC_H + D::h_off(8 * nt + 2 * q, m));
                        Hc[mt][nt][0] = v.x;
                        Hc[mt][nt][1] = v.y;
                        continue;
                    }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = 8 * nt + 2 * q + e;
                    double v = 0.0;
                    if (m < S) {
                        if (n < S) v = Rg[D::REC_H + D::h_off(n, m)];
                        else if (n == S) v = Rg[D::REC_h + m];
                    }
                    Hc[mt][nt][e] = v;
                }
            }
        }
    };
    if (LEN > 0) {
        if (lane == 0) issue_stage(N1 - 1);
        load_H(N1 - 1);
    }

    int bad = 0;
    // A warp issues in order: an instruction that consumes a shuffle result stalls everything behind it for the ~28 cycles
    // of the shuffle.  The source order below therefore issues shuffles EARLY and consumes them LATE, with independent
    // tensor ops in between (the affine reductions are software-pipelined through the contraction steps of the products).
    constexpr bool HAS_MG = (NX % 8 == 4);          // the last contraction step is a merged one
    constexpr int S_MG = NS - 1;
    PHASE_DECL
    PHASE_START();
#pragma unroll 1
    for (int it = 0; it < LEN; ++it) {
        const int k = N1 - 1 - it;
        // merged-step operand fragments of P+ and F+ (shuffles; consumed by the LAST step of the first products)
        double Pm[XT], Fm[XT];
        if constexpr (HAS_MG) {
#pragma unroll
            for (int c = 0; c < XT; ++c) {
                Pm[c] = frag_from_acc<true>(Pc[c][S_MG / 2], 0, q);
                Fm[c] = frag_from_acc<true>(Fc[c][S_MG / 2], 0, q);
            }
        }
        mbar_wait(&bar[0], it & 1);
        PHASE(0);
        // ---- [E c] fragments, once per stage: Ef[s][t] = Ec(row of step s / lane q, record column 8t + r)
        double Ef[NS][TT];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int t = 0; t < TT; ++t) Ef[s][t] = (8 * t + r <= S) ? rec[(4 * s + q) + (8 * t + r) * NX] : 0.0;
        double cval[XT][2];   // AFX: c at the columns the lane's accumulator slots own
        if constexpr (AFX) {
#pragma unroll
            for (int c = 0; c < XT; ++c)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = 8 * c + 2 * q + e;
                    cval[c][e] = i < NX ? rec[D::er(i) + S * NX] : 0.0;
                }
        }
        __syncwarp();
        if (lane == 0 && it + 1 < LEN) {   // the buffer is free again: fetch stage k-1 while this stage computes
            fence_proxy_async();
            issue_stage(k - 1);
        }
        PHASE(1);
        // AFX: t = P+ c + p+ and F+ c as row dot products over the accumulator fragments (partial sums per lane here, the
        // two quad-reduction steps and the store ride along with contraction steps 0, 1, 2 of the products below)
        double at[XT], af[XT], sh1[XT], sh2[XT];
        if constexpr (AFX) {
#pragma unroll
            for (int a = 0; a < XT; ++a) {
                at[a] = 0.0; af[a] = 0.0;
#pragma unroll
                for (int c = 0; c < XT; ++c)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        at[a] = fma(Pc[a][c][e], cval[c][e], at[a]);
                        if (pdp) af[a] = fma(Fc[a][c][e], cval[c][e], af[a]);
                    }
            }
        }
        auto aff_pipe = [&](int st) {
            if constexpr (AFX) {
#pragma unroll
                for (int a = 0; a < XT; ++a) {
                    if (st == 0) {
                        sh1[a] = __shfl_xor_sync(0xffffffffu, at[a], 1);
                        if (pdp) sh2[a] = __shfl_xor_sync(0xffffffffu, af[a], 1);
                    } else if (st == 1) {
                        at[a] += sh1[a];
                        if (pdp) af[a] += sh2[a];
                        sh1[a] = __shfl_xor_sync(0xffffffffu, at[a], 2);
                        if (pdp) sh2[a] = __shfl_xor_sync(0xffffffffu, af[a], 2);
                    } else {
                        const int i = 8 * a + r;
                        if (q == 0 && i < NX) {
                            ts[i] = at[a] + sh1[a] + pn[i];
                            if (pdp) fcs[i] = af[a] + sh2[a];
                        }
                    }
                }
            }
        };
        // ---- T = [A B c]^T P+  ((S+1) x NX)  and  FE = F+ [A B c]  (NX x (S+1))
        double Tc[TT][XT][2], FEc[XT][TT][2];
#pragma unroll
        for (int t = 0; t < TT; ++t)
#pragma unroll
            for (int c = 0; c < XT; ++c) Tc[t][c][0] = Tc[t][c][1] = FEc[c][t][0] = FEc[c][t][1] = 0.0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int bs = s / 2, es = s & 1;
            const bool mg = HAS_MG && s == S_MG;
            double Pf[XT], Ff[XT];
#pragma unroll
            for (int c = 0; c < XT; ++c) {
                Pf[c] = mg ? Pm[c] : Pc[c][bs][es];
                Ff[c] = mg ? Fm[c] : Fc[c][bs][es];
            }
#pragma unroll
            for (int t = 0; t < TT; ++t)
#pragma unroll
                for (int c = 0; c < XT; ++c) dmma_m8n8k4(Tc[t][c][0], Tc[t][c][1], Ef[s][t], Pf[c]);
            if (pdp) {
#pragma unroll
                for (int a = 0; a < XT; ++a)
#pragma unroll
                    for (int t = 0; t < TT; ++t) dmma_m8n8k4(FEc[a][t][0], FEc[a][t][1], Ff[a], Ef[s][t]);
            }
            if (s < 3) aff_pipe(s);
        }
#pragma unroll
        for (int st = NS; st < 3; ++st) aff_pipe(st);
        // affine row of T:  (P+ c)^T + p+^T
        if (!AFX && r == S % 8) {
#pragma unroll
            for (int c = 0; c < XT; ++c)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = 8 * c + 2 * q + e;
                    if (j < NX) Tc[S / 8][c][e] += pn[j];
                }
        }
        PHASE(2);
        // merged-step operand fragments of T (shuffles; consumed by the last step of the next product)
        double Tm[TT];
        if constexpr (HAS_MG) {
#pragma unroll
            for (int nt = 0; nt < TT; ++nt) Tm[nt] = frag_from_acc<true>(Tc[nt][S_MG / 2], 0, q);
        }
        double tk[NS];        // AFX: t at the contraction index the lane owns in step s
        if constexpr (AFX) {
            __syncwarp();     // ts was written by the quads' first lanes during the products
#pragma unroll
            for (int s = 0; s < NS; ++s) tk[s] = ts[D::eri(4 * s + q)];
        }
        // ---- M' = [H~' | h~'] + [A B]^T (PE)   (S x (S+1)); PE = T^T is the B operand straight from T's accumulators
        double Mc[ST][TT][2];
#pragma unroll
        for (int mt = 0; mt < ST; ++mt)
#pragma unroll
            for (int nt = 0; nt < TT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m = 8 * mt + r, n = 8 * nt + 2 * q + e;
                    double v = Hc[mt][nt][e];
                    if (m < S && n == m) v += sigma;                       // H + sigma I      (lqr_solver_parallel.hpp:129-130)
                    if (m < S && n == S) v = fma(-sigma, Wc[mt], v);       // h - sigma w_prev (lqr_solver_parallel.hpp:131-132)
                    Mc[mt][nt][e] = v;
                }
        // AFX: g' = h~' - sigma w_prev + [A B]^T t as column dot products over the [E c] fragments, pipelined like t
        double ag[ST], sg[ST];
        auto g_pipe = [&](int st) {
            if constexpr (AFX) {
#pragma unroll
                for (int mt = 0; mt < ST; ++mt) {
                    if (st == 0) {
                        ag[mt] = 0.0;
#pragma unroll
                        for (int s = 0; s < NS; ++s) ag[mt] = fma(Ef[s][mt], tk[s], ag[mt]);
                        sg[mt] = __shfl_xor_sync(0xffffffffu, ag[mt], 1);
                    } else if (st == 1) {
                        ag[mt] += sg[mt];
                        sg[mt] = __shfl_xor_sync(0xffffffffu, ag[mt], 2);
                    } else {
                        const int m = 8 * mt + r;
                        if (q == 0 && m < S) gs[m] = fma(-sigma, Wc[mt], hv[mt]) + ag[mt] + sg[mt];
                    }
                }
            }
        };
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int bs = s / 2, es = s & 1;
            const bool mg = HAS_MG && s == S_MG;
            double Bf[TT];
#pragma unroll
            for (int nt = 0; nt < TT; ++nt) Bf[nt] = mg ? Tm[nt] : Tc[nt][bs][es];
#pragma unroll
            for (int mt = 0; mt < ST; ++mt)
#pragma unroll
                for (int nt = 0; nt < TT; ++nt) dmma_m8n8k4(Mc[mt][nt][0], Mc[mt][nt][1], Ef[s][mt], Bf[nt]);
            if (s < 3) g_pipe(s);
        }
#pragma unroll
        for (int st = NS; st < 3; ++st) g_pipe(st);
        PHASE(4);
        if (it + 1 < LEN) load_H(k - 1);   // next stage's H~, h~ (consumed a whole stage from now)
        // ---- the nu columns Qxu / Quu, F+B and the affine columns go to shared memory (column access for the solve)
#pragma unroll
        for (int mt = 0; mt < ST; ++mt)
#pragma unroll
            for (int nt = 0; nt < TT; ++nt) {
                if (8 * nt + 7 < NX) continue;   // compile-time: a tile of Qxx only
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m = 8 * mt + r, n = 8 * nt + 2 * q + e;
                    if (m < S) {
                        if (n >= NX && n < S) Ms[m + (n - NX) * LDMS] = Mc[mt][nt][e];
                        else if (n == S) gs[m] = Mc[mt][nt][e];
                    }
                }
            }
        if (pdp) {
#pragma unroll
            for (int a = 0; a < XT; ++a)
#pragma unroll
                for (int nt = 0; nt < TT; ++nt) {
                    if (8 * nt + 7 < NX) continue;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int i = 8 * a + r, n = 8 * nt + 2 * q + e;
                        if (i < NX) {
                            if (n >= NX && n < S) FBs[i + (n - NX) * LDFB] = FEc[a][nt][e];
                            else if (n == S) fcs[i] = FEc[a][nt][e];
                        }
                    }
                }
        }
        __syncwarp();
        PHASE(5);
        // ---- Quu = L D L^T in registers (every lane, redundantly), one right-hand side per lane:
        //      z = -Quu^-1 r,  r in [Qux | Qu | (F+B)^T]   ->   Z = [K | d | Gt]        (as in seg_backward_kernel)
        {
            double Lr[NU][NU], dr[NU], dd[NU];
#pragma unroll
            for (int j = 0; j < NU; ++j)
#pragma unroll
                for (int i = j + 1; i < NU; ++i) Lr[i][j] = Ms[(NX + i) + j * LDMS];
#pragma unroll
            for (int c = 0; c < NU; ++c) {
                double vc[NU];
#pragma unroll
                for (int qq = 0; qq < c; ++qq) vc[qq] = Lr[c][qq] * dd[qq];
                double a = Ms[(NX + c) + c * LDMS];
#pragma unroll
                for (int qq = 0; qq < c; ++qq) a = fma(-Lr[c][qq], vc[qq], a);
                double rc = rcp_newton(a);
                if (!(a > 0.0)) {   // (rare) keep the sweep finite, report through the status word
                    if (!bad) bad = k + 1;
                    a = fabs(a) + 1e-300;
                    rc = rcp_newton(a);
                }
                dr[c] = rc;
                dd[c] = a;
#pragma unroll
                for (int i = c + 1; i < NU; ++i) {
                    double v = Lr[i][c];
#pragma unroll
                    for (int qq = 0; qq < c; ++qq) v = fma(-Lr[i][qq], vc[qq], v);
                    Lr[i][c] = v * rc;
                }
            }
            PHASE(6);
            const int nrhs = NX + 1 + (pdp ? NX : 0);
            double* fk = fac_b + (size_t)k * D::FREC;
            for (int c = lane; c < nrhs; c += 32) {
                const double* src;
                int stride;
                if (c < NX) { src = Ms + c; stride = LDMS; }                    // Qux(m, c) = Qxu(c, m)
                else if (c == NX) { src = gs + NX; stride = 1; }                // Qu(m)
                else { src = FBs + (c - NX - 1); stride = LDFB; }               // (F+B)(c', m)
                double y[NU], z[NU];
#pragma unroll
                for (int m = 0; m < NU; ++m) y[m] = src[m * stride];
#pragma unroll
                for (int m = 1; m < NU; ++m) {
                    double v = y[m];
#pragma unroll
                    for (int qq = 0; qq < m; ++qq) v = fma(-Lr[m][qq], y[qq], v);
                    y[m] = v;
                }
#pragma unroll
                for (int m = NU - 1; m >= 0; --m) {
                    double v = y[m] * dr[m];
#pragma unroll
                    for (int qq = m + 1; qq < NU; ++qq) v = fma(-Lr[qq][m], z[qq], v);
                    z[m] = v;
                }
                if constexpr (NU % 2 == 0) {   // factor record -> shared (operands below) and global, 16-byte vectors
#pragma unroll
                    for (int m = 0; m < NU; m += 2) {
                        const double2 v = make_double2(-z[m], -z[m + 1]);
                        *reinterpret_cast<double2*>(Z + m + c * NU) = v;
                        *reinterpret_cast<double2*>(fk + m + c * NU) = v;
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < NU; ++m) {
                        Z[m + c * NU] = -z[m];
                        fk[m + c * NU] = -z[m];
                    }
                }
            }
        }
        PHASE(7);
        __syncwarp();
        PHASE(8);
        // ---- P = Qxx + Qxu K ;  F = F+A + (F+B) K ;  C += (F+B)(-Gt)      (contraction over the inputs, natural order)
        {
            double Aq[UK][XT], Afb[UK][XT], Bk[UK][XT], Bg[UK][XT];
#pragma unroll
            for (int kt = 0; kt < UK; ++kt)
#pragma unroll
                for (int a = 0; a < XT; ++a) {
                    const int m = 4 * kt + q, i = 8 * a + r;
                    const bool in = m < NU && i < NX;
                    Aq[kt][a] = in ? Ms[i + m * LDMS] : 0.0;                              // Qxu(i, m)
                    Afb[kt][a] = (in && pdp) ? FBs[i + m * LDFB] : 0.0;                   // (F+B)(i, m)
                    Bk[kt][a] = in ? Z[m + i * NU] : 0.0;                                  // K(m, j = i)
                    Bg[kt][a] = (in && pdp) ? -Z[m + (NX + 1 + i) * NU] : 0.0;            // -Gt(m, j = i)
                }
#pragma unroll
            for (int a = 0; a < XT; ++a)
#pragma unroll
                for (int c = 0; c < XT; ++c) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const bool in = (8 * a + r < NX) && (8 * c + 2 * q + e < NX);
                        Pc[a][c][e] = in ? Mc[a][c][e] : 0.0;
                        Fc[a][c][e] = (in && pdp) ? FEc[a][c][e] : 0.0;
                    }
#pragma unroll
                    for (int kt = 0; kt < UK; ++kt) {
                        dmma_m8n8k4(Pc[a][c][0], Pc[a][c][1], Aq[kt][a], Bk[kt][c]);
                        if (pdp) {
                            dmma_m8n8k4(Fc[a][c][0], Fc[a][c][1], Afb[kt][a], Bk[kt][c]);
                            dmma_m8n8k4(Cc[a][c][0], Cc[a][c][1], Afb[kt][a], Bg[kt][c]);
                        }
                    }
                }
        }
        // P = Qxx + Qxu K is symmetric only up to rounding, and the asymmetry is carried from stage to stage (it reached 6e-7
        // relative over the 253-stage segments of the 2^20-stage problem).  seg_backward_kernel mirrors the lower triangle
        // every stage; here that costs 24 shuffles, so it is done every 4th stage and at the segment entry (a numpy
        // experiment on the quadrotor: error vs every-stage mirroring 1e-15 at every 4, 8e-15 at 8, 1e-13 at 16, 2e-5 never).
        if ((it & 3) == 3 || it + 1 == LEN) {
            double v0[XT][XT][2], v1[XT][XT][2];   // all shuffles first (lower tiles c <= a), the selects afterwards
#pragma unroll
            for (int a = 0; a < XT; ++a)
#pragma unroll
                for (int c = 0; c <= a; ++c)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int sl = 4 * (2 * q + e) + (r >> 1);   // lane that holds element (2q + e, r) of the tile
                        v0[a][c][e] = __shfl_sync(0xffffffffu, Pc[a][c][0], sl);
                        v1[a][c][e] = __shfl_sync(0xffffffffu, Pc[a][c][1], sl);
                    }
#pragma unroll
            for (int a = 0; a < XT; ++a)
#pragma unroll
                for (int c = 0; c <= a; ++c)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double tv = (r & 1) ? v1[a][c][e] : v0[a][c][e];
                        if (c < a) Pc[c][a][e] = tv;
                        else if (2 * q + e > r) Pc[a][a][e] = tv;
                    }
        }
        // p = Qx + Qxu d ;  f = F+c + (F+B) d + f+
        for (int i = lane; i < NX; i += 32) {
            double ap = gs[i], af = pdp ? fcs[i] + fn[i] : 0.0;
#pragma unroll
            for (int m = 0; m < NU; ++m) {
                const double dm = Z[m + NX * NU];
                ap = fma(Ms[i + m * LDMS], dm, ap);
                if (pdp) af = fma(FBs[i + m * LDFB], dm, af);
            }
            pn[i] = ap;
            fn[i] = af;
        }
        PHASE(9);
        __syncwarp();
        PHASE(10);
    }
    PHASE_PRINT("seg_backward_warp [wait | Ef | T,FE->S2 | - | M->S3 | H prefetch + staging | LDL | solve | sync | S6 + vectors]", LEN);

    // segment summary (lqr_solver_parallel.hpp:180-187): P, F, C, p, f at the segment entry
    double* sm = p.sum + ((size_t)b * p.S + seg) * D::SREC;
#pragma unroll
    for (int a = 0; a < XT; ++a)
#pragma unroll
        for (int c = 0; c < XT; ++c)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = 8 * a + r, j = 8 * c + 2 * q + e;
                if (i < NX && j < NX) {
                    sm[D::SUM_P + i + j * NX] = Pc[a][c][e];
                    sm[D::SUM_F + i + j * NX] = Fc[a][c][e];
                    sm[D::SUM_C + i + j * NX] = Cc[a][c][e];
                }
            }
    for (int i = lane; i < NX; i += 32) {
        sm[D::SUM_p + i] = pn[i];
        sm[D::SUM_f + i] = fn[i];
    }
    if (lane == 0) p.status[(size_t)b * p.S + seg] = bad;
}

}  // namespace pdplqr
