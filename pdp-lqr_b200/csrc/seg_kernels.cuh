// Segment kernels: one group of T threads (a warp, or a CTA for large nx+nu) owns one (problem, segment) and
// runs the segment-local backward Riccati sweep with sensitivity propagation, resp. the forward rollout.
//
// Replaces (B200-first redesign, not a translation):
//   LQRParallelSolver::update_problem_data      /root/reference include/clqr/lqr/lqr_solver_parallel.hpp:115-140  (fused)
//   LQRParallelSolver::reduction_per_thread     lqr_solver_parallel.hpp:164-188
//   LQRKernel::step_with_factorization          lqr_kernel.hpp:103-147
//   ParallelLQRKernel::step_with_factorization  lqr_kernel_parallel.hpp:87-136
//   ParallelLQRKernel::forward_step / LQRKernel::forward_step   lqr_kernel_parallel.hpp:170-218 / lqr_kernel.hpp:180-212
//
// Algorithmic differences from the reference (same mathematics, results agree to rounding):
//   * the value function is carried as P_k (Schur complement Qxx - Qxu Quu^-1 Qux) instead of its Cholesky
//     factor Lxx, so only the nu x nu block Quu is factorised per stage (critical path nu instead of nx+nu);
//   * the affine terms ride along as an extra column:  [M | g] = [H~ | h~] + E^T (P+ [E c] + [0 p+]);
//   * F+ [B A c] is formed up front (off the critical path): F = F+A + (F+B) K, f = F+c + (F+B) d + f+;
//   * what is stored per stage is Z = [K | d | Gt] with K = -Quu^-1 Qux, d = -Quu^-1 Qu, Gt = Luu^-T G
//     (G as in lqr_kernel_parallel.hpp:127-128), so the rollout is u = K x + d + Gt uhat with no solves;
//   * Quu is factorised as L D L^T (no square roots: a Newton-refined reciprocal per pivot, the pivot signs are the
//     Cholesky's) and the value-function updates use Z directly:  P = Qxx + Qxu K,  p = Qx + Qxu d,
//     C += (F+B)(-Gt)  [= G^T G],  so no Y = Luu^-1 [...] is formed or stored.  Round-2 clock64() breakdown: the
//     Cholesky + substitution phase was 1,700-1,900 of the 4,000 (latency mode) / 6,700 (throughput mode) cycles of a stage.
#pragma once
#include "common.cuh"

namespace pdplqr {

constexpr int odd_ld(int n) { return n | 1; }
constexpr int even_up(int n) { return (n + 1) & ~1; }

template <int NX, int NU>
struct SegDims {
    static constexpr int S = NX + NU;
    // model record (one stage): [E (NX x S) | c (NX) | H (S x S) | h (S)], column-major, padded to 16 bytes.  [E c] is
    // an NX x (S+1) matrix with leading dimension NX: as a DMMA operand (k along the contiguous index) it is read in
    // place from the TMA buffer without shared-memory bank conflicts for NX = 2, 4, 6 (mod 8) -- no transposed copy
    static constexpr int REC_E = 0;
    static constexpr int REC_C = NX * S;
    static constexpr int REC_H = REC_C + NX;
    static constexpr int REC_h = REC_H + S * S;
    // H is only ever read as the initial value of a DMMA accumulator tile (lane -> row r, columns 2 (lane%4) + {0,1}).
    // With a leading dimension S = 0 (mod 16) the 16 lanes of a half-warp would hit 4 banks 4 times each, so for those
    // S the rows of column j are stored rotated by 4 (j / 2): element (i, j) at ((i + 4 (j/2)) mod S) + j S.
    static constexpr bool H_ROT = h_rotated(S);
    PDPLQR_DEVINL static int h_off(int i, int j) { return H_ROT ? ((i + 4 * (j >> 1)) % S) + j * S : i + j * S; }
    // record order of the rows of [E c] and of the w-indices (common.cuh): identity unless the warp kernel's layout is on.
    // Everywhere below i, j are the reference's indices (w = [u; x], lqr_model.hpp:12-19); er / wi give the position.
    static constexpr bool WLAY = warp_layout(NX, NU);
    PDPLQR_DEVINL static constexpr int er(int i) { return erow_pos(i, NX, WLAY); }           // position of row i of [E c]
    PDPLQR_DEVINL static constexpr int eri(int p) { return erow_inv(p, NX, WLAY); }          // row at position p
    PDPLQR_DEVINL static constexpr int wi(int i) { return widx_pos(i, NX, NU, WLAY); }       // position of w-index i
    PDPLQR_DEVINL static constexpr int wic(int j) { return j < S ? widx_pos(j, NX, NU, WLAY) : j; }   // ... of column j of [E c]
    PDPLQR_DEVINL static int H_at(int i, int j) { return REC_H + h_off(wi(i), wi(j)); }      // H(i,j) inside the record
    PDPLQR_DEVINL static constexpr int h_at(int i) { return REC_h + wi(i); }
    static constexpr int REC = even_up(REC_h + S);
    static constexpr int REC_EC = even_up(NX * S + NX);  // prefix the rollout needs
    // factor record (one stage): Z = [K (NU x NX) | d (NU) | Gt (NU x NX)], column-major
    static constexpr int NRHS = 2 * NX + 1;
    static constexpr int FREC = even_up(NU * NRHS);
    // affine cache (one stage), kept for backward_without_factorization: [Quu^-1 (NU x NU) | P+c | F+c | F+B (NX x NU)]
    static constexpr int AR_QI = 0, AR_PC = NU * NU, AR_FC = AR_PC + NX, AR_FB = AR_FC + NX;
    static constexpr int AREC = even_up(AR_FB + NX * NU);
    // segment summary: [P | F | C | p | f]
    static constexpr int SUM_P = 0, SUM_F = NX * NX, SUM_C = 2 * NX * NX, SUM_p = 3 * NX * NX, SUM_f = 3 * NX * NX + NX;
    static constexpr int SREC = 3 * NX * NX + 2 * NX;
};

struct SegParams {
    int N, S, batch;
    int interior;            // 1: this handle is a horizon shard whose last segment ends at an interface, not at the terminal
    int seg_mode, seg_len0;  // partition in closed form (no global loads in the kernel prologue), see seg_first()
    const int* seg_start;    // [S]  (kept for reference / debugging)
    const int* seg_len;      // [S]
    const double* model;     // [batch][N][REC]
    const double* HN;        // [batch][NX*NX]
    const double* hN;        // [batch][NX]
    const double* ws_prev;   // [batch][N*S+NX] or nullptr (== zeros)
    double sigma;
    double* fac;             // [batch][N][FREC]
    double* sum;             // [batch][S][SREC]
    int* status;             // [batch][S]  first non-positive pivot (1 + stage) met by each segment, 0 = none
    double* aff;             // [batch][N][AREC] affine cache, or nullptr (not kept)
    // constraints (all nullptr / 0 when the problem has none)
    const int* ncs;          // [N+1] rows per stage
    const long long* coff;   // [N+2] prefix offsets of the constraint vectors
    const long long* doff;   // [N+2] prefix offsets into the (16-byte padded) device copy of D
    const double* Dm;        // [batch][d_total_dev]
    long long d_total, nc_total;
    int ncmax;
    const double *ys, *zs, *rho, *inv_rho;   // [batch][nc_total]
    // selection-matrix constraints (every row of every D_k has at most one non-zero): D is replaced by the column
    // index (-1 for an all-zero row) and the value of each row; nullptr -> dense D
    const int* sel_col;      // [batch][nc_total]
    const double* sel_val;   // [batch][nc_total]
    const double* xhat;      // [batch][S][NX]  (entry state of each segment; == x0 when S == 1)
    const double* uhat;      // [batch][S][NX]  (costate at each segment's exit)
    double* ws_out;          // [batch][N*S+NX]
};

// first stage of segment i (i = S gives N).  seg_mode 0: equal lengths, the first N % S segments one stage longer
// (library-chosen segmentation); seg_mode 1: the reference's rule -- every segment but the last has seg_len0 stages
// (lqr_solver_parallel.hpp:73-80).
PDPLQR_DEVINL int seg_first(const SegParams& p, int i) {
    if (i >= p.S) return p.N;
    if (p.seg_mode == 1) return i * p.seg_len0;
    const int q = p.N / p.S, r = p.N % p.S;
    return i * q + min(i, r);
}

// compile-time choice of the register tile for an M x N product on T threads
struct Tile { int tm, tn; };
constexpr Tile pick_tile(int M, int N, int T) {
    Tile best{1, 1};
    double best_cost = 1e30;
    const int cand[][2] = {{1, 1}, {2, 1}, {1, 2}, {2, 2}, {3, 1}, {1, 3}, {2, 3}, {3, 2}, {4, 1}, {1, 4},
                           {2, 4}, {4, 2}, {3, 3}, {4, 3}, {3, 4}, {4, 4}};
    for (auto& c : cand) {
        int mt = (M + c[0] - 1) / c[0], nt = (N + c[1] - 1) / c[1];
        int rounds = (mt * nt + T - 1) / T;
        double cost = rounds * (c[0] * c[1] + 1.0 * (c[0] + c[1]) + 2.0);
        if (cost < best_cost) { best_cost = cost; best = Tile{c[0], c[1]}; }
    }
    return best;
}

template <int NX, int NU>
struct BwdSmem {
    using D = SegDims<NX, NU>;
    static constexpr int S = D::S;
    // Leading dimensions.  Bank conflicts are counted per half-warp (16 lanes; scripts/micro/lds_bench.cu).  A DMMA
    // OPERAND fragment of a column-major array (lane -> row r = lane/4, k = lane%4: address r + k ld, or k + r ld
    // for the other orientation) is conflict-free for ld = 4 or 12 (mod 16) -- hence 4 (mod 8) for PF, PFE, YT.  An
    // ACCUMULATOR fragment (lane -> row r, columns 2 (lane%4) + {0,1}: r + 2 (lane%4) ld) wants ld = 2 (mod 4): used
    // for Ma, which is never an operand; the accumulator stores into PF / PFE keep a 2-way conflict (no plain ld serves
    // both patterns; shifting every column pair by 4 elements does, but its index arithmetic cost what it saved).
    static constexpr int LDPF = ld4mod8(2 * NX);    // PF: [P+; F+] stacked, (2NX) x NX
    static constexpr int LDPE = ld4mod8(2 * NX);    // PFE: [P+;F+] [E c], (2NX) x (S+1)
    static constexpr int LDM = S + ((2 - S % 4) + 4) % 4;   // Ma: [M | g], S x (S+1)
    static constexpr int o_rec = 0;                                 // REC (TMA destination, 16B aligned), single buffer:
                                                                    // the next record is fetched right after its last reader (S3)
    static constexpr int o_Z = o_rec + D::REC;                      // FREC
    static constexpr int o_PF = o_Z + D::FREC;
    static constexpr int o_PFE = o_PF + LDPF * NX;
    static constexpr int o_Ma = o_PFE + LDPE * (S + 1);
    // C (NX x NX) accumulates over the whole segment.  For small NX it lives in the DMMA accumulator registers of the
    // group's first warp (2 x 2 tiles, 8 doubles per lane) instead of a shared-memory array that is read and written
    // every stage: fewer wavefronts and 1.1 KB less per group (14 instead of 13 CTAs/SM at nx12/nu4).
    static constexpr bool C_REGS = (NX <= 16) && ((long long)NX * NX * NU >= PDPLQR_DMMA_MIN_MACS);
    static constexpr int o_Cn = o_Ma + LDM * (S + 1);
    static constexpr int o_pn = o_Cn + (C_REGS ? 0 : NX * NX);
    static constexpr int o_fn = o_pn + NX;
    static constexpr int o_dinv = o_fn + NX;
    static constexpr int o_wp = o_dinv + NU;
    static constexpr int o_pc = o_wp + S;                           // P+ c   (NX)
    static constexpr int o_qi = o_pc + NX;                          // Quu^-1 (NU x NU)
    static constexpr int o_bar = even_up(o_qi + NU * NU);           // 2 mbarriers
    static constexpr int DOUBLES = even_up(o_bar + 2);
    static constexpr size_t BYTES = (size_t)DOUBLES * 8;
    // run-time tail (only when the problem has constraints):
    //   dense D:      D[even(ncmax*S)] | rho[ncmax] | rho.*g[ncmax]
    //   selection D:                     rho[ncmax] | rho.*g[ncmax] | val[ncmax] | 2 x (dg[S] | dh[S]) | col[ncmax] (int)
    static size_t bytes(int ncmax, bool sel = false) {
        if (ncmax <= 0) return BYTES;
        return BYTES + (size_t)((sel ? 0 : even_up(ncmax * S)) + 2 * ncmax + (sel ? ncmax + even_up(ncmax) / 2 + 1 + 4 * S : 0)) * 8;
    }
};

// ------------------------------------------------------------------------------------------------
// Backward: segment-local Riccati sweep (+ sensitivities F, f, C for non-last segments).
// CON = false compiles every constraint path out (unconstrained problems pay nothing for them).
// -DPDPLQR_PHASE_CLOCKS: thread 0 of two CTAs accumulates clock64() per phase of the stage loop and prints cycles per
// stage at the end (instrumented builds only: PDPLQR_VARIANT=prof PDPLQR_CFLAGS=-DPDPLQR_PHASE_CLOCKS, scripts/prof_phases.py)
#ifndef PDPLQR_SYM_S3
#define PDPLQR_SYM_S3 1
#endif
#ifndef PDPLQR_S3_VEC
#define PDPLQR_S3_VEC 1
#endif
#ifndef PDPLQR_AFF_PSHIFT
#define PDPLQR_AFF_PSHIFT 1
#endif
#ifndef PDPLQR_FOLD_AHEAD
#define PDPLQR_FOLD_AHEAD 1
#endif
#ifdef PDPLQR_PHASE_CLOCKS
#define PHASE_DECL long long ph_acc[12] = {0}, ph_t = 0; const bool ph_on = (threadIdx.x == 0) && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2);
#define PHASE_START() do { if (ph_on) ph_t = clock64(); } while (0)
#define PHASE(i) do { if (ph_on) { const long long t_ = clock64(); ph_acc[i] += t_ - ph_t; ph_t = t_; } } while (0)
#define PHASE_PRINT(name, n) do { if (ph_on && (n) > 0) printf("%s blk %d T %d stages %d | cycles/stage: wait %lld fold %lld S2 %lld sync %lld S3 %lld sync+issue %lld chol %lld solve %lld sync+bulk %lld S6 %lld tail %lld | total %lld\n", name, (int)blockIdx.x, (int)blockDim.x, (int)(n), ph_acc[0]/(n), ph_acc[1]/(n), ph_acc[2]/(n), ph_acc[3]/(n), ph_acc[4]/(n), ph_acc[5]/(n), ph_acc[6]/(n), ph_acc[7]/(n), ph_acc[8]/(n), ph_acc[9]/(n), ph_acc[10]/(n), (ph_acc[0]+ph_acc[1]+ph_acc[2]+ph_acc[3]+ph_acc[4]+ph_acc[5]+ph_acc[6]+ph_acc[7]+ph_acc[8]+ph_acc[9]+ph_acc[10])/(n)); } while (0)
#else
#define PHASE_DECL
#define PHASE_START() do {} while (0)
#define PHASE(i) do {} while (0)
#define PHASE_PRINT(name, n) do {} while (0)
#endif

template <int NX, int NU, int T, bool CON>
__global__ void __launch_bounds__(T) seg_backward_kernel(SegParams p) {
    using D = SegDims<NX, NU>;
    using L = BwdSmem<NX, NU>;
    constexpr int S = D::S;
    extern __shared__ __align__(16) double smem[];
    pdl_wait();      // the predecessor kernel of the solve chain has completed (no-op without the launch attribute)
    pdl_trigger();   // the next kernel of the chain may be scheduled from here on
    const int tid = threadIdx.x;
    const int g = blockIdx.x;
    const int b = g / p.S, seg = g % p.S;
    const int N0 = seg_first(p, seg), LEN = seg_first(p, seg + 1) - N0, N1 = N0 + LEN;
    const bool is_last = (seg == p.S - 1) && !p.interior;
    const bool pdp = !is_last;

    double* rec = smem + L::o_rec;
    double* Z = smem + L::o_Z;
    double* PF = smem + L::o_PF;
    double* PFE = smem + L::o_PFE;
    double* Ma = smem + L::o_Ma;
    double* Cn = smem + L::o_Cn;
    double* pn = smem + L::o_pn;
    double* fn = smem + L::o_fn;
    double* dinv = smem + L::o_dinv;
    double* wp = smem + L::o_wp;
    double* pc_s = smem + L::o_pc;
    double* Qi = smem + L::o_qi;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::o_bar);
    // run-time tail: constraint matrix ring + rho, rho.*g of the current stage
    const int ncmax = CON ? p.ncmax : 0;
    const bool sel = CON && p.sel_col != nullptr;   // selection-matrix constraints: no dense D at all
    const int DSTRIDE = sel ? 0 : even_up(ncmax * S);
    double* Dbuf = smem + L::DOUBLES;
    double* rho_s = Dbuf + DSTRIDE;
    double* rg_s = rho_s + ncmax;
    double* sval_s = rg_s + ncmax;
    double* dg_s = sval_s + ncmax;                  // selection mode: diag(D^T rho D) | D^T (rho o g) of a stage, scatter-added
                                                    // by the row threads (shared-memory atomics); two buffers: the rows of the
                                                    // NEXT stage are folded while this stage's factorisation runs
    int* scol_s = reinterpret_cast<int*>(dg_s + 4 * S);

    const size_t ws_len = (size_t)p.N * S + NX;
    const double* model_b = p.model + (size_t)b * p.N * D::REC;
    const double* ws_b = p.ws_prev ? p.ws_prev + (size_t)b * ws_len : nullptr;
    double* fac_b = p.fac + (size_t)b * p.N * D::FREC;
    double* aff_b = p.aff ? p.aff + (size_t)b * p.N * D::AREC : nullptr;
    const double* D_b = ncmax > 0 ? p.Dm + (size_t)b * p.d_total : nullptr;
    const size_t cbase = (size_t)b * p.nc_total;
    const double sigma = p.sigma;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    // segment terminal condition (lqr_kernel_parallel.hpp:51-67; lqr_kernel.hpp:79-91 for the last segment)
    for (int e = tid; e < NX * NX; e += T) {
        const int i = e % NX, j = e / NX;
        double Pv = 0.0, Fv = (i == j) ? 1.0 : 0.0;
        if (is_last) {
            Pv = p.HN[(size_t)b * NX * NX + e] + ((i == j) ? sigma : 0.0);
            Fv = 0.0;
        }
        PF[i + j * L::LDPF] = Pv;
        PF[NX + i + j * L::LDPF] = Fv;
        if constexpr (!L::C_REGS) Cn[e] = 0.0;
    }
    for (int i = tid; i < NX; i += T) {
        double pv = 0.0;
        if (is_last) pv = p.hN[(size_t)b * NX + i] - (ws_b ? sigma * ws_b[(size_t)p.N * S + i] : 0.0);
        pn[i] = pv;
        fn[i] = 0.0;
    }
    if (is_last && ncmax > 0 && p.ncs[p.N] > 0) {
        // terminal fold-in (lqr_kernel.hpp:82-87): P_N += D_N^T rho D_N ; p_N -= D_N^T (rho o (z - y/rho))
        const int ncN = p.ncs[p.N];
        const double* DN = D_b + p.doff[p.N];
        const size_t co = cbase + p.coff[p.N];
        group_sync<T>();
        if (sel) {   // D_N rows are scaled unit vectors: only the diagonal of P_N and p_N change
            for (int i = tid; i < NX; i += T) {
                double dgi = 0.0, dhi = 0.0;
                for (int r = 0; r < ncN; ++r)
                    if (p.sel_col[co + r] == i) {
                        const double v = p.sel_val[co + r], rr = p.rho[co + r];
                        dgi = fma(rr * v, v, dgi);
                        dhi = fma(v, rr * (p.zs[co + r] - p.inv_rho[co + r] * p.ys[co + r]), dhi);
                    }
                PF[i + i * L::LDPF] += dgi;
                pn[i] -= dhi;
            }
        } else {
            for (int e = tid; e < NX * NX; e += T) {
                const int i = e % NX, j = e / NX;
                double acc = 0.0;
                for (int r = 0; r < ncN; ++r) acc = fma(DN[r + i * ncN] * p.rho[co + r], DN[r + j * ncN], acc);
                PF[i + j * L::LDPF] += acc;
            }
            for (int i = tid; i < NX; i += T) {
                double acc = 0.0;
                for (int r = 0; r < ncN; ++r)
                    acc = fma(DN[r + i * ncN], p.rho[co + r] * (p.zs[co + r] - p.inv_rho[co + r] * p.ys[co + r]), acc);
                pn[i] -= acc;
            }
        }
    }
    if (sel)
        for (int i = tid; i < 4 * S; i += T) dg_s[i] = 0.0;
    group_sync<T>();
    auto issue_stage = [&](int kk, int bufi) {   // one elected thread: stage record (+ constraint matrix) of stage kk
        const int nck = (ncmax > 0 && !sel) ? p.ncs[kk] : 0;
        const uint32_t dbytes = (uint32_t)even_up(nck * S) * 8;
        mbar_expect_tx(&bar[bufi], D::REC * 8 + dbytes);
        bulk_g2s(rec + bufi * D::REC, model_b + (size_t)kk * D::REC, D::REC * 8, &bar[bufi]);
        if (nck > 0) bulk_g2s(Dbuf + bufi * DSTRIDE, D_b + p.doff[kk], dbytes, &bar[bufi]);
    };
    if (tid == 0 && LEN > 0) issue_stage(N1 - 1, 0);
    // The record holds [E c] (NX x (S+1), leading dimension NX): both big products read it in place.

    int bad = 0;
    constexpr int CT = (NX + 7) / 8;          // 8 x 8 tiles per side of C
    double creg[L::C_REGS ? CT : 1][L::C_REGS ? CT : 1][2];
#pragma unroll
    for (int a = 0; a < (L::C_REGS ? CT : 1); ++a)
#pragma unroll
        for (int c = 0; c < (L::C_REGS ? CT : 1); ++c) creg[a][c][0] = creg[a][c][1] = 0.0;
    // w_prev of the stage is fetched one stage ahead into registers: a load consumed right away would stall the group
    // for an L2 round trip at the top of every stage
    constexpr int NW = (S + T - 1) / T;
    double wreg[NW];
    auto fetch_w = [&](int kk) {
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            const int i = tid + q * T;
            wreg[q] = (ws_b && i < S) ? ws_b[(size_t)kk * S + i] : 0.0;
        }
    };
    if (LEN > 0) fetch_w(N1 - 1);
    // the constraint vectors of a stage (rho, z, y, 1/rho and, for selection-matrix constraints, the column and value of
    // every row) are fetched one stage ahead into registers as well: consumed right after their loads they cost the group a
    // DRAM round trip at the top of every stage (3,600 of 17,000 cycles per stage at nx30/nu10/nc44, profiles/r2_phase_clocks.txt)
    constexpr int CR = CON ? 2 : 1;          // rows per thread held in registers
    double pr_rho[CR], pr_z[CR], pr_ir[CR], pr_y[CR], pr_v[CR];
    int pr_c[CR], pr_n = 0;
    // (row count and offset of the stage AFTER the prefetched one are fetched yet another stage ahead: the vector loads
    // need them for their addresses, and an address that is itself still in flight stalls the in-order group)
    int nx_n = 0;
    long long nx_co = 0;
    auto fetch_con = [&](int kk) {   // vectors of stage kk (its count / offset are in nx_n / nx_co), then count / offset of kk-1
        pr_n = nx_n;
        const size_t co = cbase + nx_co;
#pragma unroll
        for (int q = 0; q < CR; ++q) {
            const int r = (tid + T - T / 2) % T + q * T;
            pr_c[q] = -1; pr_v[q] = 0.0;
            if (r < pr_n) {
                pr_rho[q] = p.rho[co + r]; pr_z[q] = p.zs[co + r]; pr_ir[q] = p.inv_rho[co + r]; pr_y[q] = p.ys[co + r];
                if (sel) { pr_c[q] = p.sel_col[co + r]; pr_v[q] = p.sel_val[co + r]; }
            }
        }
        if (kk - 1 >= N0) { nx_n = p.ncs[kk - 1]; nx_co = p.coff[kk - 1]; }
    };
    // fold of one stage's rows: g = z - y/rho ; keep rho and rho.*g (lqr_solver_parallel.hpp:134-137, lqr_kernel.hpp:110);
    // selection-matrix rows scatter-add diag(D^T rho D) and D^T (rho o g) into dgb | dgb + S.  It runs one stage AHEAD, right
    // after S3 of the previous stage, on the threads that have no right-hand side in S4 / S5 (rows are dealt out from the
    // middle of the group): at the top of the stage it cost 2,000 of 13,000 cycles at nx30/nu10/nc44.
    const int ctid = (tid + T - T / 2) % T;
    auto fold_stage = [&](int kk, double* dgb) {
        const int n = pr_n;
        if (n <= 0) return;
        auto row = [&](int r, double rr, double zz, double ir, double yy, int cj, double v) {
            rho_s[r] = rr;
            const double rg = rr * (zz - ir * yy);
            rg_s[r] = rg;
            if (sel && cj >= 0) {   // D_k rows are scaled unit vectors: D^T rho D is diagonal
                atomicAdd(&dgb[cj], rr * v * v);
                atomicAdd(&dgb[S + cj], v * rg);
            }
        };
#pragma unroll
        for (int q = 0; q < CR; ++q)   // rows fetched a stage ahead (registers)
            if (ctid + q * T < n) row(ctid + q * T, pr_rho[q], pr_z[q], pr_ir[q], pr_y[q], pr_c[q], pr_v[q]);
        if (n > CR * T) {              // (more than CR rows per thread: straight from global memory)
            const size_t co = cbase + p.coff[kk];
            for (int r = ctid + CR * T; r < n; r += T)
                row(r, p.rho[co + r], p.zs[co + r], p.inv_rho[co + r], p.ys[co + r], sel ? p.sel_col[co + r] : -1,
                    sel ? p.sel_val[co + r] : 0.0);
        }
    };
    int nck_next = 0;
    if (CON && LEN > 0) {
        nx_n = ncmax > 0 ? p.ncs[N1 - 1] : 0;
        nx_co = ncmax > 0 ? p.coff[N1 - 1] : 0;
        fetch_con(N1 - 1);
        if constexpr (PDPLQR_FOLD_AHEAD) {
            fold_stage(N1 - 1, dg_s);
            nck_next = pr_n;
            if (LEN > 1) fetch_con(N1 - 2);
        }
    }
    constexpr bool Z_BULK = (D::FREC % 2 == 0) && ((NU * D::NRHS) % 2 == 0) && ((NU * (NX + 1)) % 2 == 0);
    PHASE_DECL
    PHASE_START();
#pragma unroll 1
    for (int it = 0; it < LEN; ++it) {
        const int k = N1 - 1 - it;
        const int buf = 0;
        const double* R = rec;
        mbar_wait(&bar[0], it & 1);              // requested after S3 of the previous stage (or in the prologue)
        PHASE(0);
        double* dg_c = dg_s + (it & 1) * 2 * S;
        double* dh_c = dg_c + S;
        if constexpr (CON && !PDPLQR_FOLD_AHEAD) {   // (debug switch: fold at the top of the stage)
            nck_next = pr_n;
            fold_stage(k, dg_c);
            if (it + 1 < LEN) fetch_con(k - 1);
        }
        const int nck = CON ? nck_next : 0;      // (this stage's rows were folded while the previous stage factorised)
#pragma unroll
        for (int q = 0; q < NW; ++q)
            if (tid + q * T < S) wp[tid + q * T] = wreg[q];
        if (it + 1 < LEN) fetch_w(k - 1);
        PHASE(1);
        // S2: PFE = [P+; F+] * [E c]  (+ p+ on the last column of the P rows)
        {
            constexpr int MM = 2 * NX;
            constexpr Tile tl = pick_tile(MM, S + 1, T);
            // (kk is the storage position of a row of [E c]: the contraction runs in the record's row order)
            auto la = [&](int i, int kk) { return PF[i + D::eri(kk) * L::LDPF]; };
            auto lb = [&](int kk, int j) { return R[kk + D::wic(j) * NX]; };     // [E c](row at kk, j)
            auto epi = [&](int i, int j, double v) {
                if (j == S && i < NX) { pc_s[i] = v; v += pn[i]; }
                PFE[i + j * L::LDPE] = v;
            };
            if (pdp) gmm<MM, S + 1, NX, tl.tm, tl.tn, T>(tid, la, lb, epi);
            else {
                constexpr Tile t2 = pick_tile(NX, S + 1, T);
                gmm<NX, S + 1, NX, t2.tm, t2.tn, T>(tid, la, lb, epi);
            }
        }
        PHASE(2);
        group_sync<T>();
        PHASE(3);

        // S3: [M | g] = [H + sigma I | h - sigma w_prev] + E^T * PE        (update_problem_data fused in)
        {
            constexpr Tile tl = pick_tile(S, S + 1, T);
            auto la = [&](int i, int kk) { return R[kk + D::wi(i) * NX]; };      // E^T(i, row at kk)
            auto lb = [&](int kk, int j) { return PFE[D::eri(kk) + j * L::LDPE]; };
            auto epi = [&](int i, int j, double v) {
                double base;
                if (j < S) base = R[D::H_at(i, j)] + ((i == j) ? sigma : 0.0);
                else base = R[D::h_at(i)] - sigma * wp[i];
                if (sel && nck > 0) {   // selection-matrix fold-in (dg, dh were scatter-added at the top of the stage)
                    if (i == j) base += dg_c[i];
                    else if (j == S) base -= dh_c[i];
                }
                Ma[i + j * L::LDM] = base + v;
            };
            // only the tiles on / below the diagonal + the g column where the product runs on the tensor cores (every reader
            // of Ma below -- S4, S5, S6 -- stays in that part)
            constexpr bool SYM = PDPLQR_SYM_S3 && (long long)S * (S + 1) * NX >= PDPLQR_DMMA_MIN_MACS &&
                                 (sym_lower_tiles(S) + T / 32 - 1) / (T / 32) <= 10;
            if constexpr (SYM) {
                // H is symmetric in the record (pack_model_kernel mirrors the lower triangle, the one the reference's LLT
                // reads): the lane's pair H(i,j), H(i,j+1) is read as the contiguous H(j,i), H(j+1,i) with one 16-byte load --
                // conflict-free, where two 8-byte loads down a column pair cost 4 wavefronts each at S = 0 (mod 8)
                constexpr bool VEC = PDPLQR_S3_VEC && !D::WLAY && !D::H_ROT && (S % 2 == 0) && (D::REC_H % 2 == 0);
                auto epi2 = [&](int i, int j, double v0, double v1) {
                    double b0, b1 = 0.0;
                    const bool in1 = j + 1 <= S;
                    if (VEC && j + 1 < S) {
                        const double2 hv = *reinterpret_cast<const double2*>(R + D::REC_H + j + i * S);
                        b0 = hv.x; b1 = hv.y;
                    } else {
                        b0 = (j < S) ? R[D::H_at(i, j)] : R[D::h_at(i)] - sigma * wp[i];
                        if (in1) b1 = (j + 1 < S) ? R[D::H_at(i, j + 1)] : R[D::h_at(i)] - sigma * wp[i];
                    }
                    if (i == j) b0 += sigma;
                    if (i == j + 1 && j + 1 < S) b1 += sigma;
                    if (sel && nck > 0) {
                        if (i == j) b0 += dg_c[i];
                        else if (j == S) b0 -= dh_c[i];
                        if (i == j + 1 && j + 1 < S) b1 += dg_c[i];
                        else if (j + 1 == S) b1 -= dh_c[i];
                    }
                    Ma[i + j * L::LDM] = b0 + v0;
                    if (in1) Ma[i + (j + 1) * L::LDM] = b1 + v1;
                };
                dmma_sym_lower<S, T / 32>(tid >> 5, tid & 31, NX, la, lb, epi2);
            } else
                gmm<S, S + 1, NX, tl.tm, tl.tn, T>(tid, la, lb, epi);
            if (nck > 0 && !sel) {  // M += D^T diag(rho) D ; g -= D^T (rho o g_c)      (lqr_kernel.hpp:106-112)
                const double* Dk = Dbuf + buf * DSTRIDE;
                auto lda = [&](int i, int r) { return Dk[r + i * nck]; };
                auto ldb = [&](int r, int j) { return j < S ? rho_s[r] * Dk[r + j * nck] : -rg_s[r]; };
                auto epd = [&](int i, int j, double v) { Ma[i + j * L::LDM] += v; };
                gmm_rt<S, S + 1, tl.tm, tl.tn, T>(tid, nck, lda, ldb, epd);
            }
        }
        PHASE(4);
        if constexpr (Z_BULK) {          // the previous stage's factor record has left shared memory long ago: S5 may write Z
            if (tid == 0 && it > 0) bulk_wait_read<0>();   // (waited for at the end of S6 it held thread 0, and with it the
        }                                                  //  end-of-stage barrier, for the drain time of the store)
        group_sync<T>();
        if (tid == 0 && it + 1 < LEN) {  // the record (H, h) and D had their last readers in S3: fetch stage k-1 into
            fence_proxy_async();         // the same buffer; it lands while S4-S6 run and is waited for at the top of the next stage
            issue_stage(k - 1, 0);
        }
        if (sel && nck > 0)
            for (int i = tid; i < 2 * S; i += T) dg_c[i] = 0.0;   // ready for the scatter-add of the stage after the next
        if (CON && PDPLQR_FOLD_AHEAD && it + 1 < LEN) {    // the next stage's rows (fetched a stage ago), then the fetch for the stage after it
            fold_stage(k - 1, dg_s + ((it + 1) & 1) * 2 * S);
            nck_next = pr_n;
            if (it + 2 < LEN) fetch_con(k - 2);
        }

        PHASE(5);
        // S4: Quu = L D L^T (unit lower L).  NU <= 12: every solving thread factorises its own register copy (no
        //     barriers, no square roots); larger NU: cooperative Cholesky of the leading block of Ma in place.
        constexpr bool REG_CHOL = NU <= 12;
        double Lr[REG_CHOL ? NU : 1][REG_CHOL ? NU : 1];   // Lr[i][j], i > j: L(i,j); indexed through static_for only, so
                                                            // it is registers at every NU (with `#pragma unroll` loops the
                                                            // 10 x 10 factor of nx30/nu10 went to local memory)
        double dr[REG_CHOL ? NU : 1];                       // 1 / D(c)
        const int n1 = NX + 1, n2 = pdp ? NX : 0, n3 = aff_b ? NU : 0;
        if constexpr (REG_CHOL) {
            if (tid < n1 + n2 + n3) {
                double dd[NU];                              // D(q)
                static_for<0, NU>([&](auto jc) {
                    constexpr int j = jc;
                    static_for<j + 1, NU>([&](auto ic) {
                        constexpr int i = ic;
                        Lr[i][j] = Ma[i + j * L::LDM];
                    });
                });
                static_for<0, NU>([&](auto cc) {
                    constexpr int c = cc;
                    double vc[NU > 1 ? NU : 2];             // vc[q] = L(c,q) D(q): row c of L D (off the pivot chain for q < c-1)
                    static_for<0, c>([&](auto qc) { constexpr int q = qc; vc[q] = Lr[c][q] * dd[q]; });
                    double a = Ma[c + c * L::LDM];
                    static_for<0, c>([&](auto qc) { constexpr int q = qc; a = fma(-Lr[c][q], vc[q], a); });
                    double r = rcp_newton(a);              // the pivot test stays off the dependent chain
                    if (!(a > 0.0)) {                      // (rare) keep the sweep finite, report through the status word
                        if (!bad) bad = k + 1;
                        a = fabs(a) + 1e-300;
                        r = rcp_newton(a);
                    }
                    dr[c] = r;
                    dd[c] = a;
                    static_for<c + 1, NU>([&](auto ic) {
                        constexpr int i = ic;
                        double v = Lr[i][c];
                        static_for<0, c>([&](auto qc) { constexpr int q = qc; v = fma(-Lr[i][q], vc[q], v); });
                        Lr[i][c] = v * r;
                    });
                });
            }
        } else {
            const int info = group_chol<NU, T>(tid, Ma, L::LDM, dinv);
            if (info && !bad) bad = k + 1;
        }
        PHASE(6);
        // S5: one right-hand side per thread: z = -Quu^-1 r,  r in [Qux | Qu | (F+B)^T]  (unit vectors -> Quu^-1)
        {
            for (int q = tid; q < n1 + n2 + n3; q += T) {
                const int c = (q < n1 + n2) ? q : D::NRHS + (q - n1 - n2);   // >= NRHS: unit vectors -> Quu^-1
                // branch-free gather of the right-hand side (a per-element if / else chain here compiled into 4 x 4 divergent
                // paths with their shared-memory latencies exposed one after the other: ~500 of the stage's cycles)
                const double* src;
                int stride;
                if (c < NX) { src = Ma + (NU + c); stride = L::LDM; }                                // Qux(m,c) = Qxu(c,m)
                else if (c == NX) { src = Ma + S * L::LDM; stride = 1; }                             // Qu(m)
                else if (c < D::NRHS) { src = PFE + (NX + (c - NX - 1)); stride = L::LDPE; }         // (F+ B)(c', m)
                else { src = Ma; stride = 0; }                                                       // unit vector (below)
                double y[NU];
#pragma unroll
                for (int m = 0; m < NU; ++m) {
                    const double r = src[m * stride];
                    y[m] = (c < D::NRHS) ? r : ((m == c - D::NRHS) ? 1.0 : 0.0);
                }
                double z[NU];
                if constexpr (REG_CHOL) {      // L D L^T: forward (unit lower), scale, backward
                    static_for<1, NU>([&](auto mc) {
                        constexpr int m = mc;
                        double v = y[m];
                        static_for<0, m>([&](auto qc) { constexpr int qq = qc; v = fma(-Lr[m][qq], y[qq], v); });
                        y[m] = v;
                    });
                    static_for<0, NU>([&](auto mc) {
                        constexpr int m = NU - 1 - mc;
                        double v = y[m] * dr[m];
                        static_for<m + 1, NU>([&](auto qc) { constexpr int qq = qc; v = fma(-Lr[qq][m], z[qq], v); });
                        z[m] = v;
                    });
                } else {                       // Cholesky factor in Ma, reciprocal diagonal in dinv
#pragma unroll
                    for (int m = 0; m < NU; ++m) {
                        double v = y[m];
#pragma unroll
                        for (int qq = 0; qq < m; ++qq) v = fma(-Ma[m + qq * L::LDM], y[qq], v);
                        y[m] = v * dinv[m];
                    }
#pragma unroll
                    for (int m = NU - 1; m >= 0; --m) {
                        double v = y[m];
#pragma unroll
                        for (int qq = m + 1; qq < NU; ++qq) v = fma(-Ma[qq + m * L::LDM], z[qq], v);
                        z[m] = v * dinv[m];
                    }
                }
                if (c < D::NRHS) {
#pragma unroll
                    for (int m = 0; m < NU; ++m) Z[m + c * NU] = -z[m];
                } else {
#pragma unroll
                    for (int m = 0; m < NU; ++m) Qi[m + (c - D::NRHS) * NU] = z[m];
                }
            }
        }
        PHASE(7);
        if constexpr (Z_BULK) fence_proxy_async();   // Z is picked up by a bulk copy below
        group_sync<T>();
        if constexpr (Z_BULK) {   // factor record -> global, one TMA store; it drains while S6 runs
            if (tid == 0) {
                bulk_s2g(fac_b + (size_t)k * D::FREC, Z, (pdp ? NU * D::NRHS : NU * (NX + 1)) * 8);
                bulk_commit();
            }
        }

        PHASE(8);
        // S6: P = Qxx + Qxu K, p = Qx + Qxu d  |  C += (F+B)(-Gt)  |  [F f] = F+[A c] + (F+B)[K d] + [0 f+]
        {
            constexpr Tile tl = pick_tile(NX, NX + 1, T);
            auto la = [&](int i, int m) { return Ma[(NU + i) + m * L::LDM]; };      // Qxu(i,m)
            auto lb = [&](int m, int j) { return Z[m + j * NU]; };                   // [K d](m,j)
            auto epi = [&](int i, int j, double v) {
                if (j < NX) {   // the lower triangle is computed, the upper one mirrored: P stays exactly symmetric
                    if (i >= j) {
                        const double pv = Ma[(NU + i) + (NU + j) * L::LDM] + v;
                        PF[i + j * L::LDPF] = pv;
                        PF[j + i * L::LDPF] = pv;
                    }
                } else {
                    pn[i] = Ma[(NU + i) + S * L::LDM] + v;
                }
            };
            gmm<NX, NX + 1, NU, tl.tm, tl.tn, T>(tid, la, lb, epi);
            if (pdp) {
                if constexpr (L::C_REGS) {
                    if (tid < 32) {   // C += (F+B)(-Gt), accumulated in the tensor-core accumulator registers of warp 0
                        const int fr = tid >> 2, fq = tid & 3;
#pragma unroll
                        for (int kt = 0; kt < (NU + 3) / 4; ++kt) {
                            const int m = kt * 4 + fq;
                            double fa[CT], fb[CT];
#pragma unroll
                            for (int a = 0; a < CT; ++a) {
                                const bool in = m < NU && 8 * a + fr < NX;
                                fa[a] = in ? PFE[(NX + 8 * a + fr) + m * L::LDPE] : 0.0;            // (F+B)(i, m)
                                fb[a] = in ? -Z[m + (NX + 1 + 8 * a + fr) * NU] : 0.0;              // -Gt(m, j)
                            }
#pragma unroll
                            for (int a = 0; a < CT; ++a)
#pragma unroll
                                for (int c = 0; c < CT; ++c) dmma_m8n8k4(creg[a][c][0], creg[a][c][1], fa[a], fb[c]);
                        }
                    }
                } else {
                    constexpr Tile tc = pick_tile(NX, NX, T);
                    auto lga = [&](int i, int m) { return PFE[(NX + i) + m * L::LDPE]; };
                    auto lgb = [&](int m, int j) { return -Z[m + (NX + 1 + j) * NU]; };
                    auto epc = [&](int i, int j, double v) { Cn[i + j * NX] += v; };
                    gmm<NX, NX, NU, tc.tm, tc.tn, T>(tid, lga, lgb, epc);
                }
                auto lfa = [&](int i, int m) { return PFE[(NX + i) + m * L::LDPE]; };     // (F+ B)(i,m)
                auto lfb = [&](int m, int j) { return Z[m + j * NU]; };                    // [K d](m,j)
                auto epf = [&](int i, int j, double v) {
                    if (j < NX) PF[(NX + i) + j * L::LDPF] = PFE[(NX + i) + (NU + j) * L::LDPE] + v;   // F+A + (F+B)K
                    else fn[i] += PFE[(NX + i) + S * L::LDPE] + v;                                      // F+c + (F+B)d + f+
                };
                gmm<NX, NX + 1, NU, tl.tm, tl.tn, T>(tid, lfa, lfb, epf);
            }
            if constexpr (!Z_BULK) {   // factor record -> global (coalesced)
                double* fk = fac_b + (size_t)k * D::FREC;
                const int nz = pdp ? NU * D::NRHS : NU * (NX + 1);
                for (int e = tid; e < nz; e += T) fk[e] = Z[e];
            }   // (Z_BULK: the bulk store's read of Z is waited for just before Z is written again, after S3 of the next stage)
            if (aff_b) {  // what backward_without_factorization needs: Quu^-1, P+c, F+c, F+B
                double* ak = aff_b + (size_t)k * D::AREC;
                for (int e = tid; e < NU * NU; e += T) ak[D::AR_QI + e] = Qi[e];
                for (int i = tid; i < NX; i += T) {
                    ak[D::AR_PC + i] = pc_s[i];
                    if (pdp) ak[D::AR_FC + i] = PFE[(NX + i) + S * L::LDPE];
                }
                if (pdp)
                    for (int e = tid; e < NX * NU; e += T) ak[D::AR_FB + e] = PFE[(NX + e % NX) + (e / NX) * L::LDPE];
            }
        }
        PHASE(9);
        group_sync<T>();
        PHASE(10);
    }
    PHASE_PRINT("seg_backward", LEN);
    if constexpr (Z_BULK) {
        if (tid == 0 && LEN > 0) bulk_wait_read<0>();   // shared memory must outlive the last bulk store's read
    }

    // segment summary (lqr_solver_parallel.hpp:180-187): P, F, C, p, f at the segment entry
    double* sm = p.sum + ((size_t)b * p.S + seg) * D::SREC;
    for (int e = tid; e < NX * NX; e += T) {
        const int i = e % NX, j = e / NX;
        sm[D::SUM_P + e] = PF[i + j * L::LDPF];
        sm[D::SUM_F + e] = PF[NX + i + j * L::LDPF];
        if constexpr (!L::C_REGS) sm[D::SUM_C + e] = Cn[e];
    }
    if constexpr (L::C_REGS) {
        if (tid < 32) {   // accumulator fragment: lane -> row 8a + lane/4, columns 8c + 2 (lane%4) + {0, 1}
            const int fr = tid >> 2, fq = tid & 3;
#pragma unroll
            for (int a = 0; a < CT; ++a)
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    const int i = 8 * a + fr, j = 8 * c + 2 * fq;
                    if (i < NX && j < NX) sm[D::SUM_C + i + j * NX] = creg[a][c][0];
                    if (i < NX && j + 1 < NX) sm[D::SUM_C + i + (j + 1) * NX] = creg[a][c][1];
                }
        }
    }
    for (int i = tid; i < NX; i += T) {
        sm[D::SUM_p + i] = pn[i];
        sm[D::SUM_f + i] = fn[i];
    }
    if (tid == 0) p.status[(size_t)b * p.S + seg] = bad;   // one slot per (problem, segment): no clearing launch needed
}

// ------------------------------------------------------------------------------------------------
// Backward without factorisation (affine-only re-solve with the cached K, Quu^-1, P+c, F+c, F+B):
//     t = P+c + p+ ;  g = h~ + E^T t ;  d = -Quu^-1 g_u ;  p = g_x + K^T g_u ;  f = F+c + (F+B) d + f+
// Replaces LQRKernel::step_without_factorization (lqr_kernel.hpp:149-178), ParallelLQRKernel::
// step_without_factorization (lqr_kernel_parallel.hpp:138-168) and reduction_without_factorization
// (lqr_solver_parallel.hpp:190-211).  One warp (or, for nx + nu >= 32, one 128-thread CTA) per (problem, segment);
// only d (inside Z) and the segment summary's p, f change.
template <int NX, int NU>
struct AffSmem {
    using D = SegDims<NX, NU>;
    static constexpr int S = D::S;
    // of the stage record only [E | c] and h are needed (H is 56 % of the record at nx=30): two bulk copies
    static constexpr int H_OFF = D::REC_h & ~1;              // 16-byte aligned start of the h chunk inside the record
    static constexpr int H_LEN = (D::REC - H_OFF);           // doubles copied (h plus at most one neighbour)
    static constexpr int RLITE = D::REC_EC + H_LEN;          // doubles per ring slot
    // Ring slots of the factor record and of the affine cache: a launch whose groups are all last / only segments (one segment
    // per problem: the ADMM batch) needs only [K | d] and [Quu^-1 | P+c] of them -- 29 instead of 39 KB per group at nx30/nu10,
    // 7 instead of 5 resident groups per SM ("lite" layout, chosen by the launcher, recomputed by the kernel from p.S)
    static constexpr int FSL_LITE = even_up(NU * (NX + 1)), ASL_LITE = even_up(D::AR_FC);
    struct Layout { int fsl, asl, o_fac, o_aff, o_t, o_g, o_d, o_pn, o_fn, o_bar, doubles; };
    PDPLQR_HD static constexpr Layout layout(bool lite) {
        Layout l{};
        l.fsl = lite ? FSL_LITE : D::FREC;
        l.asl = lite ? ASL_LITE : D::AREC;
        l.o_fac = 2 * RLITE;                 // rec: 2 x RLITE (TMA) at 0
        l.o_aff = l.o_fac + 2 * l.fsl;       // 2 x fsl  (TMA)
        l.o_t = l.o_aff + 2 * l.asl;         // 2 x asl  (TMA); t (NX)
        l.o_g = l.o_t + NX;                  // g (S)
        l.o_d = l.o_g + S;                   // d (NU)
        l.o_pn = l.o_d + NU;
        l.o_fn = l.o_pn + NX;
        l.o_bar = even_up(l.o_fn + NX);
        l.doubles = even_up(l.o_bar + 2);
        return l;
    }
    static size_t bytes(int ncmax, bool sel = false, bool lite = false) {
        const size_t base = (size_t)layout(lite).doubles * 8;
        if (ncmax <= 0) return base;
        return base + (size_t)((sel ? 0 : 2 * even_up(ncmax * S)) + ncmax + (sel ? ncmax + even_up(ncmax) / 2 + 1 + 4 * S : 0)) * 8;
    }
};

template <int NX, int NU, int T>
__global__ void __launch_bounds__(T) seg_affine_kernel(SegParams p) {
    using D = SegDims<NX, NU>;
    using L = AffSmem<NX, NU>;
    constexpr int S = D::S;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int gidx = blockIdx.x;
    const int b = gidx / p.S, seg = gidx % p.S;
    const int N0 = seg_first(p, seg), LEN = seg_first(p, seg + 1) - N0, N1 = N0 + LEN;
    const bool is_last = (seg == p.S - 1) && !p.interior;
    const bool pdp = !is_last;

    const typename L::Layout lay = L::layout(p.S == 1 && !p.interior);   // (the launcher sized the allocation the same way)
    const int FSL = lay.fsl, ASL = lay.asl;
    double* rec = smem;
    double* fac = smem + lay.o_fac;
    double* aff = smem + lay.o_aff;
    double* gv = smem + lay.o_g;
    double* dv = smem + lay.o_d;
    double* pn = smem + lay.o_pn;
    double* fn = smem + lay.o_fn;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + lay.o_bar);
    const int ncmax = p.ncmax;
    const bool sel = p.sel_col != nullptr;
    const int DSTRIDE = sel ? 0 : even_up(ncmax * S);
    double* Dbuf = smem + lay.doubles;
    double* rg_s = Dbuf + 2 * DSTRIDE;
    double* sval_s = rg_s + ncmax;
    double* dh_s = sval_s + ncmax + S;              // (same tail layout as the factorising kernel: dg unused here)

    const size_t ws_len = (size_t)p.N * S + NX;
    const double* model_b = p.model + (size_t)b * p.N * D::REC;
    const double* ws_b = p.ws_prev ? p.ws_prev + (size_t)b * ws_len : nullptr;
    double* fac_b = p.fac + (size_t)b * p.N * D::FREC;
    const double* aff_b = p.aff + (size_t)b * p.N * D::AREC;
    const double* D_b = ncmax > 0 ? p.Dm + (size_t)b * p.d_total : nullptr;
    const size_t cbase = (size_t)b * p.nc_total;
    const double sigma = p.sigma;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    // terminal: p_N = h_N - sigma w_N - D_N^T (rho o g)   (lqr_kernel.hpp:93-101), zero for interior interfaces
    for (int i = tid; i < NX; i += T) {
        double pv = 0.0;
        if (is_last) {
            pv = p.hN[(size_t)b * NX + i] - (ws_b ? sigma * ws_b[(size_t)p.N * S + i] : 0.0);
            if (ncmax > 0 && p.ncs[p.N] > 0) {
                const int ncN = p.ncs[p.N];
                const double* DN = D_b + p.doff[p.N];
                const size_t co = cbase + p.coff[p.N];
                double acc = 0.0;
                for (int r = 0; r < ncN; ++r) {
                    const double dv = sel ? (p.sel_col[co + r] == i ? p.sel_val[co + r] : 0.0) : DN[r + i * ncN];
                    acc = fma(dv, p.rho[co + r] * (p.zs[co + r] - p.inv_rho[co + r] * p.ys[co + r]), acc);
                }
                pv -= acc;
            }
        }
        pn[i] = pv;
        fn[i] = 0.0;
    }
    if (sel)
        for (int i = tid; i < S; i += T) dh_s[i] = 0.0;
    group_sync<T>();
    auto issue_stage = [&](int kk, int bufi) {
        const int nck = (ncmax > 0 && !sel) ? p.ncs[kk] : 0;
        const uint32_t dbytes = (uint32_t)even_up(nck * S) * 8;
        // last / only segments need just [K | d] of the factor record and [Quu^-1 | P+c] of the affine cache
        const uint32_t fdoubles = pdp ? D::FREC : even_up(NU * (NX + 1));
        const uint32_t adoubles = pdp ? D::AREC : even_up(D::AR_FC);
        mbar_expect_tx(&bar[bufi], (L::RLITE + fdoubles + adoubles) * 8 + dbytes);
        bulk_g2s(rec + bufi * L::RLITE, model_b + (size_t)kk * D::REC, D::REC_EC * 8, &bar[bufi]);
        bulk_g2s(rec + bufi * L::RLITE + D::REC_EC, model_b + (size_t)kk * D::REC + L::H_OFF, L::H_LEN * 8, &bar[bufi]);
        bulk_g2s(fac + bufi * FSL, fac_b + (size_t)kk * D::FREC, fdoubles * 8, &bar[bufi]);
        bulk_g2s(aff + bufi * ASL, aff_b + (size_t)kk * D::AREC, adoubles * 8, &bar[bufi]);
        if (nck > 0) bulk_g2s(Dbuf + bufi * DSTRIDE, D_b + p.doff[kk], dbytes, &bar[bufi]);
    };
    if (tid == 0 && LEN > 0) issue_stage(N1 - 1, 0);
    // w_prev and the constraint vectors of a stage are fetched one stage ahead into registers (consumed right after their
    // loads they cost the group a DRAM round trip per stage -- this sweep is the common ADMM iteration)
    constexpr int NWA = (S + T - 1) / T, CR = 2;
    double wnext[NWA], pr_rho[CR], pr_z[CR], pr_ir[CR], pr_y[CR], pr_v[CR];
    int pr_c[CR], pr_n = 0;
    auto fetch_wa = [&](int kk) {
#pragma unroll
        for (int r = 0; r < NWA; ++r) {
            const int i = tid + r * T;
            wnext[r] = (i < S && ws_b) ? ws_b[(size_t)kk * S + i] : 0.0;
        }
    };
    // constraint rows are dealt out from the middle of the group: with 128 threads the rows of the next stage are then folded by
    // warps 2-3 while warp 0 forms d and warp 1 forms p
    const int ctid = (T >= 128) ? (tid + T - T / 2) % T : tid;
    int nx_n = 0;          // row count / offset of the next stage to prefetch (fetched one more stage ahead: the vector
    long long nx_co = 0;   // loads need them for their addresses)
    auto fetch_con = [&](int kk) {
        pr_n = nx_n;
        const size_t co = cbase + nx_co;
#pragma unroll
        for (int q = 0; q < CR; ++q) {
            const int r = ctid + q * T;
            pr_c[q] = -1; pr_v[q] = 0.0;
            if (r < pr_n) {
                pr_rho[q] = p.rho[co + r]; pr_z[q] = p.zs[co + r]; pr_ir[q] = p.inv_rho[co + r]; pr_y[q] = p.ys[co + r];
                if (sel) { pr_c[q] = p.sel_col[co + r]; pr_v[q] = p.sel_val[co + r]; }
            }
        }
        if (kk - 1 >= N0) { nx_n = p.ncs[kk - 1]; nx_co = p.coff[kk - 1]; }
    };
    // rows of one stage (vectors in the prefetch registers): rho o g -> rg_s, selection rows scatter-add D^T (rho o g) into dh_s.
    // Done one stage AHEAD, in the d / p phase of the previous stage (after its g phase has consumed and cleared rg_s / dh_s),
    // so that no barrier of its own separates it from the g phase that reads the result.
    auto fold_rows = [&](int kk) {
        const int n = pr_n;
        if (n <= 0) return;
        auto row = [&](int r, double rr, double zz, double ir, double yy, int cj, double v) {
            const double rg = rr * (zz - ir * yy);
            rg_s[r] = rg;
            if (sel && cj >= 0) atomicAdd(&dh_s[cj], v * rg);
        };
#pragma unroll
        for (int q = 0; q < CR; ++q)   // rows fetched a stage ahead (registers)
            if (ctid + q * T < n) row(ctid + q * T, pr_rho[q], pr_z[q], pr_ir[q], pr_y[q], pr_c[q], pr_v[q]);
        if (n > CR * T) {
            const size_t co = cbase + p.coff[kk];
            for (int r = ctid + CR * T; r < n; r += T)
                row(r, p.rho[co + r], p.zs[co + r], p.inv_rho[co + r], p.ys[co + r], sel ? p.sel_col[co + r] : -1,
                    sel ? p.sel_val[co + r] : 0.0);
        }
    };
    int nck_cur = 0;
    if (LEN > 0) {
        fetch_wa(N1 - 1);
        if (ncmax > 0) {
            nx_n = p.ncs[N1 - 1];
            nx_co = p.coff[N1 - 1];
            fetch_con(N1 - 1);
            fold_rows(N1 - 1);
            nck_cur = pr_n;
            if (LEN > 1) fetch_con(N1 - 2);
        }
    }
    group_sync<T>();
#pragma unroll 1
    for (int it = 0; it < LEN; ++it) {
        const int k = N1 - 1 - it;
        const int buf = it & 1;
        const double* R = rec + buf * L::RLITE;
        const double* hrec = R + D::REC_EC + (D::REC_h - L::H_OFF);   // h of this stage
        const double* Zk = fac + buf * FSL;
        const double* Ak = aff + buf * ASL;
        if (tid == 0 && it + 1 < LEN) {
            fence_proxy_async();
            issue_stage(k - 1, buf ^ 1);
        }
        const int nck = nck_cur;
        double wpv[NWA];
#pragma unroll
        for (int r = 0; r < NWA; ++r) wpv[r] = wnext[r];
        if (it + 1 < LEN) fetch_wa(k - 1);   // next stage's w_prev: in flight while this stage computes
        mbar_wait(&bar[buf], (it >> 1) & 1);
        // The stage is a chain of dependent mat-vecs between group barriers, and that chain -- not the 17 GB the sweep reads -- is
        // what the (problem, segment) groups resident on an SM spend their time in (ncu: 5.2 TB/s, half of all stall samples at
        // the barriers, a 30-term FMA chain on 40 of 128 threads).  So: t = P+c + p+ is formed inside the product (no barrier of
        // its own), every dot product runs on four independent accumulators, and d and p -- which both need only g -- share
        // one phase on different warps.
        // g = h - sigma w - D^T (rho o g_c) + E^T t ,  t = P+c + p+
#pragma unroll
        for (int r = 0; r < (S + T - 1) / T; ++r) {
            const int i = tid + r * T;
            if (i < S) {
                double acc = hrec[D::wi(i)] - sigma * wpv[r];
                if (nck > 0 && sel) {
                    acc -= dh_s[i];
                    dh_s[i] = 0.0;          // ready for the next stage's scatter-add (same thread, next use after a sync)
                } else if (nck > 0) {
                    const double* Dk = Dbuf + buf * DSTRIDE;
                    for (int q = 0; q < nck; ++q) acc = fma(-Dk[q + i * nck], rg_s[q], acc);
                }
                double a1 = 0.0, a2 = 0.0, a3 = 0.0;
                const double* Ei = R + D::wi(i) * NX;
#pragma unroll
                for (int q = 0; q + 3 < NX; q += 4) {
                    acc = fma(Ei[q], Ak[D::AR_PC + D::eri(q)] + pn[D::eri(q)], acc);
                    a1 = fma(Ei[q + 1], Ak[D::AR_PC + D::eri(q + 1)] + pn[D::eri(q + 1)], a1);
                    a2 = fma(Ei[q + 2], Ak[D::AR_PC + D::eri(q + 2)] + pn[D::eri(q + 2)], a2);
                    a3 = fma(Ei[q + 3], Ak[D::AR_PC + D::eri(q + 3)] + pn[D::eri(q + 3)], a3);
                }
#pragma unroll
                for (int q = NX & ~3; q < NX; ++q) acc = fma(Ei[q], Ak[D::AR_PC + D::eri(q)] + pn[D::eri(q)], acc);
                gv[i] = (acc + a1) + (a2 + a3);
            }
        }
        group_sync<T>();
        if (ncmax > 0 && it + 1 < LEN) {   // the next stage's rows, then the fetch for the stage after it
            fold_rows(k - 1);
            nck_cur = pr_n;
            if (it + 2 < LEN) fetch_con(k - 2);
        }
        // d = -Quu^-1 g_u   |   p = g_x + K^T g_u          (the latter one warp further on when the group has one)
        for (int m = tid; m < NU; m += T) {
            double acc = 0.0, a1 = 0.0;
#pragma unroll
            for (int q = 0; q + 1 < NU; q += 2) {
                acc = fma(-Ak[D::AR_QI + m + q * NU], gv[q], acc);
                a1 = fma(-Ak[D::AR_QI + m + (q + 1) * NU], gv[q + 1], a1);
            }
            if constexpr (NU & 1) acc = fma(-Ak[D::AR_QI + m + (NU - 1) * NU], gv[NU - 1], acc);
            acc += a1;
            dv[m] = acc;
            fac_b[(size_t)k * D::FREC + NU * NX + m] = acc;   // the d slot of Z = [K | d | Gt]
        }
        for (int i = ((T >= 64 && PDPLQR_AFF_PSHIFT) ? (tid + T - 32) % T : tid); i < NX; i += T) {   // (p+ had its last readers in the g phase)
            double acc = gv[NU + i], a1 = 0.0;
#pragma unroll
            for (int m = 0; m + 1 < NU; m += 2) {
                acc = fma(Zk[m + i * NU], gv[m], acc);
                a1 = fma(Zk[m + 1 + i * NU], gv[m + 1], a1);
            }
            if constexpr (NU & 1) acc = fma(Zk[NU - 1 + i * NU], gv[NU - 1], acc);
            pn[i] = acc + a1;
        }
        group_sync<T>();
        if (pdp) {           // f = F+c + (F+B) d + f+
            for (int i = tid; i < NX; i += T) {
                double af = Ak[D::AR_FC + i] + fn[i];
#pragma unroll
                for (int m = 0; m < NU; ++m) af = fma(Ak[D::AR_FB + i + m * NX], dv[m], af);
                fn[i] = af;
            }
            group_sync<T>();
        }
    }
    double* sm = p.sum + ((size_t)b * p.S + seg) * D::SREC;
    for (int i = tid; i < NX; i += T) {
        sm[D::SUM_p + i] = pn[i];
        sm[D::SUM_f + i] = fn[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Forward rollout of one segment: u_k = K x_k + d + Gt uhat ; x_{k+1} = c + A x_k + B u_k.
template <int NX, int NU>
struct FwdSmem {
    using D = SegDims<NX, NU>;
    static constexpr int o_rec = 0;                       // 2 x REC_EC
    static constexpr int o_fac = o_rec + 2 * D::REC_EC;   // 2 x FREC
    static constexpr int o_x = o_fac + 2 * D::FREC;
    static constexpr int o_u = o_x + NX;
    static constexpr int o_uh = o_u + NU;
    static constexpr int o_bar = even_up(o_uh + NX);
    static constexpr int DOUBLES = o_bar + 2;
    static constexpr size_t BYTES = (size_t)DOUBLES * 8;
};

template <int NX, int NU, int T>
__global__ void __launch_bounds__(T) seg_forward_kernel(SegParams p) {
    using D = SegDims<NX, NU>;
    using L = FwdSmem<NX, NU>;
    constexpr int S = D::S;
    extern __shared__ __align__(16) double smem[];
    pdl_wait();      // the predecessor kernel of the solve chain has completed (no-op without the launch attribute)
    pdl_trigger();   // the next kernel of the chain may be scheduled from here on
    const int tid = threadIdx.x;
    const int g = blockIdx.x;
    const int b = g / p.S, seg = g % p.S;
    const int N0 = seg_first(p, seg), LEN = seg_first(p, seg + 1) - N0, N1 = N0 + LEN;
    const bool is_last = (seg == p.S - 1) && !p.interior;

    double* rec = smem + L::o_rec;
    double* fac = smem + L::o_fac;
    double* xs = smem + L::o_x;
    double* us = smem + L::o_u;
    double* uh = smem + L::o_uh;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::o_bar);

    const size_t ws_len = (size_t)p.N * S + NX;
    const double* model_b = p.model + (size_t)b * p.N * D::REC;
    const double* fac_b = p.fac + (size_t)b * p.N * D::FREC;
    double* ws_b = p.ws_out + (size_t)b * ws_len;
    const uint32_t FD = is_last ? even_up(NU * (NX + 1)) : D::FREC;   // last segment: only [K | d] is needed
    const uint32_t TX = (D::REC_EC + FD) * 8;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < NX; i += T) {
        xs[i] = p.xhat[((size_t)b * p.S + seg) * NX + i];
        uh[i] = is_last ? 0.0 : p.uhat[((size_t)b * p.S + seg) * NX + i];
    }
    group_sync<T>();
    if (tid == 0 && LEN > 0) {
        mbar_expect_tx(&bar[0], TX);
        bulk_g2s(rec, model_b + (size_t)N0 * D::REC, D::REC_EC * 8, &bar[0]);
        bulk_g2s(fac, fac_b + (size_t)N0 * D::FREC, FD * 8, &bar[0]);
    }
#pragma unroll 1
    for (int it = 0; it < LEN; ++it) {
        const int k = N0 + it;
        const int buf = it & 1;
        const double* R = rec + buf * D::REC_EC;
        const double* Zk = fac + buf * D::FREC;
        if (tid == 0 && it + 1 < LEN) {
            fence_proxy_async();
            mbar_expect_tx(&bar[buf ^ 1], TX);
            bulk_g2s(rec + (buf ^ 1) * D::REC_EC, model_b + (size_t)(k + 1) * D::REC, D::REC_EC * 8, &bar[buf ^ 1]);
            bulk_g2s(fac + (buf ^ 1) * D::FREC, fac_b + (size_t)(k + 1) * D::FREC, FD * 8, &bar[buf ^ 1]);
        }
        mbar_wait(&bar[buf], (it >> 1) & 1);
        // u = K x + d (+ Gt uhat)
        for (int i = tid; i < NU; i += T) {   // (independent accumulators: the rollout is a chain of dependent mat-vecs)
            double acc = Zk[NU * NX + i], a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int j = 0; j + 1 < NX; j += 2) {
                acc = fma(Zk[i + j * NU], xs[j], acc);
                a1 = fma(Zk[i + (j + 1) * NU], xs[j + 1], a1);
            }
            if constexpr (NX & 1) acc = fma(Zk[i + (NX - 1) * NU], xs[NX - 1], acc);
            if (!is_last) {
#pragma unroll
                for (int j = 0; j + 1 < NX; j += 2) {
                    a2 = fma(Zk[NU * (NX + 1) + i + j * NU], uh[j], a2);
                    a3 = fma(Zk[NU * (NX + 1) + i + (j + 1) * NU], uh[j + 1], a3);
                }
                if constexpr (NX & 1) a2 = fma(Zk[NU * (NX + 1) + i + (NX - 1) * NU], uh[NX - 1], a2);
            }
            acc = (acc + a1) + (a2 + a3);
            us[i] = acc;
            ws_b[(size_t)k * S + i] = acc;
        }
        for (int i = tid; i < NX; i += T) ws_b[(size_t)k * S + NU + i] = xs[i];
        group_sync<T>();
        // x+ = c + B u + A x
        double xn[(NX + T - 1) / T];
#pragma unroll
        for (int r = 0; r < (NX + T - 1) / T; ++r) {
            const int i = tid + r * T;
            if (i < NX) {
                double acc = R[D::REC_C + D::er(i)], a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int j = 0; j + 1 < NU; j += 2) {
                    acc = fma(R[D::er(i) + D::wi(j) * NX], us[j], acc);
                    a1 = fma(R[D::er(i) + D::wi(j + 1) * NX], us[j + 1], a1);
                }
                if constexpr (NU & 1) acc = fma(R[D::er(i) + D::wi(NU - 1) * NX], us[NU - 1], acc);
#pragma unroll
                for (int j = 0; j + 1 < NX; j += 2) {
                    a2 = fma(R[D::er(i) + D::wi(NU + j) * NX], xs[j], a2);
                    a3 = fma(R[D::er(i) + D::wi(NU + j + 1) * NX], xs[j + 1], a3);
                }
                if constexpr (NX & 1) a2 = fma(R[D::er(i) + D::wi(NU + NX - 1) * NX], xs[NX - 1], a2);
                xn[r] = (acc + a1) + (a2 + a3);
            }
        }
        group_sync<T>();
#pragma unroll
        for (int r = 0; r < (NX + T - 1) / T; ++r) {
            const int i = tid + r * T;
            if (i < NX) xs[i] = xn[r];
        }
        group_sync<T>();
    }
    // the last segment owns x_N; other segments leave their exit state to the next segment's entry
    // (lqr_solver_parallel.hpp:224,235-236)
    if (is_last)
        for (int i = tid; i < NX; i += T) ws_b[(size_t)p.N * S + i] = xs[i];
}

}  // namespace pdplqr
