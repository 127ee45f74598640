// Specialised factorising backward sweep for the quadrotor class (nx % 4 == 0, nu % 4 == 0, no constraints, no
// affine cache): same mathematics and data flow as seg_backward_kernel (seg_kernels.cuh), but every small product is
// laid out "k-major" -- both operands are contiguous along their OUTPUT index (A[i + k*lda], B^T[j + k*ldb]) -- so a
// thread's TS x TS register tile is fed by 128-bit shared loads and the contraction loop is fully unrolled.  The
// generic kernel spends ~15 % of its instructions in FMAs (ncu source page, profiles/); this one roughly doubles that.
//
// Shared-memory operands (doubles, all leading dimensions even, all blocks 16-byte aligned):
//   PF  [2NX x NX]     rows 0..NX-1 = P+, NX.. = F+                      A operand of S2
//   ET  [SP  x NX]     ET[j + k*SP] = [E c](k, j), rows > S are zero      B^T operand of S2, A operand of S3
//   PEt [SP  x NX]     PEt[j + i*SP] = (P+[E c] + [0 p+])(i, j)           B^T operand of S3   (written transposed by S2)
//   FE  [NX  x (S+1)]  F+ [E c]                                           A operand of S6 (its first NU columns = F+B)
//   YT  [LDY x NU]     rows [Yx (NX) | Yg (NX) | yu]                      both operands of S6
//   ZT  [LDZ x NU]     ZT[j + m*LDZ] = [K d](m, j)                        B^T operand of S6
#pragma once
#include "common.cuh"
#include "seg_kernels.cuh"

namespace pdplqr {

constexpr int round_up4(int n) { return (n + 3) & ~3; }

template <int NX, int NU>
struct KmSmem {
    using D = SegDims<NX, NU>;
    static constexpr int S = D::S;
    static constexpr int SP = round_up4(S + 1);
    static constexpr int LDPF = 2 * NX;
    static constexpr int LDM = S;
    static constexpr int LDY = round_up4(2 * NX + 1);
    static constexpr int LDZ = round_up4(NX + 1);
    static constexpr int o_rec = 0;
    static constexpr int o_Z = o_rec + D::REC;
    static constexpr int o_ZT = o_Z + D::FREC;
    static constexpr int o_ET = o_ZT + LDZ * NU;
    static constexpr int o_PF = o_ET + SP * NX;
    static constexpr int o_PEt = o_PF + LDPF * NX;
    static constexpr int o_FE = o_PEt + SP * NX;
    static constexpr int o_Ma = even_up(o_FE + NX * (S + 1));
    static constexpr int o_YT = o_Ma + LDM * (S + 1);
    static constexpr int o_Cn = even_up(o_YT + LDY * NU);
    static constexpr int o_pn = o_Cn + NX * NX;
    static constexpr int o_fn = o_pn + NX;
    static constexpr int o_wp = o_fn + NX;
    static constexpr int o_bar = even_up(o_wp + S);
    static constexpr int DOUBLES = even_up(o_bar + 2);
    static constexpr size_t BYTES = (size_t)DOUBLES * 8;
    static constexpr bool ELIGIBLE = (NX % 4 == 0) && (NU % 4 == 0) && (NU <= 8) && (S <= 32);
};

// acc(r, c) += sum_k A[r + k*lda] * B[c + k*ldb]   (TS in {2, 4}; A, B 16-byte aligned, lda / ldb even)
template <int TS, int K>
PDPLQR_DEVINL void km_tile(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                           double (&acc)[TS][TS]) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a[TS], b[TS];
#pragma unroll
        for (int h = 0; h < TS / 2; ++h) {
            const double2 va = *reinterpret_cast<const double2*>(A + k * lda + 2 * h);
            const double2 vb = *reinterpret_cast<const double2*>(B + k * ldb + 2 * h);
            a[2 * h] = va.x; a[2 * h + 1] = va.y;
            b[2 * h] = vb.x; b[2 * h + 1] = vb.y;
        }
#pragma unroll
        for (int r = 0; r < TS; ++r)
#pragma unroll
            for (int c = 0; c < TS; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
    }
}

template <int NX, int NU, int T>
__global__ void __launch_bounds__(T) seg_backward_km_kernel(SegParams p) {
    using D = SegDims<NX, NU>;
    using L = KmSmem<NX, NU>;
    static_assert(L::ELIGIBLE, "seg_backward_km_kernel: nx, nu must be multiples of 4 with nx + nu <= 32");
    constexpr int S = D::S, SP = L::SP;
    constexpr int TS = (T == 32) ? 4 : 2;      // register tile edge
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int g = blockIdx.x;
    const int b = g / p.S, seg = g % p.S;
    const int N0 = seg_first(p, seg), LEN = seg_first(p, seg + 1) - N0, N1 = N0 + LEN;
    const bool is_last = (seg == p.S - 1) && !p.interior;
    const bool pdp = !is_last;

    double* rec = smem + L::o_rec;
    double* Z = smem + L::o_Z;
    double* ZT = smem + L::o_ZT;
    double* ET = smem + L::o_ET;
    double* PF = smem + L::o_PF;
    double* PEt = smem + L::o_PEt;
    double* FE = smem + L::o_FE;
    double* Ma = smem + L::o_Ma;
    double* YT = smem + L::o_YT;
    double* Cn = smem + L::o_Cn;
    double* pn = smem + L::o_pn;
    double* fn = smem + L::o_fn;
    double* wp = smem + L::o_wp;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::o_bar);

    const size_t ws_len = (size_t)p.N * S + NX;
    const double* model_b = p.model + (size_t)b * p.N * D::REC;
    const double* ws_b = p.ws_prev ? p.ws_prev + (size_t)b * ws_len : nullptr;
    double* fac_b = p.fac + (size_t)b * p.N * D::FREC;
    const double sigma = p.sigma;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
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
        Cn[e] = 0.0;
    }
    for (int i = tid; i < NX; i += T) {
        double pv = 0.0;
        if (is_last) pv = p.hN[(size_t)b * NX + i] - (ws_b ? sigma * ws_b[(size_t)p.N * S + i] : 0.0);
        pn[i] = pv;
        fn[i] = 0.0;
    }
    for (int e = tid; e < SP * NX; e += T) { ET[e] = 0.0; PEt[e] = 0.0; }   // padding rows stay zero
    for (int e = tid; e < L::LDY * NU; e += T) YT[e] = 0.0;
    for (int e = tid; e < L::LDZ * NU; e += T) ZT[e] = 0.0;
    group_sync<T>();
    auto issue_stage = [&](int kk) {
        mbar_expect_tx(&bar[0], D::REC * 8);
        bulk_g2s(rec, model_b + (size_t)kk * D::REC, D::REC * 8, &bar[0]);
    };
    if (tid == 0 && LEN > 0) issue_stage(N1 - 1);
    auto transpose_stage = [&]() {   // ET(j, k') = [E c](k', j)
        for (int e = tid; e < NX * (S + 1); e += T) {
            const int kk = e % NX, j = e / NX;
            ET[j + kk * SP] = rec[e];
        }
    };
    if (LEN > 0) {
        mbar_wait(&bar[0], 0);
        transpose_stage();
    }
    group_sync<T>();

    int bad = 0;
#pragma unroll 1
    for (int it = 0; it < LEN; ++it) {
        const int k = N1 - 1 - it;
        if (tid < S) wp[tid] = ws_b ? ws_b[(size_t)k * S + tid] : 0.0;

        // S2: [PEt ; FE] = [P+ ; F+] [E c]   (+ p+ on column S of the P rows), TS x TS register tiles
        {
            const int mrows = pdp ? 2 * NX : NX;
            const int MT = mrows / TS;
            constexpr int NT = SP / TS;
            for (int t = tid; t < MT * NT; t += T) {
                const int ti = t % MT, tj = t / MT;
                const int i0 = ti * TS, j0 = tj * TS;
                double acc[TS][TS];
#pragma unroll
                for (int r = 0; r < TS; ++r)
#pragma unroll
                    for (int c = 0; c < TS; ++c) acc[r][c] = 0.0;
                km_tile<TS, NX>(PF + i0, L::LDPF, ET + j0, SP, acc);
                if (i0 < NX) {   // P rows -> PEt (transposed), p+ added on column S
#pragma unroll
                    for (int r = 0; r < TS; ++r) {
                        const int i = i0 + r;
#pragma unroll
                        for (int c = 0; c < TS; ++c)
                            if (j0 + c == S) acc[r][c] += pn[i];
#pragma unroll
                        for (int h = 0; h < TS / 2; ++h)
                            *reinterpret_cast<double2*>(PEt + j0 + 2 * h + i * SP) = make_double2(acc[r][2 * h], acc[r][2 * h + 1]);
                    }
                } else {         // F rows -> FE (column-major NX x (S+1))
                    const int i = i0 - NX;
#pragma unroll
                    for (int c = 0; c < TS; ++c) {
                        if (j0 + c <= S) {
#pragma unroll
                            for (int h = 0; h < TS / 2; ++h)
                                *reinterpret_cast<double2*>(FE + i + 2 * h + (j0 + c) * NX) = make_double2(acc[2 * h][c], acc[2 * h + 1][c]);
                        }
                    }
                }
            }
        }
        group_sync<T>();

        // S3: [M | g] = [H + sigma I | h - sigma w_prev] + E^T PE : lower-triangular blocks of M and the last column
        {
            constexpr int NB = S / TS;                       // blocks per side
            constexpr int NTRI = NB * (NB + 1) / 2;
            for (int t = tid; t < NTRI + S; t += T) {
                if (t < NTRI) {
                    // t -> (bi, bj), bi >= bj, row by row of the lower triangle
                    int bi = 0, acc_t = 0;
                    while (acc_t + bi + 1 <= t) { acc_t += bi + 1; ++bi; }
                    const int bj = t - acc_t;
                    const int i0 = bi * TS, j0 = bj * TS;
                    double acc[TS][TS];
#pragma unroll
                    for (int r = 0; r < TS; ++r)
#pragma unroll
                        for (int c = 0; c < TS; ++c)
                            acc[r][c] = rec[D::REC_H + D::h_off(i0 + r, j0 + c)] + ((i0 + r == j0 + c) ? sigma : 0.0);
                    km_tile<TS, NX>(ET + i0, SP, PEt + j0, SP, acc);
#pragma unroll
                    for (int c = 0; c < TS; ++c)
#pragma unroll
                        for (int h = 0; h < TS / 2; ++h)
                            *reinterpret_cast<double2*>(Ma + i0 + 2 * h + (j0 + c) * L::LDM) = make_double2(acc[2 * h][c], acc[2 * h + 1][c]);
                } else {
                    const int i = t - NTRI;
                    double a = rec[D::REC_h + i] - sigma * wp[i];
#pragma unroll
                    for (int kk = 0; kk < NX; ++kk) a = fma(ET[i + kk * SP], PEt[S + kk * SP], a);
                    Ma[i + S * L::LDM] = a;
                }
            }
        }
        group_sync<T>();
        if (tid == 0 && it + 1 < LEN) {  // H, h had their last readers in S3: fetch the next record into the same buffer
            fence_proxy_async();
            issue_stage(k - 1);
        }

        // S4 + S5: register Cholesky of Quu, one right-hand side per thread
        const int n1 = NX + 1, n2 = pdp ? NX : 0;
        if (tid < n1 + n2) {
            double Lr[NU][NU], dr[NU];
#pragma unroll
            for (int j = 0; j < NU; ++j)
#pragma unroll
                for (int i = j; i < NU; ++i) Lr[i][j] = Ma[i + j * L::LDM];
#pragma unroll
            for (int c = 0; c < NU; ++c) {
                double a = Lr[c][c];
#pragma unroll
                for (int q = 0; q < c; ++q) a = fma(-Lr[c][q], Lr[c][q], a);
                if (!(a > 0.0)) { if (!bad) bad = k + 1; a = fabs(a) + 1e-300; }
                const double r = rsqrt(a);
                dr[c] = r;
                Lr[c][c] = a * r;
#pragma unroll
                for (int i = c + 1; i < NU; ++i) {
                    double v = Lr[i][c];
#pragma unroll
                    for (int q = 0; q < c; ++q) v = fma(-Lr[i][q], Lr[c][q], v);
                    Lr[i][c] = v * r;
                }
            }
            for (int c = tid; c < n1 + n2; c += T) {
                double y[NU], z[NU];
                const int yrow = (c < NX) ? c : (c == NX ? 2 * NX : NX + (c - NX - 1));   // row of YT
#pragma unroll
                for (int m = 0; m < NU; ++m) {
                    double r;
                    if (c < NX) r = Ma[(NU + c) + m * L::LDM];            // Qux(m,c) = Qxu(c,m)
                    else if (c == NX) r = Ma[m + S * L::LDM];             // Qu(m)
                    else r = FE[(c - NX - 1) + m * NX];                    // (F+ B)(c', m)
                    y[m] = r;
                }
#pragma unroll
                for (int m = 0; m < NU; ++m) {
                    double v = y[m];
#pragma unroll
                    for (int q = 0; q < m; ++q) v = fma(-Lr[m][q], y[q], v);
                    y[m] = v * dr[m];
                    YT[yrow + m * L::LDY] = y[m];
                }
#pragma unroll
                for (int m = NU - 1; m >= 0; --m) {
                    double v = y[m];
#pragma unroll
                    for (int q = m + 1; q < NU; ++q) v = fma(-Lr[q][m], z[q], v);
                    z[m] = v * dr[m];
                }
#pragma unroll
                for (int m = 0; m < NU; ++m) {
                    Z[m + c * NU] = -z[m];
                    if (c <= NX) ZT[c + m * L::LDZ] = -z[m];
                }
            }
        }
        group_sync<T>();

        // S6: P = Qxx - Yx^T Yx | C += Yg^T Yg | [F f] = F+[A c] + (F+B)[K d] + [0 f+] | p = Qx - Yx^T yu
        {
            constexpr int NBX = NX / TS;
            constexpr int NP = NBX * NBX;                 // tiles of P, and of C
            constexpr int NFJ = L::LDZ / TS;              // column tiles of [F f]
            const int ntiles = pdp ? (2 * NP + NBX * NFJ) : NP;
            for (int t = tid; t < ntiles; t += T) {
                double acc[TS][TS];
#pragma unroll
                for (int r = 0; r < TS; ++r)
#pragma unroll
                    for (int c = 0; c < TS; ++c) acc[r][c] = 0.0;
                if (t < NP) {
                    const int i0 = (t % NBX) * TS, j0 = (t / NBX) * TS;
                    km_tile<TS, NU>(YT + i0, L::LDY, YT + j0, L::LDY, acc);
#pragma unroll
                    for (int r = 0; r < TS; ++r)
#pragma unroll
                        for (int c = 0; c < TS; ++c) {
                            const int i = i0 + r, j = j0 + c;
                            const int hi = max(i, j), lo = min(i, j);
                            PF[i + j * L::LDPF] = Ma[(NU + hi) + (NU + lo) * L::LDM] - acc[r][c];
                        }
                } else if (t < 2 * NP) {
                    const int tt = t - NP;
                    const int i0 = (tt % NBX) * TS, j0 = (tt / NBX) * TS;
                    km_tile<TS, NU>(YT + NX + i0, L::LDY, YT + NX + j0, L::LDY, acc);
#pragma unroll
                    for (int r = 0; r < TS; ++r)
#pragma unroll
                        for (int c = 0; c < TS; ++c) Cn[(i0 + r) + (j0 + c) * NX] += acc[r][c];
                } else {
                    const int tt = t - 2 * NP;
                    const int i0 = (tt % NBX) * TS, j0 = (tt / NBX) * TS;
                    km_tile<TS, NU>(FE + i0, NX, ZT + j0, L::LDZ, acc);
#pragma unroll
                    for (int r = 0; r < TS; ++r)
#pragma unroll
                        for (int c = 0; c < TS; ++c) {
                            const int i = i0 + r, j = j0 + c;
                            if (j < NX) PF[(NX + i) + j * L::LDPF] = FE[i + (NU + j) * NX] + acc[r][c];   // F+A + (F+B)K
                            else if (j == NX) fn[i] += FE[i + S * NX] + acc[r][c];                          // F+c + (F+B)d + f+
                        }
                }
            }
            for (int i = tid; i < NX; i += T) {
                double a = Ma[(NU + i) + S * L::LDM];
#pragma unroll
                for (int m = 0; m < NU; ++m) a = fma(-YT[i + m * L::LDY], YT[2 * NX + m * L::LDY], a);
                pn[i] = a;
            }
            double* fk = fac_b + (size_t)k * D::FREC;
            const int nz = pdp ? NU * D::NRHS : NU * (NX + 1);
            for (int e = tid; e < nz; e += T) fk[e] = Z[e];
        }
        if (it + 1 < LEN) {  // the next stage's record was requested after S3: wait for it and transpose it now
            mbar_wait(&bar[0], (it + 1) & 1);
            transpose_stage();
        }
        group_sync<T>();
    }

    // segment summary (lqr_solver_parallel.hpp:180-187): P, F, C, p, f at the segment entry
    double* sm = p.sum + ((size_t)b * p.S + seg) * D::SREC;
    for (int e = tid; e < NX * NX; e += T) {
        const int i = e % NX, j = e / NX;
        sm[D::SUM_P + e] = PF[i + j * L::LDPF];
        sm[D::SUM_F + e] = PF[NX + i + j * L::LDPF];
        sm[D::SUM_C + e] = Cn[e];
    }
    for (int i = tid; i < NX; i += T) {
        sm[D::SUM_p + i] = pn[i];
        sm[D::SUM_f + i] = fn[i];
    }
    if (bad && tid == 0) atomicMax(&p.status[b], bad);
}

}  // namespace pdplqr
