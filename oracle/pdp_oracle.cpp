// =====================================================================================
// oracle/pdp_oracle.cpp  --  TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
//
// CPU restatement ("port") of the reference's parallel dynamic-programming LQ solve
// (Luyao787/PDP-LQR), written without Eigen because Eigen3 is absent from this image and
// the reference therefore cannot be compiled as shipped.  Every function cites the
// reference file:line it follows; the operation order inside a stage follows the reference.
//
// PARITY STATUS: the reference ships no tests / golden vectors, so this oracle is pinned by
//   (i)   the reference's own headers compiled unmodified against an Eigen-API shim (oracle/_ref, oracle/Makefile
//         target `ref`, tests/test_reference_build.py) -- agreement <= 1e-11 incl. constraints / no-refactor,
//   (ii)  sequential == parallel (S in {2,4,8}, LU and Cholesky condensed variants),
//   (iii) an independent sparse KKT solve in numpy/scipy (tests/test_oracle.py, tests/kkt_ref.py).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library, and only as the checker / CPU baseline.
//
// Flat data layout (all FP64, column-major like Eigen's default), one problem:
//   E  [N][nx*s]   E_k = [B_k A_k]  (nx rows, s = nu+nx cols)      lqr_model.hpp:12-15
//   c  [N][nx]
//   H  [N][s*s]    H_k = [R S; S^T Q]  (u block first)              lqr_model.hpp:17-19
//   h  [N][s]      h_k = [r; q]
//   HN [nx*nx], hN [nx]   terminal cost                              lqr_model.hpp:32-35
//   D  concat_k (nc_k x dim_k) col-major, dim_k = s (k<N) or nx (k=N) lqr_model.hpp:21-24
//   ws [N*s + nx]  w_k = [u_k; x_k], w_N = x_N                      lqr_example.cpp:30-34
//   ys, zs, rho, inv_rho  [sum_k nc_k]
// Batched entry points put a leading batch dimension on every array.
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <omp.h>

namespace oracle {

// ------------------------------------------------------------------ dense helpers (col-major)
// Unblocked lower Cholesky, same recurrence as Eigen's llt_inplace<Lower>::unblocked.
// Returns 0 on success, k+1 if the k-th pivot is not positive.  Strict upper part is zeroed
// (the reference reads .matrixL(), lqr_kernel.hpp:89,126).
static int chol_lower(double* A, int n, int lda) {
    for (int k = 0; k < n; ++k) {
        double x = A[k + k * lda];
        for (int j = 0; j < k; ++j) x -= A[k + j * lda] * A[k + j * lda];
        if (!(x > 0.0)) return k + 1;
        x = std::sqrt(x);
        A[k + k * lda] = x;
        for (int i = k + 1; i < n; ++i) {
            double v = A[i + k * lda];
            for (int j = 0; j < k; ++j) v -= A[i + j * lda] * A[k + j * lda];
            A[i + k * lda] = v / x;
        }
    }
    for (int j = 1; j < n; ++j)
        for (int i = 0; i < j; ++i) A[i + j * lda] = 0.0;
    return 0;
}
// x <- L^{-1} x   (L lower, n x n)
static void trsv_lower(const double* L, int n, int ldl, double* x, int incx = 1) {
    for (int i = 0; i < n; ++i) {
        double v = x[i * incx];
        for (int j = 0; j < i; ++j) v -= L[i + j * ldl] * x[j * incx];
        x[i * incx] = v / L[i + i * ldl];
    }
}
// x <- L^{-T} x
static void trsv_lower_t(const double* L, int n, int ldl, double* x, int incx = 1) {
    for (int i = n - 1; i >= 0; --i) {
        double v = x[i * incx];
        for (int j = i + 1; j < n; ++j) v -= L[j + i * ldl] * x[j * incx];
        x[i * incx] = v / L[i + i * ldl];
    }
}
// C(m x n) = alpha*op(A)*op(B) + beta*C ; ta/tb: 0 = N, 1 = T.  Plain triple loop.
static void gemm(int ta, int tb, int m, int n, int k, double alpha, const double* A, int lda,
                 const double* B, int ldb, double beta, double* C, int ldc) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
            double acc = 0.0;
            for (int l = 0; l < k; ++l) {
                double a = ta ? A[l + i * lda] : A[i + l * lda];
                double b = tb ? B[j + l * ldb] : B[l + j * ldb];
                acc += a * b;
            }
            C[i + j * ldc] = alpha * acc + (beta == 0.0 ? 0.0 : beta * C[i + j * ldc]);
        }
}
static void gemv(int ta, int m, int n, double alpha, const double* A, int lda, const double* x,
                 double beta, double* y) {
    // y (m if !ta else n) = alpha*op(A)*x + beta*y, A is m x n
    if (!ta) {
        for (int i = 0; i < m; ++i) {
            double acc = 0.0;
            for (int j = 0; j < n; ++j) acc += A[i + j * lda] * x[j];
            y[i] = alpha * acc + (beta == 0.0 ? 0.0 : beta * y[i]);
        }
    } else {
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int i = 0; i < m; ++i) acc += A[i + j * lda] * x[i];
            y[j] = alpha * acc + (beta == 0.0 ? 0.0 : beta * y[j]);
        }
    }
}
// LU with partial pivoting (Eigen::PartialPivLU semantics: row swaps, unit-lower L, U).
struct LU {
    int n = 0;
    std::vector<double> a;
    std::vector<int> piv;
    void compute(const double* A, int n_) {
        n = n_;
        a.assign(A, A + (size_t)n * n);
        piv.resize(n);
        for (int k = 0; k < n; ++k) {
            int p = k;
            double best = std::fabs(a[k + k * n]);
            for (int i = k + 1; i < n; ++i)
                if (std::fabs(a[i + k * n]) > best) { best = std::fabs(a[i + k * n]); p = i; }
            piv[k] = p;
            if (p != k)
                for (int j = 0; j < n; ++j) std::swap(a[k + j * n], a[p + j * n]);
            double d = a[k + k * n];
            for (int i = k + 1; i < n; ++i) a[i + k * n] /= d;
            for (int j = k + 1; j < n; ++j) {
                double u = a[k + j * n];
                for (int i = k + 1; i < n; ++i) a[i + j * n] -= a[i + k * n] * u;
            }
        }
    }
    void solve_inplace(double* B, int nrhs, int ldb) const {
        for (int r = 0; r < nrhs; ++r) {
            double* b = B + (size_t)r * ldb;
            for (int k = 0; k < n; ++k) std::swap(b[k], b[piv[k]]);
            for (int i = 0; i < n; ++i) {
                double v = b[i];
                for (int j = 0; j < i; ++j) v -= a[i + j * n] * b[j];
                b[i] = v;
            }
            for (int i = n - 1; i >= 0; --i) {
                double v = b[i];
                for (int j = i + 1; j < n; ++j) v -= a[i + j * n] * b[j];
                b[i] = v / a[i + i * n];
            }
        }
    }
};

// ------------------------------------------------------------------ problem view
struct Dims {
    int nx, nu, N, s;
    std::vector<int> ncs;      // N+1
    std::vector<int64_t> coff; // prefix offsets of constraint vectors, N+2
    std::vector<int64_t> doff; // prefix offsets into D, N+2
    int64_t nc_total = 0, d_total = 0;
    void init(int nx_, int nu_, int N_, const int* ncs_) {
        nx = nx_; nu = nu_; N = N_; s = nx + nu;
        ncs.assign(N + 1, 0);
        if (ncs_) ncs.assign(ncs_, ncs_ + N + 1);
        coff.assign(N + 2, 0); doff.assign(N + 2, 0);
        for (int k = 0; k <= N; ++k) {
            coff[k + 1] = coff[k] + ncs[k];
            doff[k + 1] = doff[k] + (int64_t)ncs[k] * (k < N ? s : nx);
        }
        nc_total = coff[N + 1]; d_total = doff[N + 1];
    }
    int dim(int k) const { return k < N ? s : nx; }
    int64_t ws_off(int k) const { return (int64_t)k * s; }
    int64_t ws_len() const { return (int64_t)N * s + nx; }
};
struct Model {  // pointers into caller-owned flat arrays (the reference also keeps a reference:
                // lqr_solver_parallel.hpp:52)
    const double *E = nullptr, *c = nullptr, *H = nullptr, *h = nullptr, *HN = nullptr, *hN = nullptr,
                 *D = nullptr;
};

// ------------------------------------------------------------------ per-stage scratch
// Mirrors LQRKernelData (lqr_kernel.hpp:8-75) + ParallelLQRKernelData (lqr_kernel_parallel.hpp:9-47)
struct Stage {
    int dim = 0;  // s, or nx for the true terminal node
    std::vector<double> H, h, g, L, lp;
    std::vector<double> G, F, C, f, K, d;  // PDP extras
    void init(int nx, int nu, int nc, bool terminal, bool pdp) {
        dim = terminal ? nx : nx + nu;
        H.assign((size_t)dim * dim, 0.0); h.assign(dim, 0.0); L.assign((size_t)dim * dim, 0.0);
        lp.assign(dim, 0.0); g.assign(nc, 0.0);
        if (pdp) {
            G.assign((size_t)nu * nx, 0.0); F.assign((size_t)nx * nx, 0.0); C.assign((size_t)nx * nx, 0.0);
            f.assign(nx, 0.0); K.assign((size_t)nu * nx, 0.0); d.assign(nu, 0.0);
        }
    }
    double* Lxx(int nx) { return L.data() + (dim - nx) + (size_t)(dim - nx) * dim; }
    const double* Lxx(int nx) const { return L.data() + (dim - nx) + (size_t)(dim - nx) * dim; }
    double* p(int nx) { return lp.data() + (dim - nx); }
    const double* p(int nx) const { return lp.data() + (dim - nx); }
};

struct Scratch {  // V, M, Pb, Pb_tmp, rhoD, rhog  (lqr_kernel.hpp:14-21)
    std::vector<double> V, M, Pb, Pbt, rhoD, rhog, BtFt, Ftmp, ftmp;
    void init(int nx, int nu, int ncmax) {
        int s = nx + nu;
        V.assign((size_t)s * nx, 0); M.assign((size_t)s * s, 0); Pb.assign(nx, 0); Pbt.assign(nx, 0);
        rhoD.assign((size_t)std::max(ncmax, 1) * s, 0); rhog.assign(std::max(ncmax, 1), 0);
        BtFt.assign((size_t)nu * nx, 0); Ftmp.assign((size_t)nx * nx, 0); ftmp.assign(nx, 0);
    }
};

// ------------------------------------------------------------------ stage kernels
// constraint fold-in:  H += D^T diag(rho) D ; h -= D^T (rho .* g)     lqr_kernel.hpp:82-87,106-112
static void fold_constraints(const double* D, const double* rho, int nc, int dim, Stage& st, Scratch& sc,
                             bool with_H) {
    if (nc <= 0) return;
    if (with_H) {
        for (int j = 0; j < dim; ++j)
            for (int i = 0; i < nc; ++i) sc.rhoD[i + (size_t)j * nc] = rho[i] * D[i + (size_t)j * nc];
        gemm(1, 0, dim, dim, nc, 1.0, D, nc, sc.rhoD.data(), nc, 1.0, st.H.data(), dim);
    }
    for (int i = 0; i < nc; ++i) sc.rhog[i] = rho[i] * st.g[i];
    gemv(1, nc, dim, -1.0, D, nc, sc.rhog.data(), 1.0, st.h.data());
}

// LQRKernel::terminal_step_with_factorization            lqr_kernel.hpp:79-91
static int terminal_fact(const double* D, const double* rho, int nc, int nx, Stage& st, Scratch& sc) {
    fold_constraints(D, rho, nc, st.dim, st, sc, true);
    st.L = st.H;
    int info = chol_lower(st.L.data(), st.dim, st.dim);
    st.lp = st.h;
    (void)nx;
    return info;
}
// LQRKernel::terminal_step_without_factorization         lqr_kernel.hpp:93-101
static void terminal_nofact(const double* D, const double* rho, int nc, Stage& st, Scratch& sc) {
    fold_constraints(D, rho, nc, st.dim, st, sc, false);
    st.lp = st.h;
}
// affine tail shared by both step variants               lqr_kernel.hpp:138-146 / 168-177
static void affine_tail(const double* E, const double* c, int nx, int nu, const Stage& nxt, Stage& st,
                        Scratch& sc) {
    const int s = nx + nu;
    const double* Lxxn = nxt.Lxx(nx);
    const int ldn = nxt.dim;
    // Pb_tmp = Lxx_next^T c ; Pb = Lxx_next Pb_tmp + p_next
    for (int j = 0; j < nx; ++j) {
        double acc = 0;
        for (int i = 0; i < nx; ++i) acc += Lxxn[i + (size_t)j * ldn] * c[i];
        sc.Pbt[j] = acc;
    }
    const double* pn = nxt.p(nx);
    for (int i = 0; i < nx; ++i) {
        double acc = 0;
        for (int j = 0; j < nx; ++j) acc += Lxxn[i + (size_t)j * ldn] * sc.Pbt[j];
        sc.Pb[i] = acc + pn[i];
    }
    // lp = h + E^T Pb
    st.lp = st.h;
    gemv(1, nx, s, 1.0, E, nx, sc.Pb.data(), 1.0, st.lp.data());
    // lu <- Luu^{-1} lu ; p -= Lxu lu
    trsv_lower(st.L.data(), nu, s, st.lp.data());
    for (int i = 0; i < nx; ++i) {
        double acc = 0;
        for (int j = 0; j < nu; ++j) acc += st.L[(nu + i) + (size_t)j * s] * st.lp[j];
        st.lp[nu + i] -= acc;
    }
}
// LQRKernel::step_with_factorization                      lqr_kernel.hpp:103-147
static int step_fact(const double* E, const double* c, const double* D, const double* rho, int nc, int nx,
                     int nu, const Stage& nxt, Stage& st, Scratch& sc) {
    const int s = nx + nu;
    fold_constraints(D, rho, nc, s, st, sc, true);
    // V = E^T Lxx_next  (s x nx)
    gemm(1, 0, s, nx, nx, 1.0, E, nx, nxt.Lxx(nx), nxt.dim, 0.0, sc.V.data(), s);
    // M = H + V V^T
    sc.M = st.H;
    gemm(0, 1, s, s, nx, 1.0, sc.V.data(), s, sc.V.data(), s, 1.0, sc.M.data(), s);
    st.L = sc.M;
    int info = chol_lower(st.L.data(), s, s);
    affine_tail(E, c, nx, nu, nxt, st, sc);
    return info;
}
// LQRKernel::step_without_factorization                   lqr_kernel.hpp:149-178
static void step_nofact(const double* E, const double* c, const double* D, const double* rho, int nc, int nx,
                        int nu, const Stage& nxt, Stage& st, Scratch& sc) {
    fold_constraints(D, rho, nc, nx + nu, st, sc, false);
    affine_tail(E, c, nx, nu, nxt, st, sc);
}
// PDP extras of ParallelLQRKernel::step_with_factorization   lqr_kernel_parallel.hpp:97-135
static void pdp_extras_fact(const double* E, const double* c, int nx, int nu, const Stage& nxt, Stage& st,
                            Scratch& sc) {
    const int s = nx + nu;
    const double* Luu = st.L.data();
    const double* B = E;                       // E.topLeftCorner(nx,nu)
    const double* A = E + (size_t)nu * nx;     // E.topRightCorner(nx,nx)
    // K = -Lxu^T ; d = -lu ; Luu^T-solve both
    for (int j = 0; j < nx; ++j)
        for (int i = 0; i < nu; ++i) st.K[i + (size_t)j * nu] = -st.L[(nu + j) + (size_t)i * s];
    for (int i = 0; i < nu; ++i) st.d[i] = -st.lp[i];
    for (int j = 0; j < nx; ++j) trsv_lower_t(Luu, nu, s, st.K.data() + (size_t)j * nu);
    trsv_lower_t(Luu, nu, s, st.d.data());
    // BtFt = B^T F_next^T ; G = -BtFt ; Luu-solve
    gemm(1, 1, nu, nx, nx, 1.0, B, nx, nxt.F.data(), nx, 0.0, sc.BtFt.data(), nu);
    for (size_t i = 0; i < st.G.size(); ++i) st.G[i] = -sc.BtFt[i];
    for (int j = 0; j < nx; ++j) trsv_lower(Luu, nu, s, st.G.data() + (size_t)j * nu);
    // F_tmp = A + B K ; F = F_next F_tmp
    std::memcpy(sc.Ftmp.data(), A, sizeof(double) * nx * nx);
    gemm(0, 0, nx, nx, nu, 1.0, B, nx, st.K.data(), nu, 1.0, sc.Ftmp.data(), nx);
    gemm(0, 0, nx, nx, nx, 1.0, nxt.F.data(), nx, sc.Ftmp.data(), nx, 0.0, st.F.data(), nx);
    // f_tmp = c + B d ; f = F_next f_tmp + f_next
    for (int i = 0; i < nx; ++i) sc.ftmp[i] = c[i];
    gemv(0, nx, nu, 1.0, B, nx, st.d.data(), 1.0, sc.ftmp.data());
    st.f = nxt.f;
    gemv(0, nx, nx, 1.0, nxt.F.data(), nx, sc.ftmp.data(), 1.0, st.f.data());
    // C = C_next + G^T G
    st.C = nxt.C;
    gemm(1, 0, nx, nx, nu, 1.0, st.G.data(), nu, st.G.data(), nu, 1.0, st.C.data(), nx);
}
// PDP extras of ParallelLQRKernel::step_without_factorization  lqr_kernel_parallel.hpp:147-167
static void pdp_extras_nofact(const double* E, const double* c, int nx, int nu, const Stage& nxt, Stage& st,
                              Scratch& sc) {
    const int s = nx + nu;
    for (int i = 0; i < nu; ++i) st.d[i] = -st.lp[i];
    trsv_lower_t(st.L.data(), nu, s, st.d.data());
    for (int i = 0; i < nx; ++i) sc.ftmp[i] = c[i];
    gemv(0, nx, nu, 1.0, E, nx, st.d.data(), 1.0, sc.ftmp.data());
    st.f = nxt.f;
    gemv(0, nx, nx, 1.0, nxt.F.data(), nx, sc.ftmp.data(), 1.0, st.f.data());
}
// LQRKernel::forward_step (lqr_kernel.hpp:180-204) and the PDP version
// (lqr_kernel_parallel.hpp:170-205): u = Luu^{-T}(-lu - Lxu^T x [+ G uhat]); x+ = c + A x + B u
static void forward_step(const double* E, const double* c, int nx, int nu, const Stage& st, const double* uhat,
                         double* w, double* x_next) {
    const int s = nx + nu;
    double* u = w;
    const double* x = w + nu;
    for (int i = 0; i < nu; ++i) {
        double acc = -st.lp[i];
        for (int j = 0; j < nx; ++j) acc -= st.L[(nu + j) + (size_t)i * s] * x[j];
        if (uhat)
            for (int j = 0; j < nx; ++j) acc += st.G[i + (size_t)j * nu] * uhat[j];
        u[i] = acc;
    }
    trsv_lower_t(st.L.data(), nu, s, u);
    if (x_next) {
        for (int i = 0; i < nx; ++i) {
            double acc = c[i];
            for (int j = 0; j < nx; ++j) acc += E[i + (size_t)(nu + j) * nx] * x[j];
            for (int j = 0; j < nu; ++j) acc += E[i + (size_t)j * nx] * u[j];
            x_next[i] = acc;
        }
    }
}

// ------------------------------------------------------------------ condensed interface system
// CondensedSystemLUSolver (condensed_system.hpp:32-147) and CondensedSystemCholeskySolver (:151-299)
struct Condensed {
    int nx = 0, S = 0, type = 0;  // type 0 = LU, 1 = CHOLESKY
    struct Seg {
        std::vector<double> A, C, P, At, Pinv, c, p, PC, PA, Dm, cbar, xhat, uhat, Pchol, Cchol;
        LU lu;
    };
    std::vector<Seg> w;
    void init(int nx_, int S_, int type_) {
        nx = nx_; S = S_; type = type_;
        w.resize(S);
        for (auto& g : w) {
            size_t n2 = (size_t)nx * nx;
            g.A.assign(n2, 0); g.C.assign(n2, 0); g.P.assign(n2, 0); g.At.assign(n2, 0); g.Pinv.assign(n2, 0);
            g.PC.assign(n2, 0); g.PA.assign(n2, 0); g.Dm.assign(n2, 0); g.Pchol.assign(n2, 0); g.Cchol.assign(n2, 0);
            g.c.assign(nx, 0); g.p.assign(nx, 0); g.cbar.assign(nx, 0); g.xhat.assign(nx, 0); g.uhat.assign(nx, 0);
        }
    }
    // update_segment_data(Lxx, A, C, p, c, id)   condensed_system.hpp:64-74 / 183-195
    void update_full(const double* Lxx, int ldl, const double* F, const double* C, const double* p, const double* f,
                     int id) {
        Seg& g = w[id];
        gemm(0, 1, nx, nx, nx, 1.0, Lxx, ldl, Lxx, ldl, 0.0, g.P.data(), nx);
        std::memcpy(g.A.data(), F, sizeof(double) * nx * nx);
        std::memcpy(g.C.data(), C, sizeof(double) * nx * nx);
        std::memcpy(g.p.data(), p, sizeof(double) * nx);
        std::memcpy(g.c.data(), f, sizeof(double) * nx);
        if (type == 1) {
            for (int j = 0; j < nx; ++j)
                for (int i = 0; i < nx; ++i) g.At[i + (size_t)j * nx] = F[j + (size_t)i * nx];
            std::fill(g.Pinv.begin(), g.Pinv.end(), 0.0);
            for (int i = 0; i < nx; ++i) g.Pinv[i + (size_t)i * nx] = 1.0;
        }
    }
    // update_segment_data(p, c, id)              condensed_system.hpp:76-80 / 197-201
    void update_affine(const double* p, const double* f, int id) {
        std::memcpy(w[id].p.data(), p, sizeof(double) * nx);
        std::memcpy(w[id].c.data(), f, sizeof(double) * nx);
    }
    static void chol_solve_inplace(const double* Lc, int n, double* B, int nrhs) {
        for (int r = 0; r < nrhs; ++r) {
            trsv_lower(Lc, n, n, B + (size_t)r * n);
            trsv_lower_t(Lc, n, n, B + (size_t)r * n);
        }
    }
    bool backward() {
        if (type == 0) {  // condensed_system.hpp:82-103
            for (int i = S - 2; i >= 0; --i) {
                Seg& g = w[i];
                const Seg& nx_ = w[i + 1];
                gemm(0, 0, nx, nx, nx, 1.0, g.C.data(), nx, nx_.P.data(), nx, 0.0, g.PC.data(), nx);
                for (int d = 0; d < nx; ++d) g.PC[d + (size_t)d * nx] += 1.0;
                gemm(0, 0, nx, nx, nx, 1.0, nx_.P.data(), nx, g.A.data(), nx, 0.0, g.PA.data(), nx);
                g.lu.compute(g.PC.data(), nx);
                g.Dm = g.A;
                g.lu.solve_inplace(g.Dm.data(), nx, nx);
                gemm(1, 0, nx, nx, nx, 1.0, g.Dm.data(), nx, g.PA.data(), nx, 1.0, g.P.data(), nx);
            }
            return true;
        }
        if (S < 2) return true;  // reference touches workspace_[1] unconditionally (UB for S=1)
        // condensed_system.hpp:203-250
        for (int i = S - 2; i >= 0; --i) {
            Seg& g = w[i];
            Seg& n1 = w[i + 1];
            n1.Pchol = n1.P;
            if (chol_lower(n1.Pchol.data(), nx, nx)) return false;
            chol_solve_inplace(n1.Pchol.data(), nx, n1.Pinv.data(), nx);
            for (size_t e = 0; e < g.C.size(); ++e) g.C[e] += n1.Pinv[e];
            g.Cchol = g.C;
            if (chol_lower(g.Cchol.data(), nx, nx)) return false;
            if (i >= 1) {
                chol_solve_inplace(g.Cchol.data(), nx, g.A.data(), nx);
                gemm(0, 0, nx, nx, nx, 1.0, g.At.data(), nx, g.A.data(), nx, 1.0, g.P.data(), nx);
            }
        }
        return true;
    }
    void forward(const double* x0) {
        if (type == 0) {  // condensed_system.hpp:105-138
            for (int i = S - 2; i >= 0; --i) {
                Seg& g = w[i];
                const Seg& n1 = w[i + 1];
                g.cbar = n1.p;
                gemv(0, nx, nx, 1.0, n1.P.data(), nx, g.c.data(), 1.0, g.cbar.data());
                gemv(1, nx, nx, 1.0, g.Dm.data(), nx, g.cbar.data(), 1.0, g.p.data());
            }
            std::memcpy(w[0].xhat.data(), x0, sizeof(double) * nx);
            for (int i = 0; i < S - 1; ++i) {
                Seg& g = w[i];
                Seg& n1 = w[i + 1];
                gemv(0, nx, nx, 1.0, g.A.data(), nx, g.xhat.data(), 1.0, g.c.data());
                gemv(0, nx, nx, -1.0, g.C.data(), nx, n1.p.data(), 1.0, g.c.data());
                n1.xhat = g.c;
                g.lu.solve_inplace(n1.xhat.data(), 1, nx);
                g.uhat = n1.p;
                gemv(0, nx, nx, 1.0, n1.P.data(), nx, n1.xhat.data(), 1.0, g.uhat.data());
            }
            return;
        }
        // condensed_system.hpp:252-290
        for (int i = S - 2; i >= 0; --i) {
            Seg& g = w[i];
            Seg& n1 = w[i + 1];
            chol_solve_inplace(n1.Pchol.data(), nx, n1.p.data(), 1);
            for (int e = 0; e < nx; ++e) g.c[e] += n1.p[e];
            if (i >= 1) gemv(1, nx, nx, 1.0, g.A.data(), nx, g.c.data(), 1.0, g.p.data());
        }
        std::memcpy(w[0].xhat.data(), x0, sizeof(double) * nx);
        for (int i = 0; i < S - 1; ++i) {
            Seg& g = w[i];
            Seg& n1 = w[i + 1];
            g.uhat = g.c;
            gemv(1, nx, nx, 1.0, g.At.data(), nx, g.xhat.data(), 1.0, g.uhat.data());  // At^T xhat = F xhat
            chol_solve_inplace(g.Cchol.data(), nx, g.uhat.data(), 1);
            for (int e = 0; e < nx; ++e) n1.xhat[e] = -n1.p[e];
            gemv(0, nx, nx, 1.0, n1.Pinv.data(), nx, g.uhat.data(), 1.0, n1.xhat.data());
        }
    }
};

// ------------------------------------------------------------------ solver (sequential + PDP)
struct Solver {
    Dims dm;
    Model md;
    bool parallel = false;
    int S = 1;
    int nthreads = 1;
    std::vector<int> seg_start, seg_len;
    std::vector<std::vector<Stage>> seg;  // seg[i][k], Nseg_i + 1 entries (interface node shared)
    std::vector<Scratch> scratch;         // one per segment
    Condensed cond;
    int status = 0;                       // first non-PD pivot seen (0 = ok)

    // ctor: LQRSolver (lqr_solver.hpp:31-39) / LQRParallelSolver (lqr_solver_parallel.hpp:64-113)
    void init(int nx, int nu, int N, const int* ncs, bool par, int S_, bool load_balancing, int cond_type,
              int nthreads_) {
        dm.init(nx, nu, N, ncs);
        parallel = par;
        S = par ? S_ : 1;
        nthreads = std::max(1, nthreads_);
        seg_start.resize(S); seg_len.resize(S);
        const double alpha = 1.55;  // lqr_solver_parallel.hpp:70
        const double scale = load_balancing ? alpha : 1.0;
        for (int i = 0; i < S; ++i) {
            int st = (i == 0) ? 0 : seg_start[i - 1] + seg_len[i - 1];
            int len = (i < S - 1) ? int(N / (scale + S - 1)) : N - st;
            if (!par) len = N;
            seg_start[i] = st; seg_len[i] = len;
        }
        int ncmax = 0;
        for (int v : dm.ncs) ncmax = std::max(ncmax, v);
        seg.resize(S); scratch.resize(S);
        for (int i = 0; i < S; ++i) {
            seg[i].resize(seg_len[i] + 1);
            for (int k = 0; k <= seg_len[i]; ++k) {
                int gk = seg_start[i] + k;
                seg[i][k].init(nx, nu, dm.ncs[gk], gk == N, par);
            }
            scratch[i].init(nx, nu, ncmax);
        }
        if (par) cond.init(nx, S, cond_type);
    }
    const double* Ek(int k) const { return md.E + (size_t)k * dm.nx * dm.s; }
    const double* ck(int k) const { return md.c + (size_t)k * dm.nx; }
    const double* Dk(int k) const { return md.D ? md.D + dm.doff[k] : nullptr; }

    // update_problem_data     lqr_solver.hpp:41-56 / lqr_solver_parallel.hpp:115-140
    void update_problem_data(const double* ws, const double* ys, const double* zs, const double* inv_rho,
                             double sigma) {
#pragma omp parallel for num_threads(std::min(nthreads, S)) schedule(static) if (S > 1 && nthreads > 1)
        for (int i = 0; i < S; ++i) {
            for (int k = 0; k <= seg_len[i]; ++k) {
                int gk = seg_start[i] + k;
                Stage& st = seg[i][k];
                int dim = st.dim;
                const double* Hm = gk < dm.N ? md.H + (size_t)gk * dm.s * dm.s : md.HN;
                const double* hm = gk < dm.N ? md.h + (size_t)gk * dm.s : md.hN;
                std::memcpy(st.H.data(), Hm, sizeof(double) * dim * dim);
                for (int d = 0; d < dim; ++d) st.H[d + (size_t)d * dim] += sigma;
                const double* w = ws + dm.ws_off(gk);
                for (int d = 0; d < dim; ++d) st.h[d] = hm[d] - sigma * w[d];
                int nc = dm.ncs[gk];
                for (int e = 0; e < nc; ++e)
                    st.g[e] = zs[dm.coff[gk] + e] - inv_rho[dm.coff[gk] + e] * ys[dm.coff[gk] + e];
            }
        }
    }
    // reduction_per_thread / reduction_without_factorization   lqr_solver_parallel.hpp:164-211
    // (sequential: lqr_solver.hpp:58-70)
    void reduce_segment(int i, const double* rho, bool fact) {
        const int nx = dm.nx, nu = dm.nu;
        const int N0 = seg_start[i], N1 = N0 + seg_len[i];
        const bool is_last = (i == S - 1);
        Scratch& sc = scratch[i];
        Stage& term = seg[i].back();
        if (is_last) {
            const double* r = rho ? rho + dm.coff[N1] : nullptr;
            if (fact) {
                int info = terminal_fact(Dk(N1), r, dm.ncs[N1], nx, term, sc);
                if (info && !status) status = info;
            } else
                terminal_nofact(Dk(N1), r, dm.ncs[N1], term, sc);
        } else {  // lqr_kernel_parallel.hpp:61-65 / 80-84
            std::fill(term.L.begin(), term.L.end(), 0.0);
            std::fill(term.lp.begin(), term.lp.end(), 0.0);
            std::fill(term.C.begin(), term.C.end(), 0.0);
            std::fill(term.f.begin(), term.f.end(), 0.0);
            std::fill(term.F.begin(), term.F.end(), 0.0);
            for (int d = 0; d < nx; ++d) term.F[d + (size_t)d * nx] = 1.0;
        }
        for (int k = N1 - 1; k >= N0; --k) {
            Stage& st = seg[i][k - N0];
            const Stage& nxt = seg[i][k - N0 + 1];
            const double* r = rho ? rho + dm.coff[k] : nullptr;
            if (fact) {
                int info = step_fact(Ek(k), ck(k), Dk(k), r, dm.ncs[k], nx, nu, nxt, st, sc);
                if (info && !status) status = info;
                if (parallel && !is_last) pdp_extras_fact(Ek(k), ck(k), nx, nu, nxt, st, sc);
            } else {
                step_nofact(Ek(k), ck(k), Dk(k), r, dm.ncs[k], nx, nu, nxt, st, sc);
                if (parallel && !is_last) pdp_extras_nofact(Ek(k), ck(k), nx, nu, nxt, st, sc);
            }
        }
        if (parallel) {
            Stage& s0 = seg[i][0];
            if (fact)
                cond.update_full(s0.Lxx(nx), s0.dim, s0.F.data(), s0.C.data(), s0.p(nx), s0.f.data(), i);
            else
                cond.update_affine(s0.p(nx), s0.f.data(), i);
        }
    }
    bool backward(const double* rho, bool fact) {
#pragma omp parallel for num_threads(std::min(nthreads, S)) schedule(static) if (S > 1 && nthreads > 1)
        for (int i = 0; i < S; ++i) reduce_segment(i, rho, fact);
        bool ok = true;
        if (parallel && fact) ok = cond.backward();  // lqr_solver_parallel.hpp:145
        return ok;
    }
    // forward    lqr_solver.hpp:72-77 / lqr_solver_parallel.hpp:213-238
    void forward(const double* x0, double* ws) {
        const int nx = dm.nx, nu = dm.nu;
        if (!parallel) {
            std::memcpy(ws + nu, x0, sizeof(double) * nx);
            for (int k = 0; k < dm.N; ++k) {
                double* wn = ws + dm.ws_off(k + 1);
                forward_step(Ek(k), ck(k), nx, nu, seg[0][k], nullptr, ws + dm.ws_off(k),
                             (k + 1 < dm.N) ? wn + nu : wn);
            }
            return;
        }
        cond.forward(x0);
#pragma omp parallel for num_threads(std::min(nthreads, S)) schedule(static) if (S > 1 && nthreads > 1)
        for (int i = 0; i < S; ++i) {
            const int N0 = seg_start[i], N1 = N0 + seg_len[i];
            const bool is_last = (i == S - 1);
            // ws[N0].tail(nx) = xhat(i)
            double* w0 = ws + dm.ws_off(N0);
            std::memcpy((N0 < dm.N) ? w0 + nu : w0, cond.w[i].xhat.data(), sizeof(double) * nx);
            const double* uhat = is_last ? nullptr : cond.w[i].uhat.data();
            for (int k = N0; k < N1; ++k) {
                double* wn = ws + dm.ws_off(k + 1);
                double* xn = (k + 1 < dm.N) ? wn + nu : wn;
                bool update_x_next = is_last || (k < N1 - 1);
                forward_step(Ek(k), ck(k), nx, nu, seg[i][k - N0], uhat, ws + dm.ws_off(k),
                             update_x_next ? xn : nullptr);
            }
        }
    }
    Stage& stage_global(int k, int* seg_id = nullptr) {
        // owner of stage k as a non-terminal local entry (k < N), or the true terminal (k == N)
        for (int i = 0; i < S; ++i) {
            int N0 = seg_start[i], N1 = N0 + seg_len[i];
            if ((k >= N0 && k < N1) || (k == dm.N && i == S - 1)) {
                if (seg_id) *seg_id = i;
                return seg[i][k - N0];
            }
        }
        if (seg_id) *seg_id = S - 1;
        return seg[S - 1].back();
    }
};

}  // namespace oracle

// ===================================================================== C API (ctypes / bench)
using oracle::Solver;
extern "C" {

void* oracle_create(int nx, int nu, int N, const int* ncs, int parallel, int num_segments, int load_balancing,
                    int condensed_type, int nthreads) {
    if (N < 1 || nx < 1 || nu < 1) return nullptr;  // lqr_model.hpp:75-77
    if (parallel && num_segments < 1) return nullptr;
    auto* s = new Solver();
    s->init(nx, nu, N, ncs, parallel != 0, num_segments, load_balancing != 0, condensed_type, nthreads);
    return s;
}
void oracle_destroy(void* o) { delete static_cast<Solver*>(o); }
void oracle_set_model(void* o, const double* E, const double* c, const double* H, const double* h,
                      const double* HN, const double* hN, const double* D) {
    auto* s = static_cast<Solver*>(o);
    s->md.E = E; s->md.c = c; s->md.H = H; s->md.h = h; s->md.HN = HN; s->md.hN = hN; s->md.D = D;
}
void oracle_update_problem_data(void* o, const double* ws, const double* ys, const double* zs,
                                const double* inv_rho, double sigma) {
    static_cast<Solver*>(o)->update_problem_data(ws, ys, zs, inv_rho, sigma);
}
int oracle_backward(void* o, const double* rho) { return static_cast<Solver*>(o)->backward(rho, true) ? 0 : 1; }
int oracle_backward_without_factorization(void* o, const double* rho) {
    return static_cast<Solver*>(o)->backward(rho, false) ? 0 : 1;
}
void oracle_forward(void* o, const double* x0, double* ws) { static_cast<Solver*>(o)->forward(x0, ws); }
int oracle_status(void* o) { return static_cast<Solver*>(o)->status; }
int oracle_num_segments(void* o) { return static_cast<Solver*>(o)->S; }
void oracle_get_partition(void* o, int* starts, int* lens) {
    auto* s = static_cast<Solver*>(o);
    for (int i = 0; i < s->S; ++i) { starts[i] = s->seg_start[i]; lens[i] = s->seg_len[i]; }
}
// Gains as defined at lqr_kernel_parallel.hpp:105-108: K_k = -Luu^{-T} Lxu^T, d_k = -Luu^{-T} lu  (k < N).
// (Segment-local inside non-last PDP segments.)  Gt_k = Luu^{-T} G_k (zero for last segment / sequential).
void oracle_get_gains(void* o, double* K, double* d, double* Gt) {
    auto* s = static_cast<Solver*>(o);
    const int nx = s->dm.nx, nu = s->dm.nu, sd = nx + nu;
    for (int k = 0; k < s->dm.N; ++k) {
        int sid;
        oracle::Stage& st = s->stage_global(k, &sid);
        double* Kk = K + (size_t)k * nu * nx;
        double* dk = d + (size_t)k * nu;
        for (int j = 0; j < nx; ++j)
            for (int i = 0; i < nu; ++i) Kk[i + (size_t)j * nu] = -st.L[(nu + j) + (size_t)i * sd];
        for (int i = 0; i < nu; ++i) dk[i] = -st.lp[i];
        for (int j = 0; j < nx; ++j) oracle::trsv_lower_t(st.L.data(), nu, sd, Kk + (size_t)j * nu);
        oracle::trsv_lower_t(st.L.data(), nu, sd, dk);
        if (Gt) {
            double* Gk = Gt + (size_t)k * nu * nx;
            if (s->parallel && sid != s->S - 1) {
                std::memcpy(Gk, st.G.data(), sizeof(double) * nu * nx);
                for (int j = 0; j < nx; ++j) oracle::trsv_lower_t(st.L.data(), nu, sd, Gk + (size_t)j * nu);
            } else
                std::memset(Gk, 0, sizeof(double) * nu * nx);
        }
    }
}
// Value function P_k = Lxx Lxx^T, p_k = lp.tail(nx) for k = 0..N (segment-local in non-last segments).
void oracle_get_value(void* o, double* P, double* p) {
    auto* s = static_cast<Solver*>(o);
    const int nx = s->dm.nx;
    for (int k = 0; k <= s->dm.N; ++k) {
        oracle::Stage& st = s->stage_global(k);
        oracle::gemm(0, 1, nx, nx, nx, 1.0, st.Lxx(nx), st.dim, st.Lxx(nx), st.dim, 0.0, P + (size_t)k * nx * nx, nx);
        std::memcpy(p + (size_t)k * nx, st.p(nx), sizeof(double) * nx);
    }
}
// Costates lambda_k (k = 1..N) of the last solve: the step the reference has written but commented out,
// `lambda+ = Lxx+ (Lxx+^T x+) + p+` (lqr_kernel.hpp:205-211) and, inside non-last segments, `... + F+^T uhat`
// (lqr_kernel_parallel.hpp:207-216).  `ws` is the trajectory forward returned; lam[(k-1) nx + i] = lambda_k(i).
void oracle_get_costates(void* o, const double* ws, double* lam) {
    auto* s = static_cast<Solver*>(o);
    const int nx = s->dm.nx, nu = s->dm.nu, N = s->dm.N, sd = nx + nu;
    std::vector<double> t(nx);
    for (int k = 1; k <= N; ++k) {
        int sid = 0;
        oracle::Stage& st = s->stage_global(k, &sid);
        const double* x = (k < N) ? ws + (size_t)k * sd + nu : ws + (size_t)N * sd;
        double* out = lam + (size_t)(k - 1) * nx;
        const double* Lxx = st.Lxx(nx);
        for (int j = 0; j < nx; ++j) {           // t = Lxx^T x
            double acc = 0.0;
            for (int i = j; i < nx; ++i) acc += Lxx[i + (size_t)j * st.dim] * x[i];
            t[j] = acc;
        }
        for (int i = 0; i < nx; ++i) {           // lambda = Lxx t + p
            double acc = st.p(nx)[i];
            for (int j = 0; j <= i; ++j) acc += Lxx[i + (size_t)j * st.dim] * t[j];
            out[i] = acc;
        }
        if (s->parallel && sid < s->S - 1) {     // + F^T uhat of the owning segment
            const double* uh = s->cond.w[sid].uhat.data();
            for (int i = 0; i < nx; ++i) {
                double acc = 0.0;
                for (int j = 0; j < nx; ++j) acc += st.F[j + (size_t)i * nx] * uh[j];
                out[i] += acc;
            }
        }
    }
}
// Segment summaries as sent to the condensed solver (lqr_solver_parallel.hpp:180-187), read from the
// segment workspaces (the condensed solver modifies its own copies in place).
void oracle_get_summary(void* o, int seg_id, double* P, double* p, double* F, double* f, double* C) {
    auto* s = static_cast<Solver*>(o);
    const int nx = s->dm.nx;
    oracle::Stage& st = s->seg[seg_id][0];
    oracle::gemm(0, 1, nx, nx, nx, 1.0, st.Lxx(nx), st.dim, st.Lxx(nx), st.dim, 0.0, P, nx);
    std::memcpy(p, st.p(nx), sizeof(double) * nx);
    if (s->parallel) {
        std::memcpy(F, st.F.data(), sizeof(double) * nx * nx);
        std::memcpy(f, st.f.data(), sizeof(double) * nx);
        std::memcpy(C, st.C.data(), sizeof(double) * nx * nx);
    }
}
void oracle_get_interface(void* o, double* xhat, double* uhat) {
    auto* s = static_cast<Solver*>(o);
    const int nx = s->dm.nx;
    if (!s->parallel) return;
    for (int i = 0; i < s->S; ++i) {
        std::memcpy(xhat + (size_t)i * nx, s->cond.w[i].xhat.data(), sizeof(double) * nx);
        std::memcpy(uhat + (size_t)i * nx, s->cond.w[i].uhat.data(), sizeof(double) * nx);
    }
}

// Batched CPU baseline: `batch` independent problems, each a sequential LQRSolver-semantics solve
// (update_problem_data + backward + forward), OpenMP over problems.
// `mode`: 1 = factorizing backward, 0 = backward_without_factorization (needs a previous mode-1 call on the
// same pool).  `pool` (from oracle_batch_pool_create) keeps one solver per problem so cached factors persist.
struct BatchPool {
    std::vector<Solver*> solvers;
    oracle::Dims dm;
};
void* oracle_batch_pool_create(int nx, int nu, int N, const int* ncs, int batch) {
    auto* p = new BatchPool();
    p->dm.init(nx, nu, N, ncs);
    p->solvers.resize(batch);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < batch; ++b) {
        p->solvers[b] = new Solver();
        p->solvers[b]->init(nx, nu, N, ncs, false, 1, false, 0, 1);
    }
    return p;
}
void oracle_batch_pool_destroy(void* pp) {
    auto* p = static_cast<BatchPool*>(pp);
    for (auto* s : p->solvers) delete s;
    delete p;
}
int oracle_batch_solve(void* pp, const double* E, const double* c, const double* H, const double* h,
                       const double* HN, const double* hN, const double* D, const double* ws_in,
                       const double* ys, const double* zs, const double* rho, const double* inv_rho, double sigma,
                       const double* x0, double* ws_out, int mode, int nthreads) {
    auto* p = static_cast<BatchPool*>(pp);
    const oracle::Dims& dm = p->dm;
    const int batch = (int)p->solvers.size();
    const size_t nx = dm.nx, s = dm.s, N = dm.N;
    int bad = 0;
    if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16) reduction(+ : bad)
    for (int b = 0; b < batch; ++b) {
        Solver* sv = p->solvers[b];
        sv->md.E = E + (size_t)b * N * nx * s;
        sv->md.c = c + (size_t)b * N * nx;
        sv->md.H = H + (size_t)b * N * s * s;
        sv->md.h = h + (size_t)b * N * s;
        sv->md.HN = HN + (size_t)b * nx * nx;
        sv->md.hN = hN + (size_t)b * nx;
        sv->md.D = D ? D + (size_t)b * dm.d_total : nullptr;
        const size_t co = (size_t)b * dm.nc_total;
        sv->status = 0;
        sv->update_problem_data(ws_in + (size_t)b * dm.ws_len(), ys ? ys + co : nullptr, zs ? zs + co : nullptr,
                                inv_rho ? inv_rho + co : nullptr, sigma);
        sv->backward(rho ? rho + co : nullptr, mode != 0);
        sv->forward(x0 + (size_t)b * nx, ws_out + (size_t)b * dm.ws_len());
        bad += (sv->status != 0);
    }
    return bad;
}
int oracle_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
