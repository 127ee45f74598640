// lqr::LQRCudaSolver -- header-only C++17 host class that restores the reference's solver protocol on top of
// the C ABI (include/pdplqr.h, libpdplqr.so).  Drop-in for
//     lqr::LQRParallelSolver   /root/reference include/clqr/lqr/lqr_solver_parallel.hpp:19-62
//     lqr::LQRSolver           /root/reference include/clqr/lqr/lqr_solver.hpp:9-28      (num_segments = 1)
// Same constructor arguments, same four calls with the same argument meaning:
//     update_problem_data(ws, ys, zs, inv_rho_vecs, sigma) -> backward(rho_vecs) | backward_without_factorization(rho_vecs)
//     -> forward(x0, ws)
// Differences a maintainer must know:
//   * the model is uploaded in the constructor; after mutating the caller-owned LQRModel call sync_model()
//     (the reference re-reads `const LQRModel&` on every call, lqr_solver_parallel.hpp:52);
//   * errors are exceptions (std::runtime_error) carrying the C-ABI message; a non-positive-definite stage, which
//     the reference never detects (lqr_kernel.hpp:89,126), is available through not_positive_definite();
//   * additions: solve(), gains / interface accessors;
//   * the constructor takes the reference's own enum lqr::CondensedSystemSolverType: replacing LQRParallelSolver by
//     LQRCudaSolver is a change of the class name and one #include.
// With Eigen3 installed this header uses the reference's own clqr/lqr_model.hpp types; without it (this image)
// it uses the API-compatible stand-ins of pdplqr/mini_model.hpp.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../pdplqr.h"
#if defined(PDPLQR_USE_EIGEN) || __has_include(<Eigen/Dense>)
// the reference's own types: Node / LQRModel (lqr_model.hpp:8-89), scalar / VectorXs (typedefs.hpp:8-21) and the
// enum lqr::CondensedSystemSolverType, which lives next to LQRParallelSolver (lqr_solver_parallel.hpp:14-17)
#include "clqr/lqr/lqr_solver_parallel.hpp"
#include "clqr/lqr_model.hpp"
#include "clqr/typedefs.hpp"
#define PDPLQR_REFERENCE_TYPES 1
#else
#include "mini_model.hpp"
#endif

namespace lqr {

#ifndef PDPLQR_REFERENCE_TYPES
// stand-in with the reference's name and values (lqr_solver_parallel.hpp:14-17): LU = 0, CHOLESKY = 1
enum class CondensedSystemSolverType { LU = PDPLQR_CONDENSED_LU, CHOLESKY = PDPLQR_CONDENSED_CHOLESKY };
#endif
static_assert(static_cast<int>(CondensedSystemSolverType::LU) == PDPLQR_CONDENSED_LU &&
                  static_cast<int>(CondensedSystemSolverType::CHOLESKY) == PDPLQR_CONDENSED_CHOLESKY,
              "enum values must match the C ABI");
using CondensedSystemSolverTypeCuda = CondensedSystemSolverType;   // round-1 name, kept for source compatibility

class LQRCudaSolver {
public:
    LQRCudaSolver(const LQRModel& model, int num_segments, bool load_balancing = true,
                  CondensedSystemSolverTypeCuda solver_type = CondensedSystemSolverTypeCuda::CHOLESKY, int device = 0)
        : model_(model), nx_(model.n), nu_(model.m), N_(model.N) {
        int rc = pdplqr_create(&h_, nx_, nu_, N_, model.ncs.data(), 1, num_segments, load_balancing ? 1 : 0,
                               static_cast<int>(solver_type), device);
        if (rc != PDPLQR_OK)
            throw std::runtime_error("pdplqr_create failed with code " + std::to_string(rc) +
                                     (rc == PDPLQR_ERR_CUDA ? " (no usable CUDA device; there is no CPU fallback)" : ""));
        coff_.assign(N_ + 2, 0);
        for (int k = 0; k <= N_; ++k) coff_[k + 1] = coff_[k] + model.ncs[k];
        sync_model();
    }
    ~LQRCudaSolver() { pdplqr_destroy(h_); }
    LQRCudaSolver(const LQRCudaSolver&) = delete;
    LQRCudaSolver& operator=(const LQRCudaSolver&) = delete;

    // Upload E, c, H, h, D of every node (call again after changing the model).
    void sync_model() {
        const int s = nx_ + nu_;
        std::vector<scalar> E((size_t)N_ * nx_ * s), c((size_t)N_ * nx_), H((size_t)N_ * s * s), hv((size_t)N_ * s),
            HN((size_t)nx_ * nx_), hN(nx_), D;
        for (int k = 0; k <= N_; ++k) {
            const Node& nd = model_.nodes[k];
            if (k < N_) {
                copy_n(nd.E.data(), nx_ * s, &E[(size_t)k * nx_ * s]);
                copy_n(nd.c.data(), nx_, &c[(size_t)k * nx_]);
                copy_n(nd.H.data(), s * s, &H[(size_t)k * s * s]);
                copy_n(nd.h.data(), s, &hv[(size_t)k * s]);
            } else {
                copy_n(nd.H.data(), nx_ * nx_, HN.data());
                copy_n(nd.h.data(), nx_, hN.data());
            }
            if (nd.n_con > 0) D.insert(D.end(), nd.D_con.data(), nd.D_con.data() + (size_t)nd.n_con * (k < N_ ? s : nx_));
        }
        check(pdplqr_set_model(h_, E.data(), c.data(), H.data(), hv.data(), HN.data(), hN.data(),
                               D.empty() ? nullptr : D.data()));
    }

    void clear_workspace() {}  // the device workspace needs no clearing (lqr_solver_parallel.hpp:26-32)

    void update_problem_data(const std::vector<VectorXs>& ws, const std::vector<VectorXs>& ys,
                             const std::vector<VectorXs>& zs, const std::vector<VectorXs>& inv_rho_vecs,
                             const scalar sigma) {
        flatten_ws(ws, ws_flat_);
        flatten_con(ys, ys_flat_);
        flatten_con(zs, zs_flat_);
        flatten_con(inv_rho_vecs, ir_flat_);
        const bool con = coff_[N_ + 1] > 0;
        check(pdplqr_update_problem_data(h_, ws_flat_.data(), con ? ys_flat_.data() : nullptr,
                                         con ? zs_flat_.data() : nullptr, con ? ir_flat_.data() : nullptr, sigma));
    }
    void backward(const std::vector<VectorXs>& rho_vecs) {
        flatten_con(rho_vecs, rho_flat_);
        check(pdplqr_backward(h_, coff_[N_ + 1] > 0 ? rho_flat_.data() : nullptr));
    }
    void backward_without_factorization(const std::vector<VectorXs>& rho_vecs) {
        flatten_con(rho_vecs, rho_flat_);
        check(pdplqr_backward_without_factorization(h_, coff_[N_ + 1] > 0 ? rho_flat_.data() : nullptr));
    }
    void forward(const VectorXs& x0, std::vector<VectorXs>& ws) {
        const int s = nx_ + nu_;
        out_flat_.resize((size_t)N_ * s + nx_);
        check(pdplqr_forward(h_, x0.data(), out_flat_.data()));
        for (int k = 0; k <= N_; ++k) {
            const int dim = k < N_ ? s : nx_;
            // ws[N] may have nx or nx+nu entries in caller code (lqr_example.cpp:30-34): write its tail(nx)
            scalar* dst = ws[k].data() + (ws[k].size() - dim);
            copy_n(&out_flat_[(size_t)k * s], dim, dst);
        }
    }
    // Addition: the three calls fused.
    void solve(std::vector<VectorXs>& ws, const std::vector<VectorXs>& ys, const std::vector<VectorXs>& zs,
               const std::vector<VectorXs>& rho_vecs, const std::vector<VectorXs>& inv_rho_vecs, scalar sigma,
               const VectorXs& x0) {
        update_problem_data(ws, ys, zs, inv_rho_vecs, sigma);
        backward(rho_vecs);
        forward(x0, ws);
    }

    // Additions for the conic outer iteration the reference only leaves hooks for (lqr_solver_parallel.hpp:33-37,
    // lqr_model.hpp:21-24).  Every stage's constraint rows become one box cone [e_lb, e_ub] of the node.
    void set_box_cones() {
        std::vector<int> stage, row0, dim, type;
        std::vector<scalar> lb((size_t)coff_[N_ + 1]), ub((size_t)coff_[N_ + 1]);
        for (int k = 0; k <= N_; ++k) {
            const Node& nd = model_.nodes[k];
            if (nd.n_con == 0) continue;
            stage.push_back(k); row0.push_back(0); dim.push_back(nd.n_con); type.push_back(PDPLQR_CONE_BOX);
            for (int r = 0; r < nd.n_con; ++r) {
                lb[(size_t)coff_[k] + r] = clamp_inf(nd.e_lb(r));
                ub[(size_t)coff_[k] + r] = clamp_inf(nd.e_ub(r));
            }
        }
        check(pdplqr_admm_set_cones(h_, (int)stage.size(), stage.data(), row0.data(), dim.data(), type.data(), lb.data(),
                                    ub.data()));
    }
    // ADMM in OSQP form on the flat iterates (ws: N (nx+nu) + nx; zs, ys, rho: sum of ncs); returns the iteration count,
    // residuals[0] = primal, [1] = dual.
    int admm_solve(const VectorXs& x0, std::vector<scalar>& ws, std::vector<scalar>& zs, std::vector<scalar>& ys,
                   const std::vector<scalar>& rho, scalar sigma, scalar alpha, int max_iter, scalar eps_abs,
                   scalar eps_rel, int check_every, scalar residuals[2]) {
        int iters = 0;
        check(pdplqr_admm_solve(h_, x0.data(), ws.data(), zs.data(), ys.data(), rho.data(), sigma, alpha, max_iter,
                                eps_abs, eps_rel, check_every, &iters, residuals));
        return iters;
    }
    size_t ws_len() const { return (size_t)N_ * (nx_ + nu_) + nx_; }
    size_t nc_total() const { return (size_t)coff_[N_ + 1]; }
    const std::vector<long long>& constraint_offsets() const { return coff_; }

    // Accessors (additions).  K_k (nu x nx, column-major), d_k (nu): lqr_kernel_parallel.hpp:105-108.
    int num_segments() const { return pdplqr_num_segments(h_); }
    void gains(std::vector<scalar>& K, std::vector<scalar>& d) {
        K.resize((size_t)N_ * nu_ * nx_);
        d.resize((size_t)N_ * nu_);
        check(pdplqr_get_gains(h_, K.data(), d.data(), nullptr));
    }
    void interface(std::vector<scalar>& xhat, std::vector<scalar>& uhat) {  // condensed_system.hpp:140-146
        xhat.resize((size_t)num_segments() * nx_);
        uhat.resize((size_t)num_segments() * nx_);
        check(pdplqr_get_interface(h_, xhat.data(), uhat.data()));
    }
    // Costates lambda_1 .. lambda_N of the last solve (lam[(k-1) nx + i]); the reference has this step commented out
    // (lqr_kernel.hpp:205-211, lqr_kernel_parallel.hpp:207-216).  `ws` = what forward returned.
    void costates(const std::vector<VectorXs>& ws, std::vector<scalar>& lam) {
        flatten_ws(ws, out_flat_);
        lam.resize((size_t)N_ * nx_);
        check(pdplqr_get_costates(h_, out_flat_.data(), lam.data()));
    }
    bool not_positive_definite() { return pdplqr_last_status(h_, nullptr) > 0; }
    // Debug aid (PDPLQR_DEBUG_GUARDS=1 in the environment when the solver is constructed): guard bytes around the solver's
    // device allocations that were overwritten; 0 = clean, -1 = constructed without guards (include/pdplqr.h).
    long long debug_check_guards() {
        long long n = 0;
        check(pdplqr_debug_check_guards(h_, &n));
        return n;
    }
    pdplqr_handle_t handle() { return h_; }

private:
    static void copy_n(const scalar* src, size_t n, scalar* dst) {
        for (size_t i = 0; i < n; ++i) dst[i] = src[i];
    }
    static scalar clamp_inf(scalar v) { return v > 1e20 ? 1e20 : (v < -1e20 ? -1e20 : v); }
    void check(int rc) {
        if (rc != PDPLQR_OK) throw std::runtime_error(std::string("pdplqr: ") + pdplqr_last_error(h_));
    }
    void flatten_ws(const std::vector<VectorXs>& ws, std::vector<scalar>& out) {
        const int s = nx_ + nu_;
        out.assign((size_t)N_ * s + nx_, 0.0);
        for (int k = 0; k <= N_; ++k) {
            const int dim = k < N_ ? s : nx_;
            copy_n(ws[k].data() + (ws[k].size() - dim), dim, &out[(size_t)k * s]);
        }
    }
    void flatten_con(const std::vector<VectorXs>& v, std::vector<scalar>& out) {
        out.assign((size_t)coff_[N_ + 1], 0.0);
        for (int k = 0; k <= N_; ++k)
            if (model_.ncs[k] > 0) copy_n(v[k].data(), model_.ncs[k], &out[(size_t)coff_[k]]);
    }

    const LQRModel& model_;
    int nx_, nu_, N_;
    pdplqr_handle_t h_ = nullptr;
    std::vector<long long> coff_;
    std::vector<scalar> ws_flat_, ys_flat_, zs_flat_, ir_flat_, rho_flat_, out_flat_;
};

}  // namespace lqr
