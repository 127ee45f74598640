// Receding-horizon (MPC) driver over lqr::LQRCudaSolver (the reference ships no outer
// loop: examples/lqr_example.cpp solves one horizon, lqr_solver_parallel.hpp:33-37 are the hooks an outer loop calls).
// Every period: conic LQ solve from the measured state, warm-started with the previous solution shifted by one stage;
// the first control is applied to the plant x+ = A_0 x + B_0 u + c_0 (node 0 of the model).
// Same logic as pdp-lqr_b200/mpc.py.
#pragma once
#include <vector>

#include "lqr_cuda_solver.hpp"

namespace lqr {

// stage k <- stage k+1 (k < N-1); the last running stage repeats its control from the old terminal state; constraint
// blocks move with their stage when both stages have the same number of rows, otherwise they stay
inline void shift_warm_start(int nx, int nu, int N, const std::vector<long long>& coff, std::vector<scalar>& ws,
                             std::vector<scalar>& zs, std::vector<scalar>& ys) {
    const int s = nx + nu;
    for (size_t e = 0; e + s < (size_t)N * s; ++e) ws[e] = ws[e + s];
    for (int i = 0; i < nx; ++i) ws[(size_t)(N - 1) * s + nu + i] = ws[(size_t)N * s + i];
    for (int k = 0; k + 2 <= N; ++k) {   // running stages 0 .. N-2
        const long long nk = coff[k + 1] - coff[k], nk1 = coff[k + 2] - coff[k + 1];
        if (nk != nk1) continue;
        for (long long r = 0; r < nk; ++r) {
            zs[(size_t)(coff[k] + r)] = zs[(size_t)(coff[k + 1] + r)];
            ys[(size_t)(coff[k] + r)] = ys[(size_t)(coff[k + 1] + r)];
        }
    }
}

class RecedingHorizon {
public:
    struct Info { int iterations; scalar r_prim, r_dual; };

    RecedingHorizon(LQRCudaSolver& solver, const LQRModel& model, const VectorXs& x0, scalar rho = 0.1,
                    scalar sigma = 1e-6, scalar alpha = 1.6, int max_iter = 200, scalar eps = 1e-4, int check_every = 10,
                    bool warm_start = true)
        : sol_(solver), model_(model), nx_(model.n), nu_(model.m), N_(model.N), x_(x0), sigma_(sigma), alpha_(alpha),
          max_iter_(max_iter), eps_(eps), check_every_(check_every), warm_(warm_start) {
        ws_.assign(sol_.ws_len(), 0.0);
        zs_.assign(sol_.nc_total(), 0.0);
        ys_.assign(sol_.nc_total(), 0.0);
        rho_.assign(sol_.nc_total(), rho);
        sol_.set_box_cones();
    }
    const VectorXs& state() const { return x_; }
    const std::vector<scalar>& plan() const { return ws_; }

    // one control period: returns u_0 (nu values) and advances the plant
    std::vector<scalar> step(Info* info = nullptr) {
        if (warm_ && periods_ > 0) shift_warm_start(nx_, nu_, N_, sol_.constraint_offsets(), ws_, zs_, ys_);
        else {
            ws_.assign(ws_.size(), 0.0); zs_.assign(zs_.size(), 0.0); ys_.assign(ys_.size(), 0.0);
        }
        scalar res[2] = {0.0, 0.0};
        const int it = sol_.admm_solve(x_, ws_, zs_, ys_, rho_, sigma_, alpha_, max_iter_, eps_, eps_, check_every_, res);
        if (info) *info = Info{it, res[0], res[1]};
        std::vector<scalar> u(ws_.begin(), ws_.begin() + nu_);
        const Node& n0 = model_.nodes[0];
        VectorXs xn(nx_);
        for (int i = 0; i < nx_; ++i) {
            scalar acc = n0.c(i);
            for (int j = 0; j < nu_; ++j) acc += n0.E(i, j) * u[j];
            for (int j = 0; j < nx_; ++j) acc += n0.E(i, nu_ + j) * x_(j);
            xn(i) = acc;
        }
        x_ = xn;
        ++periods_;
        return u;
    }

private:
    LQRCudaSolver& sol_;
    const LQRModel& model_;
    int nx_, nu_, N_;
    VectorXs x_;
    scalar sigma_, alpha_;
    int max_iter_;
    scalar eps_;
    int check_every_;
    bool warm_;
    int periods_ = 0;
    std::vector<scalar> ws_, zs_, ys_, rho_;
};

}  // namespace lqr
