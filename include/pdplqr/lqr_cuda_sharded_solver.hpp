// lqr::LQRCudaShardedSolver -- the reference's solver protocol for ONE very long horizon spread over several GPUs of one
// process (C ABI: pdplqr_sharded_* in include/pdplqr.h; NCCL all-gather of the slice summaries over NVLink).
// Same constructor shape and the same four calls as lqr::LQRParallelSolver
// (/root/reference include/clqr/lqr/lqr_solver_parallel.hpp:19-62): here "segments" are first the devices' time slices and,
// inside every device, the GPU segments of LQRCudaSolver.  The three calls of a solve are staged on the host and run
// together inside forward() (the device sweeps of all slices, the one collective and the rollouts form one pipeline).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "lqr_cuda_solver.hpp"   // model types (the reference's own with Eigen, stand-ins otherwise) + the C ABI

namespace lqr {

class LQRCudaShardedSolver {
public:
    LQRCudaShardedSolver(const LQRModel& model, int num_devices, int segments_per_device = 0,
                         CondensedSystemSolverType solver_type = CondensedSystemSolverType::CHOLESKY,
                         const std::vector<int>& devices = {})
        : model_(model), nx_(model.n), nu_(model.m), N_(model.N) {
        int rc = pdplqr_sharded_create(&h_, nx_, nu_, N_, model.ncs.data(), num_devices,
                                       devices.empty() ? nullptr : devices.data(), segments_per_device,
                                       static_cast<int>(solver_type));
        if (rc != PDPLQR_OK)
            throw std::runtime_error("pdplqr_sharded_create failed with code " + std::to_string(rc) +
                                     (rc == PDPLQR_ERR_CUDA ? " (not enough CUDA devices, or NCCL missing; there is no CPU fallback)" : ""));
        coff_.assign(N_ + 2, 0);
        for (int k = 0; k <= N_; ++k) coff_[k + 1] = coff_[k] + model.ncs[k];
        sync_model();
    }
    ~LQRCudaShardedSolver() { pdplqr_sharded_destroy(h_); }
    LQRCudaShardedSolver(const LQRCudaShardedSolver&) = delete;
    LQRCudaShardedSolver& operator=(const LQRCudaShardedSolver&) = delete;

    void sync_model() {   // upload every node's E, c, H, h, D to the device that owns its stage
        const int s = nx_ + nu_;
        std::vector<scalar> E((size_t)N_ * nx_ * s), c((size_t)N_ * nx_), H((size_t)N_ * s * s), hv((size_t)N_ * s),
            HN((size_t)nx_ * nx_), hN(nx_), D;
        for (int k = 0; k <= N_; ++k) {
            const Node& nd = model_.nodes[k];
            if (k < N_) {
                copy_n(nd.E.data(), (size_t)nx_ * s, &E[(size_t)k * nx_ * s]);
                copy_n(nd.c.data(), nx_, &c[(size_t)k * nx_]);
                copy_n(nd.H.data(), (size_t)s * s, &H[(size_t)k * s * s]);
                copy_n(nd.h.data(), s, &hv[(size_t)k * s]);
            } else {
                copy_n(nd.H.data(), (size_t)nx_ * nx_, HN.data());
                copy_n(nd.h.data(), nx_, hN.data());
            }
            if (nd.n_con > 0) D.insert(D.end(), nd.D_con.data(), nd.D_con.data() + (size_t)nd.n_con * (k < N_ ? s : nx_));
        }
        check(pdplqr_sharded_set_model(h_, E.data(), c.data(), H.data(), hv.data(), HN.data(), hN.data(),
                                       D.empty() ? nullptr : D.data()));
    }

    void update_problem_data(const std::vector<VectorXs>& ws, const std::vector<VectorXs>& ys,
                             const std::vector<VectorXs>& zs, const std::vector<VectorXs>& inv_rho_vecs,
                             const scalar sigma) {
        flatten_ws(ws, ws_flat_);
        flatten_con(ys, ys_flat_);
        flatten_con(zs, zs_flat_);
        flatten_con(inv_rho_vecs, ir_flat_);
        sigma_ = sigma;
    }
    void backward(const std::vector<VectorXs>& rho_vecs) { flatten_con(rho_vecs, rho_flat_); }
    void forward(const VectorXs& x0, std::vector<VectorXs>& ws) {
        const int s = nx_ + nu_;
        const bool con = coff_[N_ + 1] > 0;
        out_flat_.resize((size_t)N_ * s + nx_);
        check(pdplqr_sharded_solve(h_, ws_flat_.data(), con ? ys_flat_.data() : nullptr, con ? zs_flat_.data() : nullptr,
                                   con ? rho_flat_.data() : nullptr, con ? ir_flat_.data() : nullptr, sigma_, x0.data(),
                                   out_flat_.data()));
        for (int k = 0; k <= N_; ++k) {
            const int dim = k < N_ ? s : nx_;
            copy_n(&out_flat_[(size_t)k * s], dim, ws[k].data() + (ws[k].size() - dim));
        }
    }
    int num_devices() const { return pdplqr_sharded_num_devices(h_); }

private:
    static void copy_n(const scalar* src, size_t n, scalar* dst) {
        for (size_t i = 0; i < n; ++i) dst[i] = src[i];
    }
    void check(int rc) {
        if (rc != PDPLQR_OK) throw std::runtime_error(std::string("pdplqr (sharded): ") + pdplqr_sharded_last_error(h_));
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
    pdplqr_sharded_t h_ = nullptr;
    scalar sigma_ = 0.0;
    std::vector<long long> coff_;
    std::vector<scalar> ws_flat_, ys_flat_, zs_flat_, ir_flat_, rho_flat_, out_flat_;
};

}  // namespace lqr
