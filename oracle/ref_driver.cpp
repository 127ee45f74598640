// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around the REFERENCE'S OWN solver classes, compiled unmodified from /root/reference/include
// (lqr::LQRSolver, lqr_solver.hpp:9-28; lqr::LQRParallelSolver, lqr_solver_parallel.hpp:19-62) against the Eigen-API
// shim in oracle/eigen_shim (Eigen3 itself is absent from this image).  Built by oracle/Makefile into
// oracle/_ref/libpdpref.so; used by tests/test_reference_build.py to pin the oracle port against the reference's
// actual control flow.  No reference source is copied into this repository.
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <vector>
#include <omp.h>
#include <sched.h>

#include "clqr/typedefs.hpp"
#include "clqr/lqr_model.hpp"
#include "clqr/lqr/lqr_solver.hpp"
#include "clqr/lqr/lqr_solver_parallel.hpp"

using namespace lqr;

struct RefHandle {
    std::unique_ptr<LQRModel> model;
    std::unique_ptr<LQRSolver> seq;
    std::unique_ptr<LQRParallelSolver> par;
    std::vector<VectorXs> ws, ys, zs, rho, inv_rho;
    int nx, nu, N;
};

static void fill(std::vector<VectorXs>& v, const double* flat, const LQRModel& m, bool is_ws) {
    size_t off = 0;
    v.resize(m.N + 1);
    for (int k = 0; k <= m.N; ++k) {
        const int len = is_ws ? (k < m.N ? m.n + m.m : m.n) : m.ncs[k];
        v[k].resize(len);
        for (int i = 0; i < len; ++i) v[k](i) = flat ? flat[off + i] : 0.0;
        off += len;
    }
}

extern "C" {

void* ref_create(int nx, int nu, int N, const int* ncs, int parallel, int num_segments, int load_balancing,
                 int condensed_type, const double* E, const double* c, const double* H, const double* h,
                 const double* HN, const double* hN, const double* D) {
    auto* r = new RefHandle();
    r->nx = nx; r->nu = nu; r->N = N;
    r->model.reset(new LQRModel(nx, nu, N));
    const int s = nx + nu;
    size_t doff = 0;
    for (int k = 0; k <= N; ++k) {
        const int nc = ncs ? ncs[k] : 0;
        r->model->add_node(nx, nu, nc, k, k == N);
        Node& nd = r->model->nodes[k];
        const int dim = k < N ? s : nx;
        if (k < N) {
            for (int j = 0; j < s; ++j) for (int i = 0; i < nx; ++i) nd.E(i, j) = E[(size_t)k * nx * s + i + (size_t)j * nx];
            for (int i = 0; i < nx; ++i) nd.c(i) = c[(size_t)k * nx + i];
            for (int j = 0; j < s; ++j) for (int i = 0; i < s; ++i) nd.H(i, j) = H[(size_t)k * s * s + i + (size_t)j * s];
            for (int i = 0; i < s; ++i) nd.h(i) = h[(size_t)k * s + i];
        } else {
            for (int j = 0; j < nx; ++j) for (int i = 0; i < nx; ++i) nd.H(i, j) = HN[i + (size_t)j * nx];
            for (int i = 0; i < nx; ++i) nd.h(i) = hN[i];
        }
        for (int j = 0; j < dim && nc > 0; ++j) for (int i = 0; i < nc; ++i) nd.D_con(i, j) = D[doff + i + (size_t)j * nc];
        doff += (size_t)nc * dim;
    }
    if (parallel)
        r->par.reset(new LQRParallelSolver(*r->model, num_segments, load_balancing != 0,
                                           condensed_type ? CondensedSystemSolverType::CHOLESKY : CondensedSystemSolverType::LU));
    else
        r->seq.reset(new LQRSolver(*r->model));
    return r;
}
void ref_destroy(void* p) { delete static_cast<RefHandle*>(p); }

void ref_update_problem_data(void* p, const double* ws, const double* ys, const double* zs, const double* inv_rho,
                             double sigma) {
    auto* r = static_cast<RefHandle*>(p);
    fill(r->ws, ws, *r->model, true);
    fill(r->ys, ys, *r->model, false);
    fill(r->zs, zs, *r->model, false);
    fill(r->inv_rho, inv_rho, *r->model, false);
    if (r->par) r->par->update_problem_data(r->ws, r->ys, r->zs, r->inv_rho, sigma);
    else r->seq->update_problem_data(r->ws, r->ys, r->zs, r->inv_rho, sigma);
}
void ref_backward(void* p, const double* rho, int factorize) {
    auto* r = static_cast<RefHandle*>(p);
    fill(r->rho, rho, *r->model, false);
    if (r->par) { if (factorize) r->par->backward(r->rho); else r->par->backward_without_factorization(r->rho); }
    else { if (factorize) r->seq->backward(r->rho); else r->seq->backward_without_factorization(r->rho); }
}
void ref_forward(void* p, const double* x0, double* ws_out) {
    auto* r = static_cast<RefHandle*>(p);
    VectorXs x(r->nx);
    for (int i = 0; i < r->nx; ++i) x(i) = x0[i];
    if (r->par) r->par->forward(x, r->ws);
    else r->seq->forward(x, r->ws);
    size_t off = 0;
    for (int k = 0; k <= r->N; ++k) {
        const int len = r->ws[k].size();
        for (int i = 0; i < len; ++i) ws_out[off + i] = r->ws[k](i);
        off += len;
    }
}

}  // extern "C"
