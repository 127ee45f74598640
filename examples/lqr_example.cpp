// Our own driver for BASELINE.json config 1: the quadrotor problem the reference ships
// (/root/reference examples/lqr_example.cpp:53-171; data adapted from the OSQP MPC example), solved with
// lqr::LQRCudaSolver exactly as the reference's example drives LQRParallelSolver (:211-221):
//     4 segments, load balancing, Cholesky condensed type, sigma = 1e-6, rho = 0.01, constraints disabled.
// Prints the first five inputs and the final state at full precision.
//   g++ -std=c++17 -Iinclude examples/lqr_example.cpp -Lpdp-lqr_b200 -lpdplqr -Wl,-rpath,$PWD/pdp-lqr_b200 -o lqr_example
#include <cstdio>
#include <vector>

#include "pdplqr/lqr_cuda_solver.hpp"

using namespace lqr;

int main() {
    constexpr int nx = 12, nu = 4, N = 100;
    const double A[nx][nx] = {
        {1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0., 0., 0.},
        {0., 1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0., 0.},
        {0., 0., 1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0.},
        {0.0488, 0., 0., 1., 0., 0., 0.0016, 0., 0., 0.0992, 0., 0.},
        {0., -0.0488, 0., 0., 1., 0., 0., -0.0016, 0., 0., 0.0992, 0.},
        {0., 0., 0., 0., 0., 1., 0., 0., 0., 0., 0., 0.0992},
        {0., 0., 0., 0., 0., 0., 1., 0., 0., 0., 0., 0.},
        {0., 0., 0., 0., 0., 0., 0., 1., 0., 0., 0., 0.},
        {0., 0., 0., 0., 0., 0., 0., 0., 1., 0., 0., 0.},
        {0.9734, 0., 0., 0., 0., 0., 0.0488, 0., 0., 0.9846, 0., 0.},
        {0., -0.9734, 0., 0., 0., 0., 0., -0.0488, 0., 0., 0.9846, 0.},
        {0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.9846}};
    const double B[nx][nu] = {{0., -0.0726, 0., 0.0726},      {-0.0726, 0., 0.0726, 0.},
                              {-0.0152, 0.0152, -0.0152, 0.0152}, {-0., -0.0006, -0., 0.0006},
                              {0.0006, 0., -0.0006, 0.0000},  {0.0106, 0.0106, 0.0106, 0.0106},
                              {0., -1.4512, 0., 1.4512},      {-1.4512, 0., 1.4512, 0.},
                              {-0.3049, 0.3049, -0.3049, 0.3049}, {-0., -0.0236, 0., 0.0236},
                              {0.0236, 0., -0.0236, 0.},      {0.2107, 0.2107, 0.2107, 0.2107}};
    const double Qd[nx] = {0., 0., 10., 10., 10., 10., 0., 0., 0., 5., 5., 5.};
    const double xref[nx] = {0., 0., 1., 0., 0., 0., 0., 0., 0., 0., 0., 0.};

    LQRModel model(nx, nu, N);
    for (int k = 0; k < N; ++k) {
        model.add_node(nx, nu, 0, k);
        Node& nd = model.nodes[k];
        for (int i = 0; i < nx; ++i) {
            for (int j = 0; j < nu; ++j) nd.E(i, j) = B[i][j];
            for (int j = 0; j < nx; ++j) nd.E(i, nu + j) = A[i][j];
        }
        for (int j = 0; j < nu; ++j) nd.H(j, j) = 0.1;
        for (int i = 0; i < nx; ++i) { nd.H(nu + i, nu + i) = Qd[i]; nd.h(nu + i) = -xref[i] * Qd[i]; }
    }
    model.add_node(nx, nu, 0, N, true);
    for (int i = 0; i < nx; ++i) { model.nodes[N].H(i, i) = Qd[i]; model.nodes[N].h(i) = -xref[i] * Qd[i]; }

    std::vector<VectorXs> ws(N + 1), ys(N + 1), zs(N + 1), rho_vecs(N + 1), inv_rho_vecs(N + 1);
    for (int k = 0; k <= N; ++k) ws[k].resize(k < N ? nx + nu : nx);
    VectorXs x0(nx);

    LQRCudaSolver solver(model, 4, true, CondensedSystemSolverTypeCuda::CHOLESKY);
    solver.update_problem_data(ws, ys, zs, inv_rho_vecs, 1e-6);
    solver.backward(rho_vecs);
    solver.forward(x0, ws);
    for (int i = 0; i < 5; ++i) {
        std::printf("Input %d (LQRCudaSolver):", i);
        for (int j = 0; j < nu; ++j) std::printf(" %.12f", ws[i](j));
        std::printf("\n");
    }
    std::printf("Final state (LQRCudaSolver):");
    for (int i = 0; i < nx; ++i) std::printf(" %.12e", ws[N](i));
    std::printf("\n");
    return solver.not_positive_definite() ? 1 : 0;
}
