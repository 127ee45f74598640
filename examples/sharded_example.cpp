// The quadrotor problem of examples/lqr_example.cpp with a longer horizon, solved with lqr::LQRCudaShardedSolver on every
// GPU of the box (or argv[1] of them) and with lqr::LQRCudaSolver on one GPU; prints the largest difference.
//   g++ -std=c++17 -Iinclude examples/sharded_example.cpp -Lpdp-lqr_b200 -lpdplqr -Wl,-rpath,$PWD/pdp-lqr_b200 -o sharded_example
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pdplqr/lqr_cuda_sharded_solver.hpp"

using namespace lqr;

int main(int argc, char** argv) {
    constexpr int nx = 12, nu = 4;
    const int G = argc > 1 ? std::atoi(argv[1]) : 2;
    const int N = argc > 2 ? std::atoi(argv[2]) : 4096;
    const double A[nx][nx] = {
        {1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0., 0., 0.},        {0., 1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0., 0.},
        {0., 0., 1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0.},        {0.0488, 0., 0., 1., 0., 0., 0.0016, 0., 0., 0.0992, 0., 0.},
        {0., -0.0488, 0., 0., 1., 0., 0., -0.0016, 0., 0., 0.0992, 0.}, {0., 0., 0., 0., 0., 1., 0., 0., 0., 0., 0., 0.0992},
        {0., 0., 0., 0., 0., 0., 1., 0., 0., 0., 0., 0.},         {0., 0., 0., 0., 0., 0., 0., 1., 0., 0., 0., 0.},
        {0., 0., 0., 0., 0., 0., 0., 0., 1., 0., 0., 0.},         {0.9734, 0., 0., 0., 0., 0., 0.0488, 0., 0., 0.9846, 0., 0.},
        {0., -0.9734, 0., 0., 0., 0., 0., -0.0488, 0., 0., 0.9846, 0.}, {0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.9846}};
    const double B[nx][nu] = {{0., -0.0726, 0., 0.0726},      {-0.0726, 0., 0.0726, 0.},
                              {-0.0152, 0.0152, -0.0152, 0.0152}, {-0., -0.0006, -0., 0.0006},
                              {0.0006, 0., -0.0006, 0.0000},  {0.0106, 0.0106, 0.0106, 0.0106},
                              {0., -1.4512, 0., 1.4512},      {-1.4512, 0., 1.4512, 0.},
                              {-0.3049, 0.3049, -0.3049, 0.3049}, {-0., -0.0236, 0., 0.0236},
                              {0.0236, 0., -0.0236, 0.},      {0.2107, 0.2107, 0.2107, 0.2107}};
    const double Qd[nx] = {0., 0., 10., 10., 10., 10., 0., 0., 0., 5., 5., 5.};
    LQRModel model(nx, nu, N);
    for (int k = 0; k < N; ++k) {
        model.add_node(nx, nu, 0, k);
        Node& nd = model.nodes[k];
        for (int i = 0; i < nx; ++i) {
            for (int j = 0; j < nu; ++j) nd.E(i, j) = B[i][j];
            for (int j = 0; j < nx; ++j) nd.E(i, nu + j) = A[i][j] * (1.0 + 1e-4 * std::sin(0.37 * k + i + 3 * j));   // LTV
            nd.c(i) = 1e-3 * std::cos(0.11 * k + i);
        }
        for (int j = 0; j < nu; ++j) nd.H(j, j) = 0.1;
        for (int i = 0; i < nx; ++i) { nd.H(nu + i, nu + i) = Qd[i]; nd.h(nu + i) = -(i == 2 ? 1.0 : 0.0) * Qd[i]; }
    }
    model.add_node(nx, nu, 0, N, true);
    for (int i = 0; i < nx; ++i) { model.nodes[N].H(i, i) = Qd[i]; model.nodes[N].h(i) = -(i == 2 ? 1.0 : 0.0) * Qd[i]; }

    std::vector<VectorXs> ws(N + 1), ws1(N + 1), ys(N + 1), zs(N + 1), rho_vecs(N + 1), inv_rho_vecs(N + 1);
    for (int k = 0; k <= N; ++k) { ws[k].resize(k < N ? nx + nu : nx); ws1[k].resize(k < N ? nx + nu : nx); }
    VectorXs x0(nx);
    x0(0) = 0.3; x0(5) = -0.2;

    LQRCudaShardedSolver sharded(model, G);
    sharded.update_problem_data(ws, ys, zs, inv_rho_vecs, 1e-6);
    sharded.backward(rho_vecs);
    sharded.forward(x0, ws);

    LQRCudaSolver single(model, 0);
    single.update_problem_data(ws1, ys, zs, inv_rho_vecs, 1e-6);
    single.backward(rho_vecs);
    single.forward(x0, ws1);

    double worst = 0.0, scale = 0.0;
    for (int k = 0; k <= N; ++k)
        for (int i = 0; i < ws[k].size(); ++i) {
            worst = std::fmax(worst, std::fabs(ws[k](i) - ws1[k](i)));
            scale = std::fmax(scale, std::fabs(ws1[k](i)));
        }
    std::printf("devices %d N %d max |sharded - single| / max |single| = %.3e\n", sharded.num_devices(), N, worst / scale);
    return worst / scale < 1e-9 ? 0 : 1;
}
