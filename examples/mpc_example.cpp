// Receding-horizon driver for the quadrotor of BASELINE.json config 1 with the box constraints the reference example
// defines but switches off (/root/reference examples/lqr_example.cpp:126-158: `nc = 0;`): inputs in
// [-0.9916, 2.4084] at every stage, the example's state bounds (:63-70) from stage 1 on.  25 control periods
// from rest towards z = 1 m with lqr::RecedingHorizon (include/pdplqr/receding_horizon.hpp): one conic ADMM solve per
// period on the GPU, warm-started with the shifted previous plan.
//   g++ -std=c++17 -Iinclude examples/mpc_example.cpp -Lpdp-lqr_b200 -lpdplqr -Wl,-rpath,$PWD/pdp-lqr_b200 -o mpc_example
#include <cstdio>
#include <vector>

#include "pdplqr/receding_horizon.hpp"

using namespace lqr;

int main() {
    constexpr int nx = 12, nu = 4, N = 20, T = 25;
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

    const double u_lb = -0.9916, u_ub = 2.4084, inf = 1e20;
    double x_lb[nx], x_ub[nx];
    for (int i = 0; i < nx; ++i) { x_lb[i] = -inf; x_ub[i] = inf; }
    x_lb[0] = x_lb[1] = -0.52359878; x_ub[0] = x_ub[1] = 0.52359878;   // lqr_example.cpp:63-70
    x_lb[5] = -1.0;
    x_ub[8] = 2.5;

    LQRModel model(nx, nu, N);
    for (int k = 0; k <= N; ++k) {
        const bool term = (k == N);
        const int nc = term ? nx : (k == 0 ? nu : nx + nu);      // stage 0: the state is given, only inputs are bounded
        model.add_node(nx, nu, nc, k, term);
        Node& nd = model.nodes[k];
        if (!term) {
            for (int i = 0; i < nx; ++i) {
                for (int j = 0; j < nu; ++j) nd.E(i, j) = B[i][j];
                for (int j = 0; j < nx; ++j) nd.E(i, nu + j) = A[i][j];
            }
            for (int j = 0; j < nu; ++j) nd.H(j, j) = 0.1;
            for (int i = 0; i < nx; ++i) { nd.H(nu + i, nu + i) = Qd[i]; nd.h(nu + i) = -xref[i] * Qd[i]; }
            for (int r = 0; r < nc; ++r) {                         // identity rows on [u; x] (on u only at k = 0)
                nd.D_con(r, r) = 1.0;
                nd.e_lb(r) = r < nu ? u_lb : x_lb[r - nu];
                nd.e_ub(r) = r < nu ? u_ub : x_ub[r - nu];
            }
        } else {
            for (int i = 0; i < nx; ++i) {
                nd.H(i, i) = Qd[i]; nd.h(i) = -xref[i] * Qd[i];
                nd.D_con(i, i) = 1.0; nd.e_lb(i) = x_lb[i]; nd.e_ub(i) = x_ub[i];
            }
        }
    }
    VectorXs x0(nx);
    LQRCudaSolver solver(model, 2, true, CondensedSystemSolverTypeCuda::CHOLESKY);
    RecedingHorizon mpc(solver, model, x0, /*rho=*/0.1, /*sigma=*/1e-6, /*alpha=*/1.6, /*max_iter=*/400, /*eps=*/1e-4);
    int total_iters = 0;
    double worst = 0.0;
    for (int t = 0; t < T; ++t) {
        RecedingHorizon::Info info;
        const std::vector<scalar> u = mpc.step(&info);
        total_iters += info.iterations;
        for (int j = 0; j < nu; ++j) {
            const double viol = u[j] > u_ub ? u[j] - u_ub : (u[j] < u_lb ? u_lb - u[j] : 0.0);
            if (viol > worst) worst = viol;
        }
        if (t < 3 || t == T - 1)
            std::printf("period %2d: u0 = %.6f %.6f %.6f %.6f  z = %.6f  admm iterations = %d\n", t, u[0], u[1], u[2], u[3],
                        mpc.state()(2), info.iterations);
    }
    std::printf("height after %d periods: %.6f\n", T, mpc.state()(2));
    std::printf("largest input-bound violation: %.3e\n", worst);
    std::printf("ADMM iterations in total: %d\n", total_iters);
    return solver.not_positive_definite() ? 1 : 0;
}
