/* pdplqr.h -- C ABI of the B200-native parallel dynamic-programming LQ solver (libpdplqr.so).
 *
 * This is the drop-in boundary for the reference's OpenMP path.  The reference has no FFI layer: its boundary
 * is the C++ class lqr::LQRParallelSolver (/root/reference include/clqr/lqr/lqr_solver_parallel.hpp:19-62)
 * and the duck-typed 4-call protocol it shares with lqr::LQRSolver (lqr_solver.hpp:9-28).  Every entry point
 * below names the reference interface it replaces.  The header-only C++ host class that restores the
 * reference's own signatures on top of this ABI is include/pdplqr/lqr_cuda_solver.hpp; the binding a
 * maintainer of the reference would add is shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; all matrices FP64 column-major (Eigen's default,
 * typedefs.hpp:8-21); no exceptions cross the ABI; every function returns PDPLQR_OK (0) or a negative error
 * code, and pdplqr_last_error() returns a description.  There is NO CPU fallback: if no CUDA device is
 * usable pdplqr_create fails with PDPLQR_ERR_CUDA.
 *
 * Flat problem layout (a batch of `batch` problems with identical dimensions; batch = 1 is the reference's
 * single-model case), s = nu + nx:
 *   E  [batch][N][nx*s]   E_k = [B_k A_k]                       lqr_model.hpp:12-15
 *   c  [batch][N][nx]
 *   H  [batch][N][s*s]    H_k = [R S; S^T Q] (u block first)    lqr_model.hpp:17-19
 *   h  [batch][N][s]      h_k = [r; q]
 *   HN [batch][nx*nx], hN [batch][nx]    terminal cost          lqr_model.hpp:32-35
 *   D  [batch][sum_k nc_k*dim_k]  D_k = [Du Dx] (nc_k x dim_k), dim_k = s for k < N, nx for k = N
 *                                                                lqr_model.hpp:21-24
 *   ws [batch][N*s + nx]  w_k = [u_k; x_k] (k < N), w_N = x_N   lqr_example.cpp:30-34
 *   ys, zs, rho, inv_rho [batch][sum_k nc_k]                    lqr_solver_parallel.hpp:33-39
 *   x0 [batch][nx]
 */
#ifndef PDPLQR_H
#define PDPLQR_H

#ifdef __cplusplus
extern "C" {
#endif

#define PDPLQR_OK 0
#define PDPLQR_ERR_INVALID (-1)      /* bad argument / dimensions (reference: std::runtime_error, lqr_model.hpp:75-77) */
#define PDPLQR_ERR_UNSUPPORTED (-2)  /* (nx, nu) pair not instantiated in this build */
#define PDPLQR_ERR_CUDA (-3)         /* CUDA runtime failure (incl. no device): no CPU fallback exists */
#define PDPLQR_ERR_ORDER (-4)        /* call-order contract violated (e.g. forward before backward) */
#define PDPLQR_ERR_NOT_PD (-5)       /* a stage Quu was not positive definite (the reference never checks,
                                        lqr_kernel.hpp:89,126; we report it) */

#define PDPLQR_CONDENSED_LU 0        /* CondensedSystemSolverType::LU       lqr_solver_parallel.hpp:14-17 */
#define PDPLQR_CONDENSED_CHOLESKY 1  /* CondensedSystemSolverType::CHOLESKY (both map to the device tree solve) */

typedef struct pdplqr_solver* pdplqr_handle_t;

/* Replaces LQRParallelSolver::LQRParallelSolver(model, num_segments, load_balancing, solver_type)
 * (lqr_solver_parallel.hpp:64-113) and LQRSolver::LQRSolver (lqr_solver.hpp:31-39; num_segments = 1).
 * `ncs` = constraint rows per stage (N+1 entries, LQRModel::ncs lqr_model.hpp:71) or NULL for none.
 * num_segments >= 1 uses the reference's partition rule (lqr_solver_parallel.hpp:70-80): load_balancing = 1 with the
 * 1.55 factor, 0 without; in both the LAST segment takes the remainder of N.  load_balancing = 2 (addition) splits
 * the horizon into num_segments equal parts (the first N % S one stage longer) -- the right choice for thousands of
 * GPU segments.  num_segments = 0 lets the library choose a GPU-appropriate equal segmentation.  `device` = CUDA
 * ordinal. */
int pdplqr_create(pdplqr_handle_t* out, int nx, int nu, int N, const int* ncs, int batch, int num_segments,
                  int load_balancing, int condensed_type, int device);
int pdplqr_destroy(pdplqr_handle_t h);

/* Run all work of this handle on an existing CUDA stream (cudaStream_t passed as void*). */
int pdplqr_set_stream(pdplqr_handle_t h, void* cuda_stream);

/* The reference keeps `const LQRModel&` and re-reads it on every call (lqr_solver_parallel.hpp:52).  A device
 * solver cannot alias host memory, so the model is uploaded explicitly; call again after mutating it.
 * D may be NULL when there are no constraints.  *_device takes device pointers in the same flat layout. */
int pdplqr_set_model(pdplqr_handle_t h, const double* E, const double* c, const double* H, const double* hvec,
                     const double* HN, const double* hN, const double* D);
int pdplqr_set_model_device(pdplqr_handle_t h, const double* E, const double* c, const double* H,
                            const double* hvec, const double* HN, const double* hN, const double* D);

/* Replaces update_problem_data(ws, ys, zs, inv_rho_vecs, sigma)   lqr_solver_parallel.hpp:115-140.
 * The fold-in (H += sigma I, h -= sigma w, g = z - y/rho) is fused into the backward kernels; this call
 * stages the vectors.  ys/zs/inv_rho may be NULL when there are no constraints; ws may be NULL (zeros). */
int pdplqr_update_problem_data(pdplqr_handle_t h, const double* ws, const double* ys, const double* zs,
                               const double* inv_rho, double sigma);
/* Replaces backward(rho_vecs)                        lqr_solver_parallel.hpp:142-146 (reduction + condensed backward) */
int pdplqr_backward(pdplqr_handle_t h, const double* rho);
/* Replaces backward_without_factorization(rho_vecs)  lqr_solver_parallel.hpp:148-154 */
int pdplqr_backward_without_factorization(pdplqr_handle_t h, const double* rho);
/* Replaces forward(x0, ws)                           lqr_solver_parallel.hpp:213-238.  Blocks until ws_out is written. */
int pdplqr_forward(pdplqr_handle_t h, const double* x0, double* ws_out);
/* Addition (north_star "solve()"): update_problem_data + backward + forward in one call. */
int pdplqr_solve(pdplqr_handle_t h, const double* ws_in, const double* ys, const double* zs, const double* rho,
                 const double* inv_rho, double sigma, const double* x0, double* ws_out);

/* Device-pointer variants: arguments are device arrays (same layouts), work is enqueued on the handle's stream
 * and the call returns without synchronising.  Pointers passed to update_problem_data_device are kept, not copied: the
 * arrays must stay valid and UNMODIFIED until the following backward* has run and, if pdplqr_get_costates* is going to be
 * called for this solve, until that call has run (it re-reads ws / ys / zs / rho through the same pointers). */
int pdplqr_update_problem_data_device(pdplqr_handle_t h, const double* ws, const double* ys, const double* zs,
                                      const double* inv_rho, double sigma);
int pdplqr_backward_device(pdplqr_handle_t h, const double* rho);
int pdplqr_backward_without_factorization_device(pdplqr_handle_t h, const double* rho);
int pdplqr_forward_device(pdplqr_handle_t h, const double* x0, double* ws_out);
/* update_problem_data + backward + forward on device arrays as ONE CUDA graph launch (captured on first use and again
 * whenever a pointer, sigma or the stream changes; PDPLQR_SOLVE_GRAPH=0 issues the launches one by one).  Returns without
 * synchronising.  This is the call for a single latency-bound problem (BASELINE.json config 2). */
int pdplqr_solve_device(pdplqr_handle_t h, const double* ws_in, const double* ys, const double* zs, const double* rho,
                        const double* inv_rho, double sigma, const double* x0, double* ws_out);
int pdplqr_synchronize(pdplqr_handle_t h);

/* Options (addition).  PDPLQR_OPT_AFFINE_CACHE: keep per stage Quu^-1, P+c, F+c, F+B during the factorising
 * backward so that backward_without_factorization is an affine-only sweep (the reference always keeps its whole
 * workspace, lqr_kernel.hpp:8-75).  Default: on when the problem has constraints (ADMM use), off otherwise; when
 * off, backward_without_factorization transparently runs the factorising sweep (same result). */
#define PDPLQR_OPT_AFFINE_CACHE 1
/* PDPLQR_OPT_INTERIOR_SHARD (set before set_model): this handle owns a time slice of a longer horizon that does
 * NOT contain the terminal stage, so its last segment ends at an interface like every other segment
 * (lqr_kernel_parallel.hpp:61-65) and HN / hN are ignored. */
#define PDPLQR_OPT_INTERIOR_SHARD 2
int pdplqr_set_option(pdplqr_handle_t h, int option, int value);

/* Horizon sharding across GPUs (addition; the reference's analogue is threads <-> segments with the serial
 * condensed solve in between, lqr_solver_parallel.hpp:144-145,215).  Every rank owns one handle for its time slice
 * (all but the last with PDPLQR_OPT_INTERIOR_SHARD).  After backward each rank exports the summary of its whole
 * slice, `pdplqr_summary_doubles` doubles per problem laid out [P | F | C | p | f] (lqr_solver_parallel.hpp:180-187);
 * the summaries are all-gathered (NCCL), every rank solves the small interface system of the G slices with a
 * coupler and feeds its own entry state / exit costate back with pdplqr_set_root_boundary_device before forward.
 * All pointers are device pointers; work is enqueued on the handle's stream. */
int pdplqr_summary_doubles(pdplqr_handle_t h);
int pdplqr_get_root_summary_device(pdplqr_handle_t h, double* summary);
int pdplqr_set_root_boundary_device(pdplqr_handle_t h, const double* xhat, const double* lam);
int pdplqr_coupler_create(pdplqr_handle_t* out, int nx, int nu, int num_shards, int batch, int device);
/* summaries [batch][G][summary_doubles], x0 [batch][nx] -> xhat, lam [batch][G][nx] (entry state / exit costate
 * of every slice; condensed_system.hpp:140-146).  Destroy the coupler with pdplqr_destroy. */
int pdplqr_coupler_solve_device(pdplqr_handle_t coupler, const double* summaries, const double* x0, double* xhat,
                                double* lam);

/* Single-process horizon sharding over several GPUs (addition): the same flow behind one handle.  The horizon is cut into
 * num_devices contiguous time slices (devices[d] = CUDA ordinal of slice d, NULL = 0 .. num_devices-1), every device
 * reduces its slice, the slice summaries are all-gathered with NCCL (ncclCommInitAll; libnccl.so.2 is loaded at run time),
 * every device solves the interface system redundantly and rolls out its slice.  Host arrays of the FULL horizon in the
 * flat layout above, batch = 1 (one very long problem: BASELINE.json config 5); constraint rows travel with their stages.
 * segments_per_device = 0 lets the library choose.  C++ wrapper: include/pdplqr/lqr_cuda_sharded_solver.hpp. */
typedef struct pdplqr_sharded* pdplqr_sharded_t;
int pdplqr_sharded_create(pdplqr_sharded_t* out, int nx, int nu, int N, const int* ncs, int num_devices, const int* devices,
                          int segments_per_device, int condensed_type);
int pdplqr_sharded_set_model(pdplqr_sharded_t hs, const double* E, const double* c, const double* H, const double* hvec,
                             const double* HN, const double* hN, const double* D);
/* update_problem_data + backward + forward over all devices (lqr_solver_parallel.hpp:115-238); blocks until ws_out is written */
int pdplqr_sharded_solve(pdplqr_sharded_t hs, const double* ws_in, const double* ys, const double* zs, const double* rho,
                         const double* inv_rho, double sigma, const double* x0, double* ws_out);
int pdplqr_sharded_num_devices(pdplqr_sharded_t hs);
const char* pdplqr_sharded_last_error(pdplqr_sharded_t hs);
int pdplqr_sharded_destroy(pdplqr_sharded_t hs);

/* Conic ADMM outer iteration (addition -- NOT IN THE REFERENCE, which ships only the hooks: the ws/ys/zs/rho/
 * inv_rho/sigma arguments, lqr_solver_parallel.hpp:33-37, and Node::D_con/e_lb/e_ub, lqr_model.hpp:21-24, with the
 * example's constraints disabled, lqr_example.cpp:127,158).  OSQP-form ADMM on  D_k w_k in K_k  where every stage's
 * rows are tiled by cones: type 0 = box [e_lb, e_ub], 1 = second-order cone (first row is t, ||rest|| <= t),
 * 2 = ball (||rows|| <= e_ub[first row]).  Cones are listed stage by stage.  One factorising backward, then
 * backward_without_factorization per iteration; everything stays on the device between iterations.
 * ws (warm start in / solution out), zs, ys: host arrays in the layouts above.  residuals_out[2] = primal, dual. */
#define PDPLQR_CONE_BOX 0
#define PDPLQR_CONE_SOC 1
#define PDPLQR_CONE_BALL 2
int pdplqr_admm_set_cones(pdplqr_handle_t h, int ncones, const int* stage, const int* row0, const int* dim,
                          const int* type, const double* e_lb, const double* e_ub);
int pdplqr_admm_solve(pdplqr_handle_t h, const double* x0, double* ws, double* zs, double* ys, const double* rho,
                      double sigma, double alpha, int max_iter, double eps_abs, double eps_rel, int check_every,
                      int* iters_out, double* residuals_out);
/* Same loop on device arrays (w, z, y in/out; inv_rho = 1/rho supplied by the caller).  The whole outer iteration is ONE
 * CUDA graph launch: a factorising iteration followed by a WHILE conditional node whose body is an affine-only
 * iteration (update_problem_data + backward_without_factorization + forward + projections); the convergence test runs on
 * the device every `check_every` iterations and the host reads one 120-byte control block when the graph has finished.
 * residuals_out[1] is the stationarity residual || H w~ + h + D^T y + (dynamics multipliers) ||_inf (admm_kernels.cuh). */
int pdplqr_admm_solve_device(pdplqr_handle_t h, const double* x0, double* w, double* z, double* y, const double* rho,
                             const double* inv_rho, double sigma, double alpha, int max_iter, double eps_abs,
                             double eps_rel, int check_every, int* iters_out, double* residuals_out);

/* use_graph = 0: issue the iterations from a host loop instead (debugging; also PDPLQR_ADMM_GRAPH=0).  adaptive_rho = 1:
 * on a check iteration rho is rescaled by sqrt((r_prim / n_prim) / (r_dual / n_dual)) when that factor leaves
 * [1 / rho_tau, rho_tau] (OSQP's rule; the hooks are rho_vecs / inv_rho_vecs of lqr_solver_parallel.hpp:33-40), at most
 * max_rho_updates times per solve; the next iteration re-factorises (one more graph launch per rescale).  The caller's
 * rho arrays are not modified.  Defaults: graph on, adaptation off, rho_tau 5, 10 updates. */
int pdplqr_admm_configure(pdplqr_handle_t h, int use_graph, int adaptive_rho, double rho_tau, int max_rho_updates);
/* Graph launches so far and rho rescales of the last solve (bench / tests). */
int pdplqr_admm_stats(pdplqr_handle_t h, int* graph_launches, int* rho_updates);

/* Accessors (additions; the reference keeps these in a private workspace, lqr_solver_parallel.hpp:55-60).
 * All outputs are host arrays; any pointer may be NULL to skip it.
 *   partition: starts[S], lens[S]                                   (lqr_solver_parallel.hpp:73-80)
 *   gains:     K [batch][N][nu*nx], d [batch][N][nu]                (lqr_kernel_parallel.hpp:105-108; segment-local
 *              inside non-last segments), Gt [batch][N][nu*nx] = Luu^-T G (lqr_kernel_parallel.hpp:127-128,198)
 *   interface: xhat, uhat [batch][S][nx]                            (condensed_system.hpp:140-146)
 *   summaries: P,F,C [batch][S][nx*nx], p,f [batch][S][nx]          (lqr_solver_parallel.hpp:180-187) */
int pdplqr_num_segments(pdplqr_handle_t h);
int pdplqr_get_partition(pdplqr_handle_t h, int* starts, int* lens);
int pdplqr_get_gains(pdplqr_handle_t h, double* K, double* d, double* Gt);
int pdplqr_get_interface(pdplqr_handle_t h, double* xhat, double* uhat);
int pdplqr_get_summaries(pdplqr_handle_t h, double* P, double* p, double* F, double* f, double* C);
/* Costates of the last solve: lam [batch][N][nx], lam[b][k-1] = lambda_k, the multiplier of
 * x_k = A x_{k-1} + B u_{k-1} + c_{k-1} (k = 1..N) with L = cost + sum_k lambda_{k+1}^T (E_k w_k + c_k - x_{k+1}), for
 * the ADMM-augmented data of the last update_problem_data / backward*.  This is the step the reference has written
 * but commented out (lqr_kernel.hpp:205-211: lambda+ = Lxx+ (Lxx+^T x+) + p+; lqr_kernel_parallel.hpp:207-216 adds
 * F+^T uhat); here it is recovered segment-parallel from the stage stationarity conditions, starting at the interface
 * costates.  `ws` is the trajectory returned by the last forward; call after forward, before
 * the next update_problem_data with different vectors (the device variant reads the ws / ys / zs / rho arrays of the
 * last update through the pointers it was given; after pdplqr_admm_solve* those are the loop's final iterates, which
 * have moved on by one relaxation / projection step from the ones the last LQ solve used).  Works on both the segment
 * path and the thread-per-problem path (batch of tiny systems with num_segments = 1). */
int pdplqr_get_costates(pdplqr_handle_t h, const double* ws, double* lam);
int pdplqr_get_costates_device(pdplqr_handle_t h, const double* ws, double* lam);
/* Number of problems whose last factorising backward met a non-positive pivot; per-problem codes (0 = ok,
 * else 1 + stage index) are written to `status` ([batch], may be NULL). */
int pdplqr_last_status(pdplqr_handle_t h, int* status);
const char* pdplqr_last_error(pdplqr_handle_t h);
/* Kernels launched by this handle so far (bench.py's gpu_launches). */
long long pdplqr_launch_count(pdplqr_handle_t h);
/* Bytes of one device model record / factor record (DESIGN.md data layout), for roofline bookkeeping. */
int pdplqr_record_doubles(pdplqr_handle_t h, int* model_rec, int* factor_rec);
/* (problem, segment) groups one GPU keeps resident in the throughput-mode stage sweep (SMs x resident CTAs per SM of the
 * stage kernel for this (nx, nu), from the CUDA occupancy calculator): a segment count that is a whole multiple of it
 * avoids a trailing partial wave.  num_segments = 0 in pdplqr_create uses it. */
int pdplqr_wave_size(int nx, int nu, int device);
/* Debug aid (addition; the memcheck substitute where compute-sanitizer is not available).  With PDPLQR_DEBUG_GUARDS=1 in the
 * environment at pdplqr_create, every device allocation of the handle is pre-filled with 0xFF bytes (a double read before the
 * library wrote it is a NaN and surfaces in the results) and placed between two 4 KiB guard bands; this call synchronises
 * the handle's stream and counts guard bytes that were overwritten (out-of-bounds device writes) into *corrupted_bytes
 * (0 = clean, -1 = the handle was created without guards).
 * Self-test: called with *corrupted_bytes == PDPLQR_GUARD_SELF_TEST the library first writes 3 bytes just past its first
 * allocation, so the call must report 3 (the handle should be destroyed afterwards). */
#define PDPLQR_GUARD_SELF_TEST (-12345LL)
int pdplqr_debug_check_guards(pdplqr_handle_t h, long long* corrupted_bytes);
int pdplqr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PDPLQR_H */
