// Conic ADMM outer iteration on top of the LQ solve.
// NOT IN THE REFERENCE: the reference ships only the hooks -- the ws/ys/zs/rho/inv_rho/sigma arguments of
// update_problem_data (lqr_solver_parallel.hpp:33-37) and Node::D_con, e_lb, e_ub (lqr_model.hpp:21-24, never read by
// any solver); its example disables the constraints (lqr_example.cpp:127,158).  The iteration implemented here is
// the OSQP-form ADMM those hooks are shaped for, with cone projections instead of box clamps only:
//     w~      = LQ solve with  H + sigma I + D^T rho D ,  h - sigma w - D^T (rho o (z - y/rho))     (the hot path)
//     z~      = D w~
//     w       = alpha w~ + (1 - alpha) w_prev ,      z^ = alpha z~ + (1 - alpha) z_prev
//     z       = Proj_K(z^ + y / rho)                  K = product of boxes [e_lb, e_ub], second-order cones, balls
//     y       = y + rho o (z^ - z)
//     r_prim  = || z~ - z ||_inf
//     r_dual  = || H w~ + h + D^T y + (dynamics multipliers) ||_inf , the stationarity residual of the conic problem at
//               (w~, y, lambda).  The LQ solve makes the AUGMENTED stationarity exact, so by subtraction
//               r_dual = || sigma (w~ - w_prev) + D^T (rho o ((1 - alpha)(z~ - z_prev) + (z - z_prev))) ||_inf
//               -- all stage-local quantities, no second pass over H (tests check it against the definition evaluated with
//               the costates of pdplqr_get_costates / the KKT multipliers)
//     rho adaptation (OSQP rule, optional): on a check iteration rho *= sqrt((r_prim / n_prim) / (r_dual / n_dual)) when that
//               factor leaves [1/tau, tau]; the next iteration then re-factorises
// One warp per (problem, stage) item, lanes over constraint rows for the mat-vec and over cones for the projections;
// the warps of a fixed-size grid loop over the items (a CTA per item made the kernel CTA-launch-rate bound: 2.9 ns per
// 32-thread CTA, 3.1 ms per iteration at C4 against 0.6 ms of HBM time).
// Parity is pinned against an independent numpy restatement of the same iteration kept with the test infrastructure
// ("parity unpinned" by the reference by construction).
#pragma once
#include "common.cuh"

namespace pdplqr {

enum ConeType { CONE_BOX = 0, CONE_SOC = 1, CONE_BALL = 2 };

struct AdmmParams {
    int nx, nu, N, batch, ncmax;
    const int* ncs;            // [N+1]
    const long long* coff;     // [N+2]
    const long long* doff;     // [N+2] (padded device layout of D)
    const double* Dm;          // [batch][d_total]
    long long d_total, nc_total;
    const int* sel_col;        // selection-matrix constraints (see SegParams): [batch][nc_total] or nullptr (dense D)
    const double* sel_val;
    // cones: per stage k the cones cone_first[k] .. cone_first[k+1]-1 ; each (type, first row within stage, dim)
    const int* cone_first;     // [N+2]
    const int* cone_type;
    const int* cone_row;
    const int* cone_dim;
    const int* row_box;        // [nc_total] (shared by the batch): 1 = the row belongs to a box cone
    const double* e_lb;        // [batch][nc_total]   box bounds (ball: radius in e_ub of the first row)
    const double* e_ub;
    const double* w_tilde;     // [batch][ws_len]   LQ solution of this iteration
    double* w;                 // [batch][ws_len]   in: w_prev, out: relaxed iterate
    double* z;                 // [batch][nc_total] in: z_prev, out: z
    double* y;                 // [batch][nc_total] in/out
    const double* rho;         // [batch][nc_total]
    double alpha, sigma;
    struct AdmmCtl* ctl;       // device-resident loop state: decides whether this iteration computes the residual norms
};

// Loop state of one conic solve, resident on the device: the whole outer iteration runs as ONE CUDA graph launch (a
// factorising iteration followed by a WHILE conditional node whose body is an affine-only iteration); admm_ctl_kernel
// ends every iteration, tests convergence on check iterations and sets the WHILE condition.
struct AdmmCtl {
    int iter;                  // iterations completed
    int max_iter, check_every;
    int converged;             // set on a check iteration
    int rho_update;            // the loop stopped because rho should be rescaled by rho_scale (host relaunches)
    int adaptive, n_rho_updates, max_rho_updates;
    int cont;                  // last WHILE condition (read by the host-loop fallback)
    int pad_;
    double eps_abs, eps_rel, rho_tau, rho_scale;
    double res[4];             // residuals of the last check iteration: r_prim, r_dual, n_prim, n_dual
    unsigned long long acc[4]; // running maxima of the current check iteration (bit patterns of non-negative doubles)
};

PDPLQR_DEVINL bool admm_is_check(const AdmmCtl* c) {
    const int it = c->iter + 1;
    return (it % c->check_every == 0) || it >= c->max_iter;
}

PDPLQR_DEVINL void atomic_max_nonneg(unsigned long long* addr, double v) {
    // non-negative IEEE doubles order like their bit patterns
    atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

// Tried and dropped (measured at C4, 4096 x 257 items, 50 iterations): an L2 (or L1) prefetch of everything the next item
// of the warp will read, one 128-byte line per lane, issued at the top of an item: 356 -> 374 ms per solve (376 with L1, the
// same with 6 instead of 8 CTAs per SM).  With 64 resident warps per SM the loads are covered already; the kernel is bound by
// issue slots (ncu: 70 % issue-active), and the prefetch adds 250 instructions per item.
// Also tried on the lean kernel below (590 instructions per item, where ncu shows long-scoreboard stalls at the sel_col ->
// wt[cj], w, z and bound loads one after the other): all loads of a row -- and the first 32 rows' together with w~ and w --
// issued before the first use.  The 14 more live registers spill at the 32 the 8-CTA residency leaves (C4 349 -> 385 ms), and
// with 40 / 48 / 64 registers (6 / 5 / 4 CTAs per SM) it measured 373 / 361 / 349 ms: what the single round trip gains, the
// lost warps give back.
constexpr int ADMM_WARPS = 8;   // warps per CTA of admm_update_kernel; 8 CTAs per SM (32 registers): the kernel is a chain of
                                // dependent global loads per item, so resident warps are what hides the latency (C4 453 -> 422 ms)

// ORDINARY iterations (no residual norms).  Enqueued every iteration next to admm_update_res_kernel; exactly one of the two
// does the work, decided on the device from the loop state, so the iteration can sit in a CUDA graph unchanged.  Kept
// separate (not a template flag on one body): with the residual code in the same kernel the hot loop went from 1.3 to
// 2.5 ms per iteration at C4.
#ifndef PDPLQR_ADMM_HOIST
#define PDPLQR_ADMM_HOIST 0      // 1: the first 32 rows' loads issued together with w~ (measured: 338 vs 308 ms per C4 solve at 8
#endif                           // CTAs per SM, 326 / 310 ms at 6 / 4 -- the extra live registers spill or cost residency)
#ifndef PDPLQR_ADMM_MINB
#define PDPLQR_ADMM_MINB 8
#endif
// ORDINARY iterations, part 1: relaxation w = alpha w~ + (1 - alpha) w.  No per-stage structure: a flat coalesced pass of its
// own (inside the per-stage kernel it cost that kernel its registers; the rows below need w~ only).
__global__ void __launch_bounds__(256) admm_relax_kernel(AdmmParams p) {
    if (admm_is_check(p.ctl)) return;
    const double alpha = p.alpha, oma = 1.0 - p.alpha;
    const long long nw = (long long)p.batch * ((long long)p.N * (p.nx + p.nu) + p.nx);
    const long long tstride = (long long)gridDim.x * blockDim.x;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; idx + 3 * tstride < nw; idx += 4 * tstride) {
        const double a0 = p.w_tilde[idx], a1 = p.w_tilde[idx + tstride], a2 = p.w_tilde[idx + 2 * tstride],
                     a3 = p.w_tilde[idx + 3 * tstride];
        const double w0 = p.w[idx], w1 = p.w[idx + tstride], w2 = p.w[idx + 2 * tstride], w3 = p.w[idx + 3 * tstride];
        p.w[idx] = alpha * a0 + oma * w0;
        p.w[idx + tstride] = alpha * a1 + oma * w1;
        p.w[idx + 2 * tstride] = alpha * a2 + oma * w2;
        p.w[idx + 3 * tstride] = alpha * a3 + oma * w3;
    }
    for (; idx < nw; idx += tstride) p.w[idx] = alpha * p.w_tilde[idx] + oma * p.w[idx];
}

__global__ void __launch_bounds__(ADMM_WARPS * 32, PDPLQR_ADMM_MINB) admm_update_kernel(AdmmParams p) {
    if (admm_is_check(p.ctl)) return;
    // ncu (profiles/r2_ncu_c4_kernels.txt): this kernel is ISSUE-bound (70 % of the issue slots at 64 resident warps per SM,
    // 900 warp-instructions per item), not latency-bound.  So: a box row -- the common case -- is finished in the pass that
    // forms it, in registers (one trip through z, y, rho, the bounds; no shared-memory staging, no second pass); only rows of
    // second-order cones / balls go through shared memory; (problem, stage) advance incrementally (no 64-bit divisions per
    // item); y / rho by a Newton reciprocal.
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = p.nx + p.nu, N1 = p.N + 1;
    const size_t ws_len = (size_t)p.N * s + p.nx;
    double* wt = smem + (size_t)warp * (s + 3 * p.ncmax);   // w~ of this stage   (dim)
    double* v = wt + s;                // z^ + y/rho, then its projection   (rows of non-box cones)
    double* zt = v + p.ncmax;          // z^                                (rows of non-box cones)
    const double alpha = p.alpha, oma = 1.0 - p.alpha;
    const long long items = (long long)p.batch * N1;
    const long long stride = (long long)gridDim.x * ADMM_WARPS;
    const int dk = (int)(stride % N1);
    const long long db = stride / N1;
    long long item = (long long)blockIdx.x * ADMM_WARPS + warp;
    int k = (int)(item % N1);
    long long b = item / N1;
    const bool sel = p.sel_col != nullptr;
#pragma unroll 1
    for (; item < items; item += stride) {
        const int dim = (k < p.N) ? s : p.nx;
        const int nc = p.ncs[k];
        const long long cok = p.coff[k];
        const size_t wo = (size_t)b * ws_len + (size_t)k * s;
        const size_t co = (size_t)b * p.nc_total + cok;
        const double* Dk = p.Dm + (size_t)b * p.d_total + p.doff[k];
        const int c0 = p.cone_first[k], c1 = p.cone_first[k + 1];
        k += dk; b += db;              // (problem, stage) of the next item
        if (k >= N1) { k -= N1; ++b; }
        if (nc == 0) continue;         // (the relaxation of w is the flat pass above)
        struct Row { int cj, box; double sv, zold, yold, rr, lb, ub; };
        auto load_row = [&](int r, Row& q) {
            q.cj = sel ? p.sel_col[co + r] : -1;
            q.sv = sel ? p.sel_val[co + r] : 0.0;
            q.zold = p.z[co + r]; q.yold = p.y[co + r]; q.rr = p.rho[co + r];
            q.box = p.row_box[cok + r];
            q.lb = p.e_lb[co + r]; q.ub = p.e_ub[co + r];
        };
        auto finish = [&](int r, double zh, double znew, double yold, double rr) {
            p.z[co + r] = znew;
            p.y[co + r] = yold + rr * (zh - znew);
        };
        auto do_row = [&](int r, const Row& q) {
            double acc = 0.0;
            if (sel) {
                if (q.cj >= 0) acc = q.sv * wt[q.cj];
            } else {
                for (int j = 0; j < dim; ++j) acc = fma(Dk[r + (size_t)j * nc], wt[j], acc);
            }
            const double zh = alpha * acc + oma * q.zold;
            const double vv = fma(q.yold, rcp_newton(q.rr), zh);
            if (q.box) {
                finish(r, zh, fmin(fmax(vv, q.lb), q.ub), q.yold, q.rr);
            } else {
                v[r] = vv;
                zt[r] = zh;
            }
        };
        Row q0;
        if (PDPLQR_ADMM_HOIST && lane < nc) load_row(lane, q0);   // in flight together with w~ (one round trip for the first 32 rows)
        __syncwarp();                  // the previous item's readers of the scratch are done
        for (int i = lane; i < dim; i += 32) wt[i] = p.w_tilde[wo + i];
        __syncwarp();
        if (PDPLQR_ADMM_HOIST && lane < nc) do_row(lane, q0);
        for (int r = lane + (PDPLQR_ADMM_HOIST ? 32 : 0); r < nc; r += 32) {
            Row q;
            load_row(r, q);
            do_row(r, q);
        }
        auto project_serial = [&](int r0, int d, int type) {     // one lane, whole cone (second-order cone or ball)
            if (type == CONE_SOC) {
                double nv = 0.0;
                for (int r = r0 + 1; r < r0 + d; ++r) nv = fma(v[r], v[r], nv);
                nv = sqrt(nv);
                const double t = v[r0];
                if (nv <= t) { /* inside */ }
                else if (nv <= -t) { for (int r = r0; r < r0 + d; ++r) v[r] = 0.0; }
                else {
                    const double a = 0.5 * (t + nv), sc = a / nv;
                    v[r0] = a;
                    for (int r = r0 + 1; r < r0 + d; ++r) v[r] *= sc;
                }
            } else {  // ball of radius e_ub[first row]
                double nv = 0.0;
                for (int r = r0; r < r0 + d; ++r) nv = fma(v[r], v[r], nv);
                nv = sqrt(nv);
                const double rad = p.e_ub[co + r0];
                if (nv > rad) { const double sc = rad / nv; for (int r = r0; r < r0 + d; ++r) v[r] *= sc; }
            }
        };
        if (c1 - c0 <= 8) {            // few cones per stage: the warp walks them together
            for (int c = c0; c < c1; ++c) {                      // warp-uniform
                const int type = p.cone_type[c];
                if (type == CONE_BOX) continue;
                const int r0 = p.cone_row[c], d = p.cone_dim[c];
                __syncwarp();
                if (lane == 0) project_serial(r0, d, type);
                __syncwarp();
                for (int r = r0 + lane; r < r0 + d; r += 32) finish(r, zt[r], v[r], p.y[co + r], p.rho[co + r]);
            }
        } else {                       // many cones per stage: one lane per cone
            __syncwarp();
            for (int c = c0 + lane; c < c1; c += 32) {
                const int type = p.cone_type[c];
                if (type == CONE_BOX) continue;
                const int r0 = p.cone_row[c], d = p.cone_dim[c];
                project_serial(r0, d, type);
                for (int r = r0; r < r0 + d; ++r) finish(r, zt[r], v[r], p.y[co + r], p.rho[co + r]);
            }
        }
    }
}


// CHECK iterations: the same update plus the residual norms (r_prim, r_dual and their scales) into the loop state.
__global__ void __launch_bounds__(ADMM_WARPS * 32, 4) admm_update_res_kernel(AdmmParams p) {
    constexpr bool RES = true;
    if (!admm_is_check(p.ctl)) return;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = p.nx + p.nu;
    const size_t ws_len = (size_t)p.N * s + p.nx;
    double* wt = smem + (size_t)warp * (2 * s + 3 * p.ncmax);   // w~ of this stage   (dim)
    double* wd = wt + s;               // w~ - w_prev        (dim)
    double* v = wd + s;                // z^ + y/rho         (ncmax)
    double* zt = v + p.ncmax;          // z~                 (ncmax)
    double* dz = zt + p.ncmax;         // rho o ((1 - alpha)(z~ - z_prev) + (z - z_prev))   (ncmax)
    constexpr bool compute_res = RES;
    double r_prim = 0.0, nrm = 0.0, r_dual = 0.0, nrm_d = 0.0;
    const long long items = (long long)p.batch * (p.N + 1);
#pragma unroll 1
    for (long long item = (long long)blockIdx.x * ADMM_WARPS + warp; item < items; item += (long long)gridDim.x * ADMM_WARPS) {
    const int k = (int)(item % (p.N + 1));
    const int b = (int)(item / (p.N + 1));
    const int dim = (k < p.N) ? s : p.nx;
    const int nc = p.ncs[k];
    const size_t wo = (size_t)b * ws_len + (size_t)k * s;
    __syncwarp();                      // the previous item's readers of the scratch are done
    for (int i = lane; i < dim; i += 32) {
        const double a = p.w_tilde[wo + i], wold = p.w[wo + i];
        wt[i] = a;
        if constexpr (RES) wd[i] = a - wold;
        p.w[wo + i] = p.alpha * a + (1.0 - p.alpha) * wold;
    }
    if (nc == 0) {   // unconstrained stage: only the sigma term of the dual residual
        if (compute_res) {
            __syncwarp();
            for (int j = lane; j < dim; j += 32)
                if (k > 0 || j < p.nu) r_dual = fmax(r_dual, fabs(p.sigma * wd[j]));   // (x_0 is data, not a variable)
        }
        continue;
    }
    __syncwarp();
    const double* Dk = p.Dm + (size_t)b * p.d_total + p.doff[k];
    const size_t co = (size_t)b * p.nc_total + p.coff[k];
    const bool sel = p.sel_col != nullptr;
    for (int r = lane; r < nc; r += 32) {
        double acc = 0.0;
        if (sel) {
            const int cj = p.sel_col[co + r];
            if (cj >= 0) acc = p.sel_val[co + r] * wt[cj];
        } else {
            for (int j = 0; j < dim; ++j) acc = fma(Dk[r + (size_t)j * nc], wt[j], acc);
        }
        zt[r] = acc;
        const double zh = p.alpha * acc + (1.0 - p.alpha) * p.z[co + r];
        v[r] = zh + p.y[co + r] / p.rho[co + r];
    }
    __syncwarp();
    // projections.  Few cones per stage (the usual case: one box over all variables + a cone or two): the warp walks
    // the cones together and clamps a box row-parallel (a lane per cone left 31 lanes idle for 40 serial clamps: 1,980
    // warp-instructions per stage at C4).  Many cones per stage: one lane per cone.
    auto project_serial = [&](int r0, int d, int type) {     // one lane, whole cone
        if (type == CONE_BOX) {
            for (int r = r0; r < r0 + d; ++r) v[r] = fmin(fmax(v[r], p.e_lb[co + r]), p.e_ub[co + r]);
        } else if (type == CONE_SOC) {
            double nv = 0.0;
            for (int r = r0 + 1; r < r0 + d; ++r) nv = fma(v[r], v[r], nv);
            nv = sqrt(nv);
            const double t = v[r0];
            if (nv <= t) { /* inside */ }
            else if (nv <= -t) { for (int r = r0; r < r0 + d; ++r) v[r] = 0.0; }
            else {
                const double a = 0.5 * (t + nv), sc = a / nv;
                v[r0] = a;
                for (int r = r0 + 1; r < r0 + d; ++r) v[r] *= sc;
            }
        } else {  // ball of radius e_ub[first row]
            double nv = 0.0;
            for (int r = r0; r < r0 + d; ++r) nv = fma(v[r], v[r], nv);
            nv = sqrt(nv);
            const double rad = p.e_ub[co + r0];
            if (nv > rad) { const double sc = rad / nv; for (int r = r0; r < r0 + d; ++r) v[r] *= sc; }
        }
    };
    const int c0 = p.cone_first[k], c1 = p.cone_first[k + 1];
    if (c1 - c0 <= 8) {
        for (int c = c0; c < c1; ++c) {                      // warp-uniform
            const int r0 = p.cone_row[c], d = p.cone_dim[c], type = p.cone_type[c];
            if (type == CONE_BOX) {
                for (int r = r0 + lane; r < r0 + d; r += 32) v[r] = fmin(fmax(v[r], p.e_lb[co + r]), p.e_ub[co + r]);
            } else if (lane == 0)
                project_serial(r0, d, type);
        }
    } else {
        for (int c = c0 + lane; c < c1; c += 32) project_serial(p.cone_row[c], p.cone_dim[c], p.cone_type[c]);
    }
    __syncwarp();
    for (int r = lane; r < nc; r += 32) {
        const double zold = p.z[co + r], znew = v[r], rr = p.rho[co + r];
        const double zh = p.alpha * zt[r] + (1.0 - p.alpha) * zold;
        const double ynew = p.y[co + r] + rr * (zh - znew);
        p.z[co + r] = znew;
        p.y[co + r] = ynew;
        if constexpr (RES) {
            dz[r] = rr * ((1.0 - p.alpha) * (zt[r] - zold) + (znew - zold));
            v[r] = ynew;                   // reuse: y for the D^T y norm
            r_prim = fmax(r_prim, fabs(zt[r] - znew));
            nrm = fmax(nrm, fmax(fabs(zt[r]), fabs(znew)));
        }
    }
    if (!compute_res) continue;
    __syncwarp();
    for (int j = lane; j < dim; j += 32) {
        double acc = p.sigma * wd[j], accy = 0.0;
        if (sel) {
            for (int r = 0; r < nc; ++r)
                if (p.sel_col[co + r] == j) {
                    const double dv = p.sel_val[co + r];
                    acc = fma(dv, dz[r], acc);
                    accy = fma(dv, v[r], accy);
                }
        } else {
            for (int r = 0; r < nc; ++r) {
                const double dv = Dk[r + (size_t)j * nc];
                acc = fma(dv, dz[r], acc);
                accy = fma(dv, v[r], accy);
            }
        }
        if (k > 0 || j < p.nu) {   // the rows of x_0 are not stationarity conditions (x_0 is data)
            r_dual = fmax(r_dual, fabs(acc));
            nrm_d = fmax(nrm_d, fabs(accy));
        }
    }
    }   // items
    if (!compute_res) return;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        r_prim = fmax(r_prim, __shfl_xor_sync(0xffffffffu, r_prim, off));
        r_dual = fmax(r_dual, __shfl_xor_sync(0xffffffffu, r_dual, off));
        nrm = fmax(nrm, __shfl_xor_sync(0xffffffffu, nrm, off));
        nrm_d = fmax(nrm_d, __shfl_xor_sync(0xffffffffu, nrm_d, off));
    }
    if (lane == 0) {
        atomic_max_nonneg(&p.ctl->acc[0], r_prim);
        atomic_max_nonneg(&p.ctl->acc[1], r_dual);
        atomic_max_nonneg(&p.ctl->acc[2], nrm);
        atomic_max_nonneg(&p.ctl->acc[3], nrm_d);
    }
}

// Ends an iteration (one thread): counts it, on a check iteration moves the residual maxima out, tests convergence
// (eps_abs + eps_rel * norm, OSQP form), decides on a rho rescale, and sets the WHILE condition of the graph.
__global__ void admm_ctl_kernel(AdmmCtl* c, cudaGraphConditionalHandle handle, int use_handle) {
    const bool check = admm_is_check(c);
    const int it = c->iter + 1;
    int cont = it < c->max_iter ? 1 : 0;
    if (check) {
        double r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r[i] = __longlong_as_double((long long)c->acc[i]);
            c->res[i] = r[i];
            c->acc[i] = 0ull;
        }
        const bool conv = r[0] <= c->eps_abs + c->eps_rel * r[2] && r[1] <= c->eps_abs + c->eps_rel * r[3];
        c->converged = conv ? 1 : 0;
        if (conv) cont = 0;
        if (!conv && cont && c->adaptive && c->n_rho_updates < c->max_rho_updates) {
            const double sc = sqrt((r[0] / fmax(r[2], 1e-30)) / fmax(r[1] / fmax(r[3], 1e-30), 1e-30));
            if (sc > c->rho_tau || sc < 1.0 / c->rho_tau) {
                c->rho_scale = fmin(fmax(sc, 1e-3), 1e3);
                c->rho_update = 1;
                cont = 0;   // the host rescales rho and relaunches: the graph starts with a factorising iteration
            }
        }
    }
    c->iter = it;
    c->cont = cont;
    if (use_handle) cudaGraphSetConditional(handle, (unsigned)cont);
}

__global__ void admm_inv_kernel(const double* __restrict__ rho, double* __restrict__ inv_rho, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        inv_rho[i] = 1.0 / rho[i];
}

__global__ void admm_rho_scale_kernel(double* __restrict__ rho, double* __restrict__ inv_rho, long long n, double scale) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double r = fmin(fmax(rho[i] * scale, 1e-6), 1e6);
        rho[i] = r;
        inv_rho[i] = 1.0 / r;
    }
}

}  // namespace pdplqr
