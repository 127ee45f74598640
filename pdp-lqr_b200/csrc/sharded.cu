// Single-process horizon sharding over several GPUs behind the C ABI (include/pdplqr.h, pdplqr_sharded_*): one very long
// LQ problem is cut into contiguous time slices, one per device; every device reduces its slice to ONE summary
// (P | F | C | p | f), the summaries are all-gathered with NCCL over NVLink (ncclCommInitAll: one communicator per device
// in this process), every device solves the small interface system of the G slices redundantly and rolls out its slice.
//
// The reference's analogue is threads <-> segments with the serial condensed solve on the master thread between the two
// sweeps (/root/reference include/clqr/lqr/lqr_solver_parallel.hpp:142-146, 213-238); composition of summaries is
// associative (tree_kernels.cuh), which is what makes the hierarchy segments -> device -> box legitimate.  The multi-process
// version of the same flow (one process per GPU, torch.distributed) is pdp-lqr_b200/sharding.py; this file is the entry a
// C++17 host program uses (include/pdplqr/lqr_cuda_sharded_solver.hpp).
//
// NCCL is loaded at run time (dlopen "libnccl.so.2": no link-time dependency; inside a process that already has a NCCL
// loaded, e.g. torch's, that copy is used).  There is no fallback: without NCCL or without enough devices create fails.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/pdplqr.h"

namespace {

// the few NCCL entry points used, declared here so that building the library does not need nccl.h
typedef void* ncclComm_t;
typedef int ncclResult_t;
constexpr int NCCL_FLOAT64 = 8;   // ncclDataType_t: ncclFloat64 / ncclDouble
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror(); return false; }
        auto sym = [&](const char* n) { return dlsym(lib, n); };
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(sym("ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!CommInitAll || !CommDestroy || !AllGather || !GroupStart || !GroupEnd) { err = "NCCL symbols missing"; return false; }
        return true;
    }
};
Nccl g_nccl;

struct Shard {
    int device = 0, start = 0, count = 0;
    bool last = false;
    pdplqr_handle_t h = nullptr, coupler = nullptr;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    long long c0 = 0, c1 = 0;        // this slice's range inside the full constraint vectors
    double *d_ws = nullptr, *d_out = nullptr, *d_x0 = nullptr, *d_sum = nullptr, *d_all = nullptr, *d_xhat = nullptr,
           *d_lam = nullptr, *d_ys = nullptr, *d_zs = nullptr, *d_rho = nullptr, *d_inv = nullptr;
    std::vector<int> ncs;
};

}  // namespace

struct pdplqr_sharded {
    int nx = 0, nu = 0, s = 0, N = 0, G = 0, srec = 0;
    long long nc_total = 0;
    std::vector<int> ncs;
    std::vector<long long> coff, doff;
    std::vector<Shard> shards;
    std::string err;
};

namespace {

int sfail(pdplqr_sharded* hs, int code, const std::string& msg) {
    if (hs) hs->err = msg;
    return code;
}
#define SH_CUDA(hs, expr)                                                                              \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess) return sfail(hs, PDPLQR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define SH_CALL(hs, sh, expr)                                                                          \
    do {                                                                                               \
        int rc_ = (expr);                                                                              \
        if (rc_ != PDPLQR_OK) return sfail(hs, rc_, std::string(#expr) + ": " + pdplqr_last_error((sh).h)); \
    } while (0)

}  // namespace

extern "C" {

int pdplqr_sharded_create(pdplqr_sharded_t* out, int nx, int nu, int N, const int* ncs, int num_devices, const int* devices,
                          int segments_per_device, int condensed_type) {
    if (!out) return PDPLQR_ERR_INVALID;
    *out = nullptr;
    if (nx < 1 || nu < 1 || num_devices < 1 || N < num_devices || segments_per_device < 0) return PDPLQR_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < num_devices) return PDPLQR_ERR_CUDA;
    pdplqr_sharded* hs = new pdplqr_sharded();
    hs->nx = nx; hs->nu = nu; hs->s = nx + nu; hs->N = N; hs->G = num_devices;
    hs->ncs.assign(N + 1, 0);
    if (ncs) hs->ncs.assign(ncs, ncs + N + 1);
    hs->coff.assign(N + 2, 0); hs->doff.assign(N + 2, 0);
    for (int k = 0; k <= N; ++k) {
        hs->coff[k + 1] = hs->coff[k] + hs->ncs[k];
        hs->doff[k + 1] = hs->doff[k] + (long long)hs->ncs[k] * (k < N ? hs->s : nx);
    }
    hs->nc_total = hs->coff[N + 1];
    auto bail = [&](int rc, const std::string& why) {
        static thread_local std::string keep;
        keep = why;
        pdplqr_sharded_destroy(hs);
        return rc;
    };
    std::string why;
    if (num_devices > 1 && !g_nccl.load(why)) return bail(PDPLQR_ERR_CUDA, why);
    hs->shards.resize(num_devices);
    std::vector<int> devs(num_devices);
    const int base = N / num_devices, rem = N % num_devices;
    int start = 0;
    for (int d = 0; d < num_devices; ++d) {
        Shard& sh = hs->shards[d];
        sh.device = devices ? devices[d] : d;
        devs[d] = sh.device;
        sh.start = start; sh.count = base + (d < rem ? 1 : 0); sh.last = (d == num_devices - 1);
        start += sh.count;
        // constraint rows travel with their stages; an interior slice has no terminal rows
        sh.ncs.assign(sh.count + 1, 0);
        for (int k = 0; k < sh.count; ++k) sh.ncs[k] = hs->ncs[sh.start + k];
        if (sh.last) sh.ncs[sh.count] = hs->ncs[N];
        sh.c0 = hs->coff[sh.start];
        sh.c1 = sh.last ? hs->coff[N + 1] : hs->coff[sh.start + sh.count];
        if (cudaSetDevice(sh.device) != cudaSuccess) return bail(PDPLQR_ERR_CUDA, "cudaSetDevice");
        int rc = pdplqr_create(&sh.h, nx, nu, sh.count, hs->nc_total > 0 ? sh.ncs.data() : nullptr, 1, segments_per_device, 2,
                               condensed_type, sh.device);
        if (rc != PDPLQR_OK) return bail(rc, std::string("pdplqr_create (shard): ") + pdplqr_last_error(nullptr));
        if (!sh.last && (rc = pdplqr_set_option(sh.h, PDPLQR_OPT_INTERIOR_SHARD, 1)) != PDPLQR_OK) return bail(rc, "interior shard option");
        rc = pdplqr_coupler_create(&sh.coupler, nx, nu, num_devices, 1, sh.device);
        if (rc != PDPLQR_OK) return bail(rc, "pdplqr_coupler_create");
        if (cudaStreamCreateWithFlags(&sh.stream, cudaStreamNonBlocking) != cudaSuccess) return bail(PDPLQR_ERR_CUDA, "stream");
        pdplqr_set_stream(sh.h, sh.stream);
        pdplqr_set_stream(sh.coupler, sh.stream);
        hs->srec = pdplqr_summary_doubles(sh.h);
        const size_t wsl = (size_t)sh.count * hs->s + nx, nc = (size_t)std::max<long long>(sh.c1 - sh.c0, 1);
        bool ok = cudaMalloc(&sh.d_ws, wsl * 8) == cudaSuccess && cudaMalloc(&sh.d_out, wsl * 8) == cudaSuccess &&
                  cudaMalloc(&sh.d_x0, nx * 8) == cudaSuccess && cudaMalloc(&sh.d_sum, hs->srec * 8) == cudaSuccess &&
                  cudaMalloc(&sh.d_all, (size_t)num_devices * hs->srec * 8) == cudaSuccess &&
                  cudaMalloc(&sh.d_xhat, (size_t)num_devices * nx * 8) == cudaSuccess &&
                  cudaMalloc(&sh.d_lam, (size_t)num_devices * nx * 8) == cudaSuccess;
        if (ok && hs->nc_total > 0)
            ok = cudaMalloc(&sh.d_ys, nc * 8) == cudaSuccess && cudaMalloc(&sh.d_zs, nc * 8) == cudaSuccess &&
                 cudaMalloc(&sh.d_rho, nc * 8) == cudaSuccess && cudaMalloc(&sh.d_inv, nc * 8) == cudaSuccess;
        if (!ok) return bail(PDPLQR_ERR_CUDA, "cudaMalloc (shard buffers)");
    }
    if (num_devices > 1) {
        std::vector<ncclComm_t> comms(num_devices);
        ncclResult_t r = g_nccl.CommInitAll(comms.data(), num_devices, devs.data());
        if (r != 0) return bail(PDPLQR_ERR_CUDA, std::string("ncclCommInitAll: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
        for (int d = 0; d < num_devices; ++d) hs->shards[d].comm = comms[d];
    }
    *out = hs;
    return PDPLQR_OK;
}

int pdplqr_sharded_destroy(pdplqr_sharded_t hs) {
    if (!hs) return PDPLQR_OK;
    for (Shard& sh : hs->shards) {
        cudaSetDevice(sh.device);
        if (sh.stream) cudaStreamSynchronize(sh.stream);
        if (sh.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(sh.comm);
        if (sh.h) pdplqr_destroy(sh.h);
        if (sh.coupler) pdplqr_destroy(sh.coupler);
        for (double* p : {sh.d_ws, sh.d_out, sh.d_x0, sh.d_sum, sh.d_all, sh.d_xhat, sh.d_lam, sh.d_ys, sh.d_zs, sh.d_rho, sh.d_inv})
            if (p) cudaFree(p);
        if (sh.stream) cudaStreamDestroy(sh.stream);
    }
    delete hs;
    return PDPLQR_OK;
}

// E, c, H, h [N][...], HN, hN, D: host arrays of the FULL horizon in the flat layout of pdplqr.h (batch = 1)
int pdplqr_sharded_set_model(pdplqr_sharded_t hs, const double* E, const double* c, const double* H, const double* hvec,
                             const double* HN, const double* hN, const double* D) {
    if (!hs || !E || !c || !H || !hvec || !HN || !hN) return sfail(hs, PDPLQR_ERR_INVALID, "sharded_set_model: null pointer");
    if (hs->nc_total > 0 && !D) return sfail(hs, PDPLQR_ERR_INVALID, "sharded_set_model: D is required");
    const int nx = hs->nx, s = hs->s;
    std::vector<double> zHN((size_t)nx * nx, 0.0), zhN(nx, 0.0);
    for (Shard& sh : hs->shards) {
        SH_CUDA(hs, cudaSetDevice(sh.device));
        const size_t k0 = sh.start;
        const double* Dk = D ? D + hs->doff[sh.start] : nullptr;
        SH_CALL(hs, sh, pdplqr_set_model(sh.h, E + k0 * nx * s, c + k0 * nx, H + k0 * s * s, hvec + k0 * s,
                                         sh.last ? HN : zHN.data(), sh.last ? hN : zhN.data(), Dk));
    }
    return PDPLQR_OK;
}

// update_problem_data + backward + forward over all devices; host arrays of the full horizon.  ws_in / ys / zs / rho /
// inv_rho may be NULL as in pdplqr_solve.
int pdplqr_sharded_solve(pdplqr_sharded_t hs, const double* ws_in, const double* ys, const double* zs, const double* rho,
                         const double* inv_rho, double sigma, const double* x0, double* ws_out) {
    if (!hs || !x0 || !ws_out) return sfail(hs, PDPLQR_ERR_INVALID, "sharded_solve: null pointer");
    const bool con = hs->nc_total > 0;
    if (con && (!ys || !zs || !rho || !inv_rho)) return sfail(hs, PDPLQR_ERR_INVALID, "sharded_solve: ys, zs, rho, inv_rho are required");
    const int nx = hs->nx, s = hs->s, G = hs->G;
    // local sweeps + slice summaries
    for (Shard& sh : hs->shards) {
        SH_CUDA(hs, cudaSetDevice(sh.device));
        const size_t wsl = (size_t)sh.count * s + nx;
        if (ws_in) SH_CUDA(hs, cudaMemcpyAsync(sh.d_ws, ws_in + (size_t)sh.start * s, wsl * 8, cudaMemcpyHostToDevice, sh.stream));
        SH_CUDA(hs, cudaMemcpyAsync(sh.d_x0, x0, nx * 8, cudaMemcpyHostToDevice, sh.stream));
        const size_t nc = (size_t)(sh.c1 - sh.c0);
        if (con && nc > 0) {
            SH_CUDA(hs, cudaMemcpyAsync(sh.d_ys, ys + sh.c0, nc * 8, cudaMemcpyHostToDevice, sh.stream));
            SH_CUDA(hs, cudaMemcpyAsync(sh.d_zs, zs + sh.c0, nc * 8, cudaMemcpyHostToDevice, sh.stream));
            SH_CUDA(hs, cudaMemcpyAsync(sh.d_rho, rho + sh.c0, nc * 8, cudaMemcpyHostToDevice, sh.stream));
            SH_CUDA(hs, cudaMemcpyAsync(sh.d_inv, inv_rho + sh.c0, nc * 8, cudaMemcpyHostToDevice, sh.stream));
        }
        SH_CALL(hs, sh, pdplqr_update_problem_data_device(sh.h, ws_in ? sh.d_ws : nullptr, con ? sh.d_ys : nullptr,
                                                          con ? sh.d_zs : nullptr, con ? sh.d_inv : nullptr, sigma));
        SH_CALL(hs, sh, pdplqr_backward_device(sh.h, con ? sh.d_rho : nullptr));
        SH_CALL(hs, sh, pdplqr_get_root_summary_device(sh.h, G > 1 ? sh.d_sum : sh.d_all));
    }
    // the only exchange: one all-gather of a (3 nx^2 + 2 nx)-double summary per device
    if (G > 1) {
        g_nccl.GroupStart();
        for (Shard& sh : hs->shards) {
            cudaSetDevice(sh.device);
            ncclResult_t r = g_nccl.AllGather(sh.d_sum, sh.d_all, (size_t)hs->srec, NCCL_FLOAT64, sh.comm, sh.stream);
            if (r != 0) { g_nccl.GroupEnd(); return sfail(hs, PDPLQR_ERR_CUDA, "ncclAllGather failed"); }
        }
        if (g_nccl.GroupEnd() != 0) return sfail(hs, PDPLQR_ERR_CUDA, "ncclGroupEnd failed");
    }
    // redundant interface solve, boundary values, rollout, copy back
    for (int d = 0; d < G; ++d) {
        Shard& sh = hs->shards[d];
        SH_CUDA(hs, cudaSetDevice(sh.device));
        int rc = pdplqr_coupler_solve_device(sh.coupler, sh.d_all, sh.d_x0, sh.d_xhat, sh.d_lam);
        if (rc != PDPLQR_OK) return sfail(hs, rc, std::string("coupler: ") + pdplqr_last_error(sh.coupler));
        SH_CALL(hs, sh, pdplqr_set_root_boundary_device(sh.h, sh.d_xhat + (size_t)d * nx, sh.d_lam + (size_t)d * nx));
        SH_CALL(hs, sh, pdplqr_forward_device(sh.h, sh.d_x0, sh.d_out));
        const size_t n = (size_t)sh.count * s + (sh.last ? nx : 0);
        SH_CUDA(hs, cudaMemcpyAsync(ws_out + (size_t)sh.start * s, sh.d_out, n * 8, cudaMemcpyDeviceToHost, sh.stream));
    }
    for (Shard& sh : hs->shards) {
        SH_CUDA(hs, cudaSetDevice(sh.device));
        SH_CUDA(hs, cudaStreamSynchronize(sh.stream));
    }
    return PDPLQR_OK;
}

int pdplqr_sharded_num_devices(pdplqr_sharded_t hs) { return hs ? hs->G : PDPLQR_ERR_INVALID; }
const char* pdplqr_sharded_last_error(pdplqr_sharded_t hs) { return hs ? hs->err.c_str() : "null handle"; }

}  // extern "C"
