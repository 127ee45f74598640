// Microbenchmark: FP64 pipe (DFMA) and FP64 tensor core (mma.sync.m8n8k4.f64, "DMMA") rate and latency on one SM and
// on the whole chip.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dmma(double* out, int iters, long long* cyc) {
    double c[CHAINS][2];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c[i][0] = i; c[i][1] = -i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS>
__global__ void k_dfma(double* out, int iters, long long* cyc) {
    double c[CHAINS];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <class K>
void run(const char* name, K kern, int blocks, int threads, int chains, double flop_per_inst_warp) {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, sizeof(double) * blocks * threads); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    kern<<<blocks, threads>>>(out, 16, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<blocks, threads>>>(out, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double insts = (double)iters * chains;                       // per warp
    const double warps = (double)blocks * threads / 32;
    printf("%-28s blocks=%4d threads=%4d chains=%2d : %.2f cyc/inst/warp, chip %.2f TFLOP/s\n", name, blocks, threads, chains,
           (double)h / insts, insts * warps * flop_per_inst_warp / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    // latency: one warp, one dependent chain
    run("DMMA latency (1 warp,1 chain)", k_dmma<1>, 1, 32, 1, 512);
    run("DFMA latency (1 warp,1 chain)", k_dfma<1>, 1, 32, 1, 64);
    run("DMMA 1 warp 8 chains", k_dmma<8>, 1, 32, 8, 512);
    run("DFMA 1 warp 8 chains", k_dfma<8>, 1, 32, 8, 64);
    run("DMMA 1 SM 4 warps 8 chains", k_dmma<8>, 1, 128, 8, 512);
    run("DMMA 1 SM 16 warps 8 chains", k_dmma<8>, 1, 512, 8, 512);
    run("DFMA 1 SM 16 warps 8 chains", k_dfma<8>, 1, 512, 8, 64);
    run("DMMA chip (148x4 CTAs x 256)", k_dmma<8>, 148 * 4, 256, 8, 512);
    run("DFMA chip (148x4 CTAs x 256)", k_dfma<8>, 148 * 4, 256, 8, 64);
    return 0;
}
