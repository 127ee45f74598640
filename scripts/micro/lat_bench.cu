// Microbenchmark: dependent-issue latencies that bound the latency-mode kernels (one warp unless noted).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_bench lat_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double rcp_newton(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    return r;
}

__global__ void k_lat(long long* out, double* sink, int iters) {
    __shared__ double sm[1024];
    __shared__ int ism[64];
    const int t = threadIdx.x;
    for (int i = t; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 1) % 1024);
    if (t < 64) ism[t] = 0;
    __syncthreads();
    long long t0, t1;
    // 1. dependent LDS.64 chain (pointer chasing through shared memory)
    int idx = t & 31;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) idx = (int)sm[idx];
    t1 = clock64();
    if (t == 0) out[0] = (t1 - t0) / iters;
    sink[t] = idx;
    // 2. rcp_newton chain
    double a = 1.5 + t * 1e-3;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) a = rcp_newton(a) + 1.0;
    t1 = clock64();
    if (t == 0) out[1] = (t1 - t0) / iters;
    sink[t] += a;
    // 3. IEEE division chain
    a = 1.5 + t * 1e-3;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) a = 1.0 / a + 1.0;
    t1 = clock64();
    if (t == 0) out[2] = (t1 - t0) / iters;
    sink[t] += a;
    // 4. shared atomicMax followed by a dependent read of the same word
    int v = t;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        atomicMax(&ism[0], v + i);
        __syncwarp();
        v = ism[0] + (t & 1);
    }
    t1 = clock64();
    if (t == 0) out[3] = (t1 - t0) / iters;
    sink[t] += v;
    // 5. bar.sync over the whole CTA, back to back
    __syncthreads();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) __syncthreads();
    t1 = clock64();
    if (t == 0) out[4] = (t1 - t0) / iters;
    // 6. STS -> bar.sync -> LDS (a producer / consumer round trip between warps)
    double x = t;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        sm[t] = x;
        __syncthreads();
        x = sm[(t + 32) % blockDim.x] + 1.0;
        // next store may only overwrite after everybody has read: second buffer instead of a second barrier
        sm[512 + t] = x;
        __syncthreads();
        x = sm[512 + (t + 32) % blockDim.x] + 1.0;
    }
    t1 = clock64();
    if (t == 0) out[5] = (t1 - t0) / (2 * iters);
    sink[t] += x;
    // 7. dependent DFMA, DMUL, FSEL-on-double chains
    a = 1.0 + t * 1e-9;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, 1.0000001, 1e-9);
    t1 = clock64();
    if (t == 0) out[6] = (t1 - t0) / iters;
    sink[t] += a;
    // 8. warp shuffle (64-bit = two SHFL) chain
    a = t;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0;
    t1 = clock64();
    if (t == 0) out[7] = (t1 - t0) / iters;
    sink[t] += a;
}

int main() {
    long long* d_out; double* d_sink;
    cudaMalloc(&d_out, 64 * sizeof(long long));
    cudaMalloc(&d_sink, 1024 * sizeof(double));
    const char* names[] = {"LDS.64 dependent chain", "rcp seed + 2 Newton (+1 DADD)", "IEEE 1/x (+1 DADD)", "ATOMS.MAX + syncwarp + LDS",
                           "bar.sync back to back", "STS -> bar.sync -> LDS (+1 DADD)", "DFMA dependent", "64-bit shfl (+1 DADD)"};
    for (int threads : {32, 128, 256}) {
        k_lat<<<1, threads>>>(d_out, d_sink, 2000);
        k_lat<<<1, threads>>>(d_out, d_sink, 2000);
        cudaDeviceSynchronize();
        long long h[8];
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads per CTA = %d\n", threads);
        for (int i = 0; i < 8; ++i) printf("  %-36s %5lld cycles\n", names[i], h[i]);
    }
    return 0;
}
