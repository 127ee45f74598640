// Microbenchmark: shared-memory throughput of warp-wide 64-bit / 128-bit loads as a function of the address pattern
// (how many wavefronts does an LDS cost?).  16 warps on one SM issue independent loads; cycles per warp-load x 1 SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bench lds_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int VEC>
__global__ void k_lds(long long* out, double* sink, int iters, int mod, int stride) {
    extern __shared__ __align__(16) double sm[];
    const int t = threadIdx.x, lane = t & 31;
    for (int i = t; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    // lane -> element index: (lane % mod) * stride  (mod distinct addresses per warp)
    const int base = (lane % mod) * stride * VEC;
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm + base);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const unsigned o = sbase + (unsigned)((i & 7) * 64 * VEC * 8);
        if (VEC == 1) {
            double a, b, c, d;   // asm volatile: the compiler may neither hoist nor merge the loads
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(o));
            asm volatile("ld.shared.f64 %0, [%1+4096];" : "=d"(b) : "r"(o));
            asm volatile("ld.shared.f64 %0, [%1+8192];" : "=d"(c) : "r"(o));
            asm volatile("ld.shared.f64 %0, [%1+12288];" : "=d"(d) : "r"(o));
            acc0 += a; acc1 += b; acc2 += c; acc3 += d;
        } else {
            double a0, a1, b0, b1, c0, c1, d0, d1;
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a0), "=d"(a1) : "r"(o));
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+8192];" : "=d"(b0), "=d"(b1) : "r"(o));
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+16384];" : "=d"(c0), "=d"(c1) : "r"(o));
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+20480];" : "=d"(d0), "=d"(d1) : "r"(o));
            acc0 += a0 + a1; acc1 += b0 + b1; acc2 += c0 + c1; acc3 += d0 + d1;
        }
    }
    long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + t] = acc0 + acc1 + acc2 + acc3;
    if (t == 0) out[0] = t1 - t0;
}

int main() {
    long long* d_out; double* d_sink;
    cudaMalloc(&d_out, 8 * sizeof(long long));
    cudaMalloc(&d_sink, 4096 * sizeof(double));
    const int iters = 4096, warps = 16;
    struct P { int mod, stride; const char* name; };
    const P pats[] = {{32, 1, "32 distinct, contiguous"}, {16, 1, "16 distinct, contiguous (lane % 16)"},
                      {12, 1, "12 distinct, contiguous (lane % 12)"}, {8, 1, "8 distinct, contiguous (lane % 8)"},
                      {4, 1, "4 distinct (lane % 4)"}, {2, 1, "2 distinct"}, {1, 1, "1 address (broadcast)"},
                      {8, 3, "8 distinct, stride 3"}, {16, 3, "16 distinct, stride 3"}, {12, 2, "12 distinct, stride 2"},
                      {32, 3, "32 distinct, stride 3"}};
    for (int vec = 1; vec <= 2; ++vec) {
        printf("%s, %d warps on one SM: SM cycles per warp-load (= wavefronts if the data pipe is the limit)\n",
               vec == 1 ? "LDS.64" : "LDS.128", warps);
        for (const P& p : pats) {
            for (int rep = 0; rep < 2; ++rep) {
                if (vec == 1) k_lds<1><<<1, warps * 32, 32768>>>(d_out, d_sink, iters, p.mod, p.stride);
                else k_lds<2><<<1, warps * 32, 32768>>>(d_out, d_sink, iters, p.mod, p.stride);
            }
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            long long h;
            cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
            printf("  %-40s %6.2f\n", p.name, (double)h / (iters * 4.0 * warps));
        }
    }
    return 0;
}
