#include <cuda_runtime.h>
#include <cstdio>
__global__ void ctl(cudaGraphConditionalHandle h, int* it, int maxit) {
    int v = ++(*it);
    cudaGraphSetConditional(h, v < maxit ? 1 : 0);
}
__global__ void work(int* x) { atomicAdd(x, 1); }
int main() {
    cudaStream_t s; cudaStreamCreate(&s);
    int* d; cudaMalloc(&d, 8); cudaMemset(d, 0, 8);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h;
    cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
    work<<<1, 1, 0, s>>>(d + 1);
    ctl<<<1, 1, 0, s>>>(h, d, 5);
    cudaStreamCaptureStatus st; unsigned long long id; cudaGraph_t cg; const cudaGraphNode_t* deps; size_t nd;
    cudaStreamGetCaptureInfo_v2(s, &st, &id, &cg, &deps, &nd);
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.conditional.handle = h; cp.conditional.type = cudaGraphCondTypeWhile; cp.conditional.size = 1;
    cudaGraphNode_t node;
    cudaGraphAddNode(&node, g, deps, nd, &cp);
    cudaStreamUpdateCaptureDependencies(s, &node, 1, cudaStreamSetCaptureDependencies);
    cudaStreamEndCapture(s, &cg);
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
    work<<<1, 1, 0, s>>>(d + 1);
    ctl<<<1, 1, 0, s>>>(h, d, 5);
    cudaStreamEndCapture(s, nullptr);
    cudaGraphExec_t ex; cudaError_t e = cudaGraphInstantiate(&ex, g, 0);
    printf("inst %s\n", cudaGetErrorString(e));
    cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
    int hres[2]; cudaMemcpy(hres, d, 8, cudaMemcpyDeviceToHost);
    printf("it %d work %d err %s\n", hres[0], hres[1], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
