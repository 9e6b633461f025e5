// launch_floor.cu -- what a chain of DEPENDENT decode-shaped launches costs on this part when the kernels do nothing:
// the floor under the per-launch overhead of the M = 1 stack.  128 launches per graph, 296 CTAs x 288
// threads, 110 KB of dynamic shared memory (two CTAs per SM, like gemv_kernel), programmatic dependent launch.
//   variant 0: empty body                         variant 1: griddepcontrol.wait, then read 4.6 KB (the q8_1 activations of a
//   4096-wide row) that the PREVIOUS launch wrote, write one word per CTA (the dependency a real decode step has)
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a launch_floor.cu -o launch_floor
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <int V>
__global__ void __launch_bounds__(288, 2) floor_kernel(const unsigned* __restrict__ in, unsigned* __restrict__ out) {
    extern __shared__ unsigned sm[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (V == 0) return;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned acc = 0;
    for (int i = threadIdx.x; i < 1152; i += blockDim.x) acc += __ldcg(in + i);
    sm[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 288; i++) acc += sm[i];
        out[blockIdx.x] = acc;
    }
}

template <int V>
static int run(bool pdl) {
    unsigned *a, *b;
    CK(cudaMalloc(&a, 1 << 20)); CK(cudaMalloc(&b, 1 << 20));
    CK(cudaMemset(a, 1, 1 << 20)); CK(cudaMemset(b, 1, 1 << 20));
    CK(cudaFuncSetAttribute(floor_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaGraph_t g; cudaGraphExec_t ge;
    const int n = 128;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    for (int i = 0; i < n; i++) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(296); cfg.blockDim = dim3(288); cfg.dynamicSmemBytes = 110 * 1024; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
        CK(cudaLaunchKernelEx(&cfg, floor_kernel<V>, (const unsigned*)((i & 1) ? b : a), (i & 1) ? a : b));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 6; r++) {
        CK(cudaEventRecord(e0, st));
        for (int k = 0; k < 10; k++) CK(cudaGraphLaunch(ge, st));
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r && ms < best) best = ms;
    }
    printf("variant %d, %s: %.3f us per launch\n", V, pdl ? "programmatic dependent launch" : "plain stream order", best * 1e3 / (10 * n));
    return 0;
}

int main() {
    if (run<0>(true) || run<0>(false) || run<1>(true) || run<1>(false)) return 1;
    return 0;
}
