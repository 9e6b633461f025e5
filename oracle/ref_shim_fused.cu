/*
 * ref_shim_fused.cu -- TEST INFRASTRUCTURE ONLY (see ref_shim_include.cu).
 *
 * Doorway onto the reference's fused fp16-activation kernel file (kernels/gemm/gemm_fused.cuh), included unmodified
 * from where it lies: its in-kernel quantizer quantize_fp16_to_q8_1_smem (:76-143) -- the "A2'" flavour of
 * quantize_q8_1 with fp16 input, a pairwise tree sum and 1/d taken from the half-rounded d -- and the fused GEMM
 * gemm_q4_0_fp16_fused (:157-338) that uses it.
 */
#include "kernels/gemm/gemm_fused.cuh"

namespace {
__global__ void ref_fused_quantize_kernel(const half* x, block_q8_1* y, int nblocks) {
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        quantize_fp16_to_q8_1_smem(x + (size_t)b * 32, y + b, threadIdx.x);
        __syncthreads();
    }
}
}

extern "C" {
/* device pointers; one 32-thread CTA walks blocks b, b + grid, ... through the reference's device function */
int ref_gpu_quantize_fp16_to_q8_1_smem(const void* x_f16, void* y, int nblocks, void* stream) {
    if (nblocks <= 0) return 0;
    ref_fused_quantize_kernel<<<nblocks < 2048 ? nblocks : 2048, 32, 0, (cudaStream_t)stream>>>((const half*)x_f16, (block_q8_1*)y, nblocks);
    return (int)cudaGetLastError();
}
} /* extern "C" */
