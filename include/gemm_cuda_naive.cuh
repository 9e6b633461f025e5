/*
 * include/gemm_cuda_naive.cuh -- drop-in names for the reference's one-thread-per-output
 * kernels (include/gemm_cuda_naive.cuh:267-301): C[M,N] = A[M,K] . B_w[N,K]^T with A = block_q8_1 (W4A8 / W8A8) or
 * fp32 (W4A16 / W8A16, no activation quantization).  gemm_fp32_naive (fp32 x fp32, no quantized operand) is outside
 * this build's path.
 */
#ifndef GEMM_CUDA_NAIVE_CUH
#define GEMM_CUDA_NAIVE_CUH
#include "qgemm_dropin.h"
#include "quant_types.h"

inline void gemm_w4a8_naive(const block_q8_1* A, const block_q4_0* B, float* C, int M, int N, int K,
                            cudaStream_t stream = 0) {
    qgemm_dropin_include(QGEMM_TYPE_Q4_0, A, B, C, M, N, K, stream);
}
inline void gemm_w8a8_naive(const block_q8_1* A, const block_q8_0* B, float* C, int M, int N, int K,
                            cudaStream_t stream = 0) {
    qgemm_dropin_include(QGEMM_TYPE_Q8_0, A, B, C, M, N, K, stream);
}
/* fp32 activations (include/gemm_cuda_naive.cuh:267-283) */
inline void gemm_w4a16_naive(const float* A, const block_q4_0* B, float* C, int M, int N, int K, cudaStream_t stream = 0) {
    qgemm_dropin_status(qgemm_gemm_a16(QGEMM_TYPE_Q4_0, A, B, C, M, N, K, (int64_t)N, 1, 0, (void*)stream), "gemm_w4a16_naive");
}
inline void gemm_w8a16_naive(const float* A, const block_q8_0* B, float* C, int M, int N, int K, cudaStream_t stream = 0) {
    qgemm_dropin_status(qgemm_gemm_a16(QGEMM_TYPE_Q8_0, A, B, C, M, N, K, (int64_t)N, 1, 0, (void*)stream), "gemm_w8a16_naive");
}
#endif
