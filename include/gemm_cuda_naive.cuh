/*
 * include/gemm_cuda_naive.cuh -- drop-in names for the reference's one-thread-per-output
 * kernels (include/gemm_cuda_naive.cuh:285-301): C[M,N] = A_q8_1[M,K] . B_w[N,K]^T.
 * The fp32-activation variants (gemm_fp32_naive, gemm_w4a16_naive, gemm_w8a16_naive) are outside
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
#endif
