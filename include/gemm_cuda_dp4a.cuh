/*
 * include/gemm_cuda_dp4a.cuh -- drop-in names for the reference's dp4a launchers (:409-444).
 * gemm_w4a8_dp4a in the reference reads nibbles in an interleaved order that no other part of
 * that repo (or llama.cpp) writes; here it uses the llama.cpp layout
 * like every other entry, i.e. it agrees with gemm_w4a8_reference.
 */
#ifndef GEMM_CUDA_DP4A_CUH
#define GEMM_CUDA_DP4A_CUH
#include "qgemm_dropin.h"
#include "quant_types.h"

#define QGEMM_W4A8_ALIAS(name)                                                                                  \
    inline void name(const block_q8_1* A, const block_q4_0* B, float* C, int M, int N, int K, cudaStream_t stream = 0) { \
        qgemm_dropin_include(QGEMM_TYPE_Q4_0, A, B, C, M, N, K, stream);                                         \
    }
QGEMM_W4A8_ALIAS(gemm_w4a8_dp4a)
QGEMM_W4A8_ALIAS(gemm_w4a8_tiled_dp4a)
QGEMM_W4A8_ALIAS(gemm_w4a8_vectorized_dp4a)
#undef QGEMM_W4A8_ALIAS

inline void gemm_w8a8_dp4a(const block_q8_1* A, const block_q8_0* B, float* C, int M, int N, int K,
                           cudaStream_t stream = 0) {
    qgemm_dropin_include(QGEMM_TYPE_Q8_0, A, B, C, M, N, K, stream);
}
#endif
