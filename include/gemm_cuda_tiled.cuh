/* include/gemm_cuda_tiled.cuh -- drop-in name for gemm_w4a8_tiled (reference :293-301). */
#ifndef GEMM_CUDA_TILED_CUH
#define GEMM_CUDA_TILED_CUH
#include "qgemm_dropin.h"
#include "quant_types.h"

inline void gemm_w4a8_tiled(const block_q8_1* A, const block_q4_0* B, float* C, int M, int N, int K,
                            cudaStream_t stream = 0) {
    qgemm_dropin_include(QGEMM_TYPE_Q4_0, A, B, C, M, N, K, stream);
}
#endif
