/*
 * kernels/gemm/gemm_quant_formats.cuh -- drop-in names for the five-format launchers of the
 * reference (kernels/gemm/gemm_quant_formats.cuh:343-428): output[M,N] = weight[M,K] . act[N,K]^T,
 * M = weight rows, N = tokens.
 */
#ifndef KERNELS_GEMM_QUANT_FORMATS_CUH
#define KERNELS_GEMM_QUANT_FORMATS_CUH
#include "../../compat/ggml_types.h"
#include "../../include/qgemm_dropin.h"

#define QGEMM_FORMAT_LAUNCHER(name, block_t, qtype)                                                              \
    inline void name(const block_t* weight, const block_q8_1* activation, float* output, int M, int N, int K,    \
                     cudaStream_t stream = 0) {                                                                   \
        qgemm_dropin_ggml(qtype, weight, activation, output, M, N, K, stream);                                   \
    }
QGEMM_FORMAT_LAUNCHER(gemm_q4_0_q8_1, block_q4_0, QGEMM_TYPE_Q4_0)
QGEMM_FORMAT_LAUNCHER(gemm_q4_1_q8_1, block_q4_1, QGEMM_TYPE_Q4_1)
QGEMM_FORMAT_LAUNCHER(gemm_q5_0_q8_1, block_q5_0, QGEMM_TYPE_Q5_0)
QGEMM_FORMAT_LAUNCHER(gemm_q5_1_q8_1, block_q5_1, QGEMM_TYPE_Q5_1)
QGEMM_FORMAT_LAUNCHER(gemm_q8_0_q8_1, block_q8_0, QGEMM_TYPE_Q8_0)
#endif
