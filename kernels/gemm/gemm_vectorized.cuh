/* kernels/gemm/gemm_vectorized.cuh -- drop-in names for gemm_q4_0_q8_1_vec_safe / _vec_float4
 * (reference :239-280; the reference's _vec_float4 does misaligned 8-byte loads). */
#ifndef KERNELS_GEMM_VECTORIZED_CUH
#define KERNELS_GEMM_VECTORIZED_CUH
#include "gemm_warp_optimized.cuh"
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_vec_safe)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_vec_float4)
#endif
