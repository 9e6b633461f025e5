/*
 * kernels/gemm/gemm_warp_optimized.cuh -- the reference's thirteen decode-shaped Q4_0 launchers
 * (kernels/gemm/gemm_warp_optimized.cuh:377-1210) are successive drafts of one idea; here every
 * name lands on the same entry, which picks the weight-streaming kernel for N <= 8 tokens.
 */
#ifndef KERNELS_GEMM_WARP_OPTIMIZED_CUH
#define KERNELS_GEMM_WARP_OPTIMIZED_CUH
#include "gemm_quant_formats.cuh"

#define QGEMM_Q4_0_ALIAS(name) QGEMM_FORMAT_LAUNCHER(name, block_q4_0, QGEMM_TYPE_Q4_0)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_warp)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_warp_v2)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_warp_prefetch)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_warp_multirow)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_warp_multirow8)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_smem)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_smem_large)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_vec)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_tile2d)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_tile2d_n8)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_tile2d_k256)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_tile2d_r8)
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_tile2d_large)
#endif
