/* kernels/gemm/gemm_async_copy.cuh -- drop-in name for gemm_q4_0_q8_1_async (reference :237-262). */
#ifndef KERNELS_GEMM_ASYNC_COPY_CUH
#define KERNELS_GEMM_ASYNC_COPY_CUH
#include "gemm_warp_optimized.cuh"
QGEMM_Q4_0_ALIAS(gemm_q4_0_q8_1_async)
#endif
