/* kernels/gemm/gemm_fused.cuh -- drop-in name for gemm_q4_0_fp16_fused (reference :311-338): Q4_0 weights [M rows] against
 * FP16 activations [N][K], output[m * N + n]; the activations are quantized to q8_1 with the arithmetic of the reference's
 * in-kernel quantizer quantize_fp16_to_q8_1_smem (:76-143: tree sum, 1/d from the half-rounded d, +-127 clamp) and then take
 * the q8_1 GEMM.  (The reference kernel runs that quantizer from all warps of a CTA over one set of static shared arrays,
 * :96-127 + :214-222, so its own output depends on warp timing; the single-warp function is what is reproduced.) */
#ifndef KERNELS_GEMM_FUSED_CUH
#define KERNELS_GEMM_FUSED_CUH
#include <cuda_fp16.h>

#include "../../include/qgemm_dropin.h"
#include "../../include/quant_types.h"

inline void gemm_q4_0_fp16_fused(const block_q4_0* weight, const half* fp16_activation, float* output, int M, int N, int K,
                                 cudaStream_t stream = 0) {
    qgemm_dropin_status(qgemm_gemm_f16act(QGEMM_TYPE_Q4_0, fp16_activation, weight, output, N, M, K, 1, (int64_t)N,
                                          QGEMM_STREAM_ALLOC | ((uint32_t)QGEMM_Q81_FUSED_F16 << 16), nullptr, 0, (void*)stream),
                        "gemm_q4_0_fp16_fused");
}
#endif
