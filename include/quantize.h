/*
 * include/quantize.h -- GPU quantizers of the include/ API.
 *
 * Drop-in for the device half of the reference header (include/quantize.h:343-368):
 * quantize_q4_0_cuda / quantize_q8_0_cuda / quantize_q8_1_cuda(x, y, k, stream) on device
 * pointers, k = total element count.  Rounding follows the reference's GPU kernels
 * (__float2int_rn, :250,:292,:335) so bytes match what that header produces on a GPU; pass
 * QGEMM_Q81_ROUND_AWAY to qgemm_quantize_q8_1() directly for the CPU reference's roundf().
 * The host-side *_ref functions of the reference header are test oracles and live in oracle/.
 */
#ifndef QUANTIZE_H
#define QUANTIZE_H

#include "qgemm_dropin.h"
#include "quant_types.h"

inline void quantize_q4_0_cuda(const float* x, block_q4_0* y, int64_t k, cudaStream_t stream = 0) {
    qgemm_dropin_status(qgemm_quantize_weight(QGEMM_TYPE_Q4_0, x, y, 1, k, QGEMM_Q81_ROUND_EVEN, (void*)stream), "quantize_q4_0_cuda");
}
inline void quantize_q8_0_cuda(const float* x, block_q8_0* y, int64_t k, cudaStream_t stream = 0) {
    qgemm_dropin_status(qgemm_quantize_weight(QGEMM_TYPE_Q8_0, x, y, 1, k, QGEMM_Q81_ROUND_EVEN, (void*)stream), "quantize_q8_0_cuda");
}
inline void quantize_q8_1_cuda(const float* x, block_q8_1* y, int64_t k, cudaStream_t stream = 0) {
    qgemm_dropin_status(qgemm_quantize_q8_1(x, y, 1, k, QGEMM_Q81_ROUND_EVEN, (void*)stream), "quantize_q8_1_cuda");
}

#endif /* QUANTIZE_H */
