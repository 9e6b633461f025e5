/*
 * compat/ggml_cuda_mul_mat.cuh -- a ggml_cuda_op_mul_mat-shaped hook for llama.cpp's CUDA backend.
 *
 * llama.cpp computes dst = src0 (weights, quantized, [ne00 = K, ne01 = F]) x src1 (activations) by calling, per device
 * and row slice, an op of the shape
 *
 *     void op(ggml_backend_cuda_context& ctx, const ggml_tensor* src0, const ggml_tensor* src1, ggml_tensor* dst,
 *             const char* src0_dd_i, const float* src1_ddf_i, const char* src1_ddq_i, float* dst_dd_i,
 *             int64_t row_low, int64_t row_high, int64_t src1_ncols, int64_t src1_padded_row_size, cudaStream_t stream);
 *
 * (ggml-cuda's ggml_cuda_op_mul_mat_vec_q / _q, the call sites the reference's integration guide patches,
 * docs/guides/INTEGRATION_GUIDE.md:37-44; the reference ships compat/ggml_cuda_compat.cuh for the non-GEMM ops only and
 * no such hook).  This header provides the same contract on plain arguments, so the one-line body of such an op is
 *
 *     qgemm_ggml_cuda_op_mul_mat_q(src0->type, src0_dd_i, src1_ddq_i, dst_dd_i, src0->ne[0], row_low, row_high,
 *                                  src1_ncols, src1_padded_row_size, dst->ne[0], stream);
 *
 *   src0_dd_i   weight rows [row_low, row_high) of this device, native blocks, (row_high - row_low) x K/32
 *   src1_ddq_i  block_q8_1 activations, src1_ncols rows of src1_padded_row_size / 32 blocks (llama.cpp pads K to 512)
 *   dst_dd_i    this slice of the column-major dst: dst_dd_i[col * nrows_dst + (row - row_low)] for the main device
 *               llama.cpp passes nrows_dst = ne0 and a pointer already offset to row_low
 * Returns the qgemm status.  The fp32-activation form (no quantization) is qgemm_ggml_cuda_op_mul_mat_f32act.
 */
#ifndef QGEMM_GGML_CUDA_MUL_MAT_CUH
#define QGEMM_GGML_CUDA_MUL_MAT_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "../include/qgemm.h"

static inline int qgemm_ggml_cuda_op_mul_mat_q(int src0_type, const char* src0_dd_i, const char* src1_ddq_i, float* dst_dd_i, int64_t ne00,
                                               int64_t row_low, int64_t row_high, int64_t src1_ncols, int64_t src1_padded_row_size,
                                               int64_t nrows_dst, cudaStream_t stream) {
    const int K = (int)ne00, F = (int)(row_high - row_low), T = (int)src1_ncols;
    if (src1_padded_row_size == ne00)   /* rows are back to back: one call */
        return qgemm_gemm(src0_type, src1_ddq_i, src0_dd_i, dst_dd_i, T, F, K, nrows_dst, 1, QGEMM_STREAM_ALLOC, 0, 0, (void*)stream);
    /* padded activation rows (K % 512 != 0): the blocks of a row are still contiguous, rows are not -- one call per column */
    const int64_t row_bytes = src1_padded_row_size / 32 * 36;
    for (int64_t t = 0; t < src1_ncols; t++) {
        const int rc = qgemm_gemm(src0_type, src1_ddq_i + t * row_bytes, src0_dd_i, dst_dd_i + t * nrows_dst, 1, F, K, nrows_dst, 1, 0, 0, 0,
                                  (void*)stream);
        if (rc) return rc;
    }
    return 0;
}

static inline int qgemm_ggml_cuda_op_mul_mat_f32act(int src0_type, const char* src0_dd_i, const float* src1_ddf_i, float* dst_dd_i, int64_t ne00,
                                                    int64_t row_low, int64_t row_high, int64_t src1_ncols, int64_t nrows_dst,
                                                    cudaStream_t stream) {
    return qgemm_gemm_a16(src0_type, src1_ddf_i, src0_dd_i, dst_dd_i, (int)src1_ncols, (int)(row_high - row_low), (int)ne00, nrows_dst, 1, 0,
                          (void*)stream);
}

#endif /* QGEMM_GGML_CUDA_MUL_MAT_CUH */
