/*
 * qgemm_oracle.h -- CPU oracle for the block-quantized GEMM path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the algorithms
 * of qhy991/llama.cpp-quant-gemm for the hot path (quantize_q8_1 + the five
 * W{4_0,4_1,5_0,5_1,8_0} x A{8_1} block dot products and the GEMM drivers).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load it.  The product library (libqgemm_sm100.so) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * against (a) the reference's own known-answer vectors and
 * (b) the reference's own CPU code compiled from /root/reference into
 * oracle/_ref/libqgemm_ref.so, byte-for-byte, plus the committed fixtures in
 * tests/golden/ generated from that library.
 *
 * Every function cites the reference file:line it restates (paths relative
 * to the reference root).
 */
#ifndef QGEMM_ORACLE_H
#define QGEMM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ggml_type numbering, compat/ggml_types.h:199-215 */
enum {
    QO_Q4_0 = 2,
    QO_Q4_1 = 3,
    QO_Q5_0 = 6,
    QO_Q5_1 = 7,
    QO_Q8_0 = 8,
    QO_Q8_1 = 9
};

/* quantize_q8_1 flavours.  Same bit values as
 * QGEMM_Q81_* in include/qgemm.h. */
#define QO_Q81_ROUND_AWAY 0u      /* roundf, include/quantize.h:165-193 (CPU ref)   */
#define QO_Q81_ROUND_EVEN 1u      /* __float2int_rn, include/quantize.h:302-337     */
#define QO_Q81_S_FROM_QSUM 2u     /* s = half(sum_q * d), test_framework.cuh:195-225 */
#define QO_Q81_CLAMP127 4u        /* clamp to +-127 (py ext gemm_ops.cu:75-110, framework) */
#define QO_Q81_TREE_SUM 8u        /* s = pairwise tree sum (i, i+16), (i, i+8) ...: kernels/gemm/gemm_fused.cuh:96-127 */
#define QO_Q81_ID_FROM_HALF_D 16u /* 1/d taken from the fp16-rounded d: gemm_fused.cuh:131-133                  */
#define QO_Q81_ZERO_D1 32u        /* all-zero block stores d = 1.0: schemas/definitions/quantization/quantize_q8_1.json */
/* the in-kernel quantizer of gemm_q4_0_fp16_fused (gemm_fused.cuh:76-143): fp16 input, the three properties above */
#define QO_Q81_FUSED_F16 (QO_Q81_TREE_SUM | QO_Q81_ID_FROM_HALF_D | QO_Q81_CLAMP127)

/* GEMM flags.  Same bit values as QGEMM_* in include/qgemm.h. */
#define QO_GEMM_MS_EXACT 1u       /* q4_1/q5_1: m*s instead of the reference's m*s/4 */
#define QO_GEMM_Q80_ASSOC_UNIT 2u /* q8_0: (d_w*d_a)*sumi (test_gemm_all_quants.cu:209)
                                     instead of (sumi*d_a)*d_w (gemm_reference.h:261) */
#define QO_GEMM_FMA 4u            /* emulate nvcc's default FMA contraction of the
                                     reference GPU kernels (see qgemm_oracle.c)      */

uint16_t qo_fp32_to_fp16(float f);
float qo_fp16_to_fp32(uint16_t h);

size_t qo_block_bytes(int type);

/* ---- quantizers --------------------------------------------------------- */
/* x[rows*k] fp32 -> y[rows*k/32] block_q8_1 (36 B each). */
/* fp16 input (the reference's fused kernel reads half activations, gemm_fused.cuh:76-143) */
void qo_quantize_q8_1_f16(const uint16_t *x_f16, void *y, int64_t n, unsigned flags);
void qo_silu_mul(const float *x, const float *gate, float *y, int64_t n);
void qo_rms_norm(const float *x, const float *weight, float *y, int64_t n_rows, int64_t n_cols, float eps);
void qo_quantize_q8_1(const float *x, void *y, int64_t n, unsigned flags);

/* include/quantize.h flavours (weights as test data) */
void qo_quantize_q4_0_ref(const float *x, void *y, int64_t n); /* quantize.h:35-70   */
void qo_quantize_q8_0_ref(const float *x, void *y, int64_t n); /* quantize.h:111-135 */
/* tests/framework/test_framework.cuh flavours */
void qo_to_q4_0(const float *x, void *y, int64_t n); /* :162-192 */
void qo_to_q4_1(const float *x, void *y, int64_t n); /* :256-288 */
void qo_to_q5_0(const float *x, void *y, int64_t n); /* :291-329 */
void qo_to_q5_1(const float *x, void *y, int64_t n); /* :332-367 */
void qo_to_q8_0(const float *x, void *y, int64_t n); /* :228-253 */

/* dequantizers: quantize.h:84-102,140-153,198-211 (+ q4_1/q5_x by the format
 * definitions in compat/ggml_types.h) */
void qo_dequantize(int type, const void *x, float *y, int64_t n);

/* ---- block dot products -------------------------------------------------- */
/* Integer part of one block dot (bit-exact contract). */
int32_t qo_block_sumi(int wtype, const void *wblock, const void *ablock);
/* Float value of one block dot, CPU expression order of the reference. */
float qo_block_dot(int wtype, const void *wblock, const void *ablock, unsigned flags);

/* ---- GEMM drivers -------------------------------------------------------- */
/*
 * C[t*ldc_t + f*ldc_f] = sum_b dot(W[f,b], A[t,b]), b sequential from 0.0f.
 * T = tokens (rows of the q8_1 operand), F = weight rows.
 * include/ convention: ldc_t = F, ldc_f = 1; kernels/gemm + python: ldc_t = 1, ldc_f = T.
 * Rows [t0,t1) only (so callers can slab it over threads).
 */
void qo_gemm(int wtype, const void *act_q8_1, const void *weight, float *C,
             int T, int F, int K, int64_t ldc_t, int64_t ldc_f, unsigned flags,
             int t0, int t1);

/* sumi[(t*F + f)*nb + b] for t in [t0,t1). */
void qo_gemm_sumi(int wtype, const void *act_q8_1, const void *weight, int32_t *sumi,
                  int T, int F, int K, int t0, int t1);

/*
 * SURVEY.md section 8(f).3 (next row, oracle half only -- no product kernel yet): fp32 activations against
 * dequantized Q4_0 / Q8_0 weights, include/gemm_reference.h:73-147 (CPU) and
 * include/gemm_cuda_naive.cuh:66-143 (GPU; QO_GEMM_FMA replays its mul+add contraction).
 * C[t*ldc_t + f*ldc_f] = sum over blocks, then over k in the reference's order
 * (Q4_0: element k, then k+16, for k = 0..15; Q8_0: k = 0..31), w = (q - 8) * d resp. q * d.
 */
void qo_gemm_f32act_dequant(int wtype, const float *act, const void *weight, float *C,
                            int T, int F, int K, int64_t ldc_t, int64_t ldc_f, unsigned flags);

#ifdef __cplusplus
}
#endif
#endif
