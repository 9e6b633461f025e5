/*
 * qgemm.h -- C ABI of libqgemm_sm100.so: the B200 (sm_100a) implementation of
 * the llama.cpp-layout block-quantized GEMM
 *
 *     C[T,F] = A[T,K] . B[F,K]^T      A = block_q8_1 activations (T tokens)
 *                                     B = block_q4_0/q4_1/q5_0/q5_1/q8_0 weights (F rows)
 *
 * plus the quantize_q8_1 activation kernel that feeds it.  Plain pointers and
 * sizes only; every pointer is a DEVICE pointer unless a parameter says "host".
 * The caller owns every buffer, nothing is allocated behind its back, all work
 * is enqueued asynchronously on `stream` (a cudaStream_t passed as void*).
 * Functions return 0 or a negative QGEMM_E_* code; they never exit().
 * There is no CPU fallback: without an sm_100 device every compute entry
 * returns QGEMM_E_ARCH / QGEMM_E_CUDA.
 *
 * Each entry names the reference interface it replaces (paths relative to the
 * root of qhy991/llama.cpp-quant-gemm).  The C++ drop-in headers in this
 * directory (gemm_cuda_naive.cuh, ... , quantize.h, llama_adapter.h) and
 * kernels/gemm/ (gemm_quant_formats.cuh, ...) keep the reference's function
 * names and signatures and forward to these symbols; the python package
 * quant_gemm binds them with ctypes.
 */
#ifndef QGEMM_H
#define QGEMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QGEMM_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define QGEMM_API __attribute__((visibility("default")))
#else
#define QGEMM_API
#endif

/* ---- error codes ---------------------------------------------------------- */
#define QGEMM_OK 0
#define QGEMM_E_BADARG (-1)    /* null pointer, K % 32 != 0, unknown type, negative size   */
#define QGEMM_E_ALIGN (-2)     /* a pointer breaks the format's minimum alignment           */
#define QGEMM_E_ARCH (-3)      /* current device is not sm_100                              */
#define QGEMM_E_CUDA (-4)      /* a CUDA call failed; the CUDA error is left sticky          */
#define QGEMM_E_WORKSPACE (-5) /* workspace missing or smaller than qgemm_workspace_bytes() */

/* ---- tensor types: numerically equal to ggml_type / the reference's
 *      enum QuantType (compat/ggml_types.h:199-215) -------------------------- */
#define QGEMM_TYPE_Q4_0 2
#define QGEMM_TYPE_Q4_1 3
#define QGEMM_TYPE_Q5_0 6
#define QGEMM_TYPE_Q5_1 7
#define QGEMM_TYPE_Q8_0 8
#define QGEMM_TYPE_Q8_1 9

/* ---- quantizer flags ---------------------------- */
#define QGEMM_Q81_ROUND_AWAY 0u  /* roundf(): include/quantize.h:165-193 CPU ref (default)   */
#define QGEMM_Q81_ROUND_EVEN 1u  /* __float2int_rn(): include/quantize.h:302-337 GPU kernel   */
#define QGEMM_Q81_S_FROM_QSUM 2u /* s = half(sum(q) * d): tests/framework/test_framework.cuh:195-225 */
#define QGEMM_Q81_CLAMP127 4u    /* clamp q to [-127,127]: python ext gemm_ops.cu:75-110, framework */
#define QGEMM_Q81_TREE_SUM 8u    /* s = pairwise tree sum (i,i+16),(i,i+8)..: kernels/gemm/gemm_fused.cuh:96-127  */
#define QGEMM_Q81_ID_FROM_HALF_D 16u /* 1/d from the fp16-rounded d, int8 narrowing before the clamp: gemm_fused.cuh:131-140 */
#define QGEMM_Q81_ZERO_D1 32u    /* all-zero block stores d = 1.0: schemas/definitions/quantization/quantize_q8_1.json */
/* quantize_fp16_to_q8_1_smem(), the in-kernel quantizer of gemm_q4_0_fp16_fused (gemm_fused.cuh:76-143), bit for bit: */
#define QGEMM_Q81_FUSED_F16 (QGEMM_Q81_TREE_SUM | QGEMM_Q81_ID_FROM_HALF_D | QGEMM_Q81_CLAMP127)

/* ---- GEMM flags ----------------------------------------------------------- */
#define QGEMM_MS_EXACT 0x1u     /* q4_1/q5_1: + m*s (llama.cpp-true) instead of the
                                   reference's + m*s/4 (gemm_quant_formats.cuh:148,266)     */
#define QGEMM_SEQUENTIAL 0x8u   /* one thread per output, blocks accumulated in order
                                   b = 0..K/32-1 with the reference GPU kernel's exact FMA
                                   sequence: bit-identical to kernels/gemm/
                                   gemm_quant_formats.cuh:312-334 built by nvcc.  Slow.      */
#define QGEMM_WEIGHTS_STATIC 0x10u /* caller promises no work still in flight on `stream` writes the
                                   weight buffer: the decode path then launches with programmatic
                                   dependent launch and prefetches weights under the previous
                                   kernel's tail (activations / C are still ordered normally)       */
#define QGEMM_INPUTS_READY 0x20u  /* with QGEMM_WEIGHTS_STATIC: the caller also promises that this call's
                                   activations and output are not written or read by earlier work still
                                   in flight on `stream` (e.g. the k/v projections after q, `up` after
                                   `gate`: same input, different outputs).  The decode kernel then computes
                                   without waiting for its predecessor and only waits before it exits, so
                                   stream-order completion is preserved for everything launched after it */
#define QGEMM_WEIGHTS_PREPACKED 0x40u /* `weight` is a qgemm_prepack_weights() buffer (tensor-core path only) */
#define QGEMM_STREAM_ALLOC 0x80u /* workspace == NULL and none registered: the library may take the scratch of the
                                   tensor-core path from the stream's memory pool (cudaMallocAsync /
                                   cudaFreeAsync around the launch: stream-ordered, graph-capturable, nothing
                                   persists).  The drop-in C++ headers pass it, because the reference's launcher
                                   signatures have no workspace argument.                                     */
#define QGEMM_FOLD_REFSEQ 0x1000u /* tensor-core path, q4_1 / q5_1 only: fold every block with the reference GPU kernel's
                                   exact operation sequence, d_w*d_a*sumi + m_w*s_a/4 as ((d_w*d_a)*sumi + (m_w*s_a)/4) added to
                                   the sum (gemm_quant_formats.cuh:148,266) -- bit-identical to it, one more FMA-pipe operation
                                   per output pair and block.  Default: acc = fma(m_w, s_a/4, fma(d_w, d_a*sumi, acc)), the same
                                   terms associated differently (~1e-7 of max|C| apart).  The other formats, and every other
                                   path, always use the reference sequence.                                                  */
#define QGEMM_PATH_MASK 0xF00u
#define QGEMM_PATH_AUTO 0x000u
#define QGEMM_PATH_GENERIC 0x100u /* same kernel as QGEMM_SEQUENTIAL                         */
#define QGEMM_PATH_GEMV 0x200u    /* decode: bulk-copy-staged weight stream + dp4a            */
#define QGEMM_PATH_MMA 0x300u     /* skinny: mma.sync m16n8k32 u8/s8, tokens on N             */
#define QGEMM_PATH_TCGEN05 0x400u /* prefill: tcgen05.mma kind::i8, TMEM accumulators         */

/* ---- housekeeping ----------------------------------------------------------- */
QGEMM_API int qgemm_version(void);
QGEMM_API const char *qgemm_strerror(int code);
QGEMM_API size_t qgemm_block_bytes(int type);     /* get_block_bytes(), compat/ggml_types.h:248-258 */
/* Kernels launched by this library since load / since the last reset (all threads). */
QGEMM_API int64_t qgemm_launch_count(void);
QGEMM_API void qgemm_reset_launch_count(void);
/* QGEMM_PATH_* actually taken by the calling thread's most recent qgemm_gemm*(). */
QGEMM_API uint32_t qgemm_last_path(void);
/* Text of the calling thread's most recent QGEMM_E_CUDA failure (CUDA error name and where). */
QGEMM_API const char *qgemm_last_error_detail(void);

/* ---- quantize / dequantize -------------------------------------------------- */
/*
 * x[rows][K] fp32 -> y[rows][K/32] block_q8_1.
 * Replaces quantize_q8_1_cuda() (include/quantize.h:361-368), the python
 * extension's quantize_q8_1_cuda() (python/quant_gemm/csrc/gemm_ops.cu:175-202)
 * and, on the host side of tests, quantize_row_q8_1_ref() (quantize.h:165-193).
 * Default flags reproduce quantize_row_q8_1_ref() byte for byte.
 */
QGEMM_API int qgemm_quantize_q8_1(const float *x, void *y, int64_t rows, int64_t K, uint32_t flags, void *stream);
/*
 * Same from fp16 input x_f16[rows][K] (2-byte aligned): every element enters as __half2float(x), then exactly the
 * arithmetic `flags` selects.  With QGEMM_Q81_FUSED_F16 the bytes are those quantize_fp16_to_q8_1_smem() writes
 * (kernels/gemm/gemm_fused.cuh:76-143), so quantize_q8_1_f16 + qgemm_gemm reproduces gemm_q4_0_fp16_fused (:157-338).
 */
QGEMM_API int qgemm_quantize_q8_1_f16(const void *x_f16, void *y, int64_t rows, int64_t K, uint32_t flags, void *stream);

/*
 * y = quantize_q8_1(silu(x) * gate): the SwiGLU neighbour of the FFN down projection folded into its quantizer
 * (SURVEY 8 f.3).  Replaces silu_mul_forward_f32() (kernels/activation/silu.cuh:97-108, 162-175: y = x / (1 + expf(-x))
 * * gate, same operation sequence, bit-identical intermediate values on the GPU) followed by quantize_q8_1_cuda();
 * x, gate: [rows][K] fp32; flags as for qgemm_quantize_q8_1.  One pass: 9.1 instead of 17.1 bytes per element.
 */
/*
 * y = quantize_q8_1(rms_norm(x) * weight): the normalisation in front of the q/k/v and gate/up projections folded into
 * their quantizer.  Replaces rms_norm_forward_f32() (kernels/normalization/rms_norm.cuh:249-272; arithmetic of
 * rms_norm_cpu_f32, :32-58: sum of squares in double, 1 / sqrtf(mean + eps), then x * inv_rms * weight[i] in that
 * order) followed by quantize_q8_1_cuda().  x: [rows][K] fp32, weight: [K] fp32 (16-byte aligned); row_scratch: `rows`
 * floats of device scratch (1 / rms per row).  Two launches; the normalised fp32 tensor never exists in memory.
 */
QGEMM_API int qgemm_quantize_q8_1_rms_norm(const float *x, const float *weight, void *y, int64_t rows, int64_t K, float eps,
                                           uint32_t flags, float *row_scratch, void *stream);

QGEMM_API int qgemm_quantize_q8_1_silu_mul(const float *x, const float *gate, void *y, int64_t rows, int64_t K, uint32_t flags,
                                           void *stream);

/*
 * Weight quantizers (test-data producers): x[rows][K] fp32 -> y[rows][K/32] blocks.
 * q4_0/q8_0 replace quantize_q4_0_cuda()/quantize_q8_0_cuda()
 * (include/quantize.h:343-359) and python quantize_q4_0 (gemm_ops.cu:146-173);
 * q4_1/q5_0/q5_1 follow testing::quantize::to_q4_1/q5_0/q5_1
 * (tests/framework/test_framework.cuh:256-367).  flags: QGEMM_Q81_ROUND_EVEN
 * selects __float2int_rn rounding (the include/ GPU kernels); default roundf.
 */
QGEMM_API int qgemm_quantize_weight(int wtype, const float *x, void *y, int64_t rows, int64_t K, uint32_t flags,
                          void *stream);

/*
 * blocks -> fp32.  Replaces dequantize_row_q4_0/q8_0/q8_1 (include/quantize.h:84-211)
 * and python dequantize_q4_0 (gemm_ops.cu:204-229).
 */
QGEMM_API int qgemm_dequantize(int type, const void *x, float *y, int64_t rows, int64_t K, void *stream);

/* ---- GEMM ------------------------------------------------------------------- */
/*
 * C[t*ldc_t + f*ldc_f] = sum_{b<K/32} dot(weight[f][b], act_q8_1[t][b])
 *
 *   q4_0: d_w*(d_a*sumi -  8*s_a)      q4_1: d_w*d_a*sumi + m_w*s_a/4   (QGEMM_MS_EXACT: m_w*s_a)
 *   q5_0: d_w*(d_a*sumi - 16*s_a)      q5_1: d_w*d_a*sumi + m_w*s_a/4
 *   q8_0: d_w*d_a*sumi
 *
 * One entry for both of the reference's conventions:
 *   include/ style  gemm_w4a8_*(A_q8_1[M], B_w[N], C[M,N], M,N,K)  (include/gemm_cuda_naive.cuh:285-301,
 *                   gemm_cuda_tiled.cuh:293-301, gemm_cuda_dp4a.cuh:409-444)
 *                   -> T=M, F=N, ldc_t=N, ldc_f=1
 *   ggml style      gemm_q*_q8_1(weight[M], act[N], out[M,N], M,N,K) (kernels/gemm/gemm_quant_formats.cuh:343-428,
 *                   gemm_warp_optimized.cuh:377-1210, gemm_async_copy.cuh:237, gemm_vectorized.cuh:239-280,
 *                   python gemm_q4_0_q8_1 gemm_ops.cu:231-259)
 *                   -> F=M, T=N, ldc_t=1, ldc_f=N
 *
 * workspace: qgemm_workspace_bytes() bytes of device memory, 256-byte aligned
 * (may be NULL when that returns 0).  Alignment: act 4 bytes, weight 2 bytes, C 4 bytes
 * (the reference's requirement); rows that are 16-byte aligned take the fast paths.
 * A call of tensor-core size (T >= 96) that brings no scratch -- no workspace, none registered with
 * qgemm_set_default_workspace(), no QGEMM_STREAM_ALLOC -- returns QGEMM_E_WORKSPACE instead of running on the
 * weight-streaming passes at a fraction of the speed; QGEMM_PATH_MMA requests those passes explicitly.
 * With K % 256 == 0 the scratch holds only the repacked activations (T * K * 1.0625 bytes): the weights are read in
 * their native layout.
 */
QGEMM_API size_t qgemm_workspace_bytes(int wtype, int T, int F, int K, uint32_t flags);

QGEMM_API int qgemm_gemm(int wtype, const void *act_q8_1, const void *weight, float *C, int T, int F, int K,
               int64_t ldc_t, int64_t ldc_f, uint32_t flags, void *workspace, size_t workspace_bytes,
               void *stream);

/*
 * Optional: device scratch the library may use, on the CURRENT device, whenever a qgemm_gemm*()
 * call passes workspace == NULL (the reference's launcher signatures have no workspace
 * argument, so the drop-in C++ headers always do).  The caller still owns the memory and must
 * keep it alive and un-shared between concurrently running streams.  (NULL, 0) unregisters.
 */
QGEMM_API int qgemm_set_default_workspace(void *workspace, size_t workspace_bytes);

/*
 * Offline weight pre-pack for the tensor-core path (static weights): unpacks the native blocks
 * once into the operand-tile layout the prefill kernel streams (`qgemm_prepack_bytes` bytes,
 * 256-byte aligned device memory).  Pass the packed buffer as `weight` together with
 * QGEMM_WEIGHTS_PREPACKED | QGEMM_PATH_TCGEN05 (or AUTO with T >= 96) and the per-call weight
 * prepass disappears; results are unchanged.  The native blocks stay the format of every other path.
 * Data-format neighbour of the path.
 */
QGEMM_API size_t qgemm_prepack_bytes(int wtype, int F, int K);
QGEMM_API int qgemm_prepack_weights(int wtype, const void *weight, int F, int K, void *packed, void *stream);

/*
 * Grouped decode GEMM: `nmat` weight matrices of the same type and K applied to the SAME
 * activations in one launch (fused q/k/v or gate/up projections -- what a concatenated weight
 * matrix would give, without concatenating).  Cs[m][t*ldc_t + f*ldc_f] receives matrix m's rows.
 * One weight stream across the group: no launch bubble between the members.  Results are
 * identical to nmat separate qgemm_gemm() calls on the decode path.  nmat <= 8; T <= 8 takes the
 * fused kernel, anything else falls back to one launch per matrix.
 */
QGEMM_API int qgemm_gemm_group(int wtype, const void *act_q8_1, int nmat, const void *const *weights, float *const *Cs,
                               const int *Fs, int T, int K, int64_t ldc_t, int64_t ldc_f, uint32_t flags, void *stream);

/*
 * Chained decode GEMVs (one token): `nsteps` grouped GEMVs executed in order by ONE persistent launch, with the
 * dependencies between them resolved on the device.  Equivalent to nsteps qgemm_gemm_group() calls (T = 1,
 * ldc_t ignored) issued back to back on `stream`, and to the launches the reference would issue for them
 * (kernels/gemm/gemm_warp_optimized.cuh:377-1210, one launch per projection): same sums, same order, same bits.
 * What changes is the cost of the boundaries: the weight stream of step k+1 is already running while step k drains
 * and while the activations of step k+1 are fetched.
 *
 * A step's activations are either
 *   - act_q8_1: ready-made block_q8_1[K/32], or
 *   - act_f32 (act_q8_1 == NULL): K floats that the kernel quantizes itself with quantize_q8_1's default arithmetic
 *     (include/quantize.h:165-193) -- typically the C of an EARLIER step of the same chain, so the quantize launch
 *     between two projections disappears; with gate_f32 != NULL the quantized value is silu(act_f32[i]) * gate_f32[i]
 *     (kernels/activation/silu.cuh:97-108), the SwiGLU in front of the down projection.  The CTAs quantize the vector
 *     together (one block per warp) into scratch inside `sync` and every CTA copies the result: such a step always waits
 *     for all earlier steps, whatever its flags.
 * Step flags: QGEMM_INPUTS_READY = this step's activations are not produced by an earlier step of the chain (nor by
 * work still in flight when the chain starts computing): it may begin before the earlier steps have finished.  Without
 * it a step starts when every earlier step has completed on the whole device (stream-order semantics).
 *
 * sync: qgemm_gemv_chain_sync_bytes(nsteps) bytes of device memory (arrival counters + 108 KB of quantizer scratch), 4-byte
 * aligned, ZERO before the first use; the
 * kernel leaves it zero, so one buffer serves every chain launched on the same stream.  Chains on different streams need
 * different buffers.  Lists the persistent kernel cannot take (T > 1 shapes, rows that are not 16-byte multiples, K too
 * long for register-resident activations, fewer rows than CTAs, QGEMM_MS_EXACT) are run as one launch per step --
 * same results; qgemm_last_path() tells (QGEMM_PATH_GEMV | QGEMM_PATH_CHAINED when the persistent kernel ran).
 * New entry: the reference has no multi-GEMM launch; it is the successor of its per-projection launch sequence.
 */
#define QGEMM_CHAIN_MAX_MATS 3
typedef struct qgemm_chain_step {
    const void *act_q8_1;                       /* or NULL: quantize act_f32 in the kernel */
    const float *act_f32;
    const float *gate_f32;                      /* optional, with act_f32 */
    int nmat;                                   /* 1..QGEMM_CHAIN_MAX_MATS matrices that share these activations */
    const void *weights[QGEMM_CHAIN_MAX_MATS];
    float *C[QGEMM_CHAIN_MAX_MATS];             /* C[m][f * ldc_f] */
    int F[QGEMM_CHAIN_MAX_MATS];
    int K;
    int64_t ldc_f;
    uint32_t flags;                             /* 0 or QGEMM_INPUTS_READY */
} qgemm_chain_step;
#define QGEMM_PATH_CHAINED 0x1000000u
QGEMM_API size_t qgemm_gemv_chain_sync_bytes(int nsteps);
QGEMM_API int qgemm_gemv_chain_max_steps(void);
QGEMM_API int qgemm_gemv_chain(int wtype, const qgemm_chain_step *steps, int nsteps, uint32_t flags, void *sync,
                               size_t sync_bytes, void *stream);

/*
 * One-shot hint for the calling thread's NEXT decode-path qgemm_gemm*() call: once that launch
 * has issued its own weight stream it also pulls [next_weights, next_weights + bytes) into L2
 * (cp.async.bulk.prefetch.L2), i.e. the weights of the GEMV that will follow it.  Back-to-back
 * decode GEMVs are separated by a dependency bubble in which HBM would otherwise idle; a runtime
 * that knows its layer order (llama.cpp does) hides it this way.  Pure performance hint: no effect
 * on results, ignored by the other paths.  (NULL, 0) clears it.
 */
QGEMM_API int qgemm_hint_next_weights(const void *next_weights, size_t bytes);
/* The same hint as explicit arguments of the call it applies to: no state is kept between calls. */
QGEMM_API int qgemm_gemm_hinted(int wtype, const void *act_q8_1, const void *weight, float *C, int T, int F, int K,
                                int64_t ldc_t, int64_t ldc_f, uint32_t flags, void *workspace, size_t workspace_bytes,
                                void *stream, const void *next_weights, size_t next_bytes);
QGEMM_API int qgemm_gemm_group_hinted(int wtype, const void *act_q8_1, int nmat, const void *const *weights,
                                      float *const *Cs, const int *Fs, int T, int K, int64_t ldc_t, int64_t ldc_f,
                                      uint32_t flags, void *stream, const void *next_weights, size_t next_bytes);

/*
 * Same, with fp32 activations act_f32[T][K]: quantize_q8_1 (flags' QGEMM_Q81_*
 * bits, shifted left by 16) runs first into the workspace, then the GEMM.
 * Successor of gemm_q4_0_fp16_fused() (kernels/gemm/gemm_fused.cuh:311-338) and of
 * the designed-but-unwritten gemm_w4a8() of docs/analysis/W4A8_DATAFLOW_ANALYSIS.md:93-160.
 */
QGEMM_API int qgemm_gemm_f32act(int wtype, const float *act_f32, const void *weight, float *C, int T, int F, int K,
                      int64_t ldc_t, int64_t ldc_f, uint32_t flags, void *workspace,
                      size_t workspace_bytes, void *stream);

/*
 * Same with fp16 activations act_f16[T][K] (2-byte aligned): qgemm_quantize_q8_1_f16 into the workspace, then the GEMM.
 * Successor of gemm_q4_0_fp16_fused() (kernels/gemm/gemm_fused.cuh:311-338), whose kernel quantizes half activations to
 * q8_1 in shared memory before the dot products; with QGEMM_Q81_FUSED_F16 << 16 in `flags` the quantized blocks are the
 * ones its quantize_fp16_to_q8_1_smem() produces.  workspace: T * (K/32) * 36 bytes rounded up to 256, plus
 * qgemm_workspace_bytes() for tensor-core sizes; NULL is accepted together with QGEMM_STREAM_ALLOC.
 */
QGEMM_API int qgemm_gemm_f16act(int wtype, const void *act_f16, const void *weight, float *C, int T, int F, int K,
                                int64_t ldc_t, int64_t ldc_f, uint32_t flags, void *workspace, size_t workspace_bytes,
                                void *stream);

/*
 * The FFN down projection with its SwiGLU neighbour: C = W . quantize_q8_1(silu(x) * gate).  Same contract as
 * qgemm_gemm_f32act (workspace, flags, two launches at tensor-core sizes); x, gate: [T][K] fp32.  Replaces
 * silu_mul_forward_f32 (kernels/activation/silu.cuh:162-175) + quantize_q8_1_cuda + gemm_*: bit-equal to
 * qgemm_quantize_q8_1_silu_mul followed by qgemm_gemm.
 */
QGEMM_API int qgemm_gemm_f32act_silu_mul(int wtype, const float *x, const float *gate, const void *weight, float *C, int T, int F,
                                         int K, int64_t ldc_t, int64_t ldc_f, uint32_t flags, void *workspace,
                                         size_t workspace_bytes, void *stream);

/*
 * W4A16 / W8A16: fp32 activations act_f32[T][K] against Q4_0 / Q8_0 weights with NO activation quantization,
 *     C[t*ldc_t + f*ldc_f] = sum_k act[t][k] * d_w * (q_w - 8)        (q8_0: d_w * q_w)
 * every product and sum in fp32.  Replaces gemm_w4a16_naive / gemm_w8a16_naive (include/gemm_cuda_naive.cuh:66-143,
 * 267-283; CPU: gemm_w4a16_reference / gemm_w8a16_reference, include/gemm_reference.h:73-147) and the python
 * extension's gemm_q4_0_fp32 (python/quant_gemm/csrc/gemm_ops.cu:271-463).  T <= 8: weight-streaming kernel;
 * larger T: register-tiled fp32 GEMM with in-kernel dequantization; both sum K in a parallel order (<= 1e-5 of
 * max|C| from the reference).  QGEMM_SEQUENTIAL: one thread per output in the reference kernel's order and FMA
 * contraction, bit-identical to it; also taken for K % 64 != 0 or act_f32 not 16-byte aligned.  No workspace.
 */
QGEMM_API int qgemm_gemm_a16(int wtype, const float *act_f32, const void *weight, float *C, int T, int F, int K,
                             int64_t ldc_t, int64_t ldc_f, uint32_t flags, void *stream);

/*
 * Test hook for the bit-exactness contract: sumi[(t*F + f)*(K/32) + b] =
 * the int32 block dot product exactly as the selected path computes it
 * (QGEMM_PATH_* in flags picks whose integers are dumped).
 */
QGEMM_API int qgemm_sumi(int wtype, const void *act_q8_1, const void *weight, int32_t *sumi, int T, int F, int K,
               uint32_t flags, void *workspace, size_t workspace_bytes, void *stream);

/* ---- multi-GPU: N-sharded decode with the all-gather fused into the kernel --------
 *
 * New functionality (the reference has no multi-GPU code).  Weight rows
 * are sharded across `world` GPUs of one NVLink/NVSwitch domain; every rank calls
 * qgemm_gemm_peers() with ITS rows and the kernel stores its slice of C directly into every
 * rank's copy of the gathered buffer (peer stores over NVLink) -- no separate all-gather.
 * Completion is tracked by device-side counters in peer-accessible memory:
 *   - all ranks issue the same sequence of launches; a "step" is `launches_per_step` launches
 *     followed by qgemm_peer_step_advance(step) (so the sequence can sit in a CUDA graph);
 *   - a launch waits, before reading its activations, until every rank's launches [0, wait_index)
 *     of the step have landed locally (wait_index = launch_index: everything before it);
 *     qgemm_peer_wait() does the same for the end of the current step.
 * C[r] / flag[r] are peer-mapped device pointers (e.g. torch symmetric memory, cudaIpc, or
 * cuMem fabric handles); flag words and `done`/`step` must start at zero.  Decode kernels for T <= 8,
 * the tensor-core epilogue for larger T.  QGEMM_INPUTS_READY is ignored by the peer entries: their
 * launches share one completion counter, so a launch must not run ahead of its predecessor.
 */
#define QGEMM_MAX_PEERS 8
typedef struct qgemm_peers {
    int world, rank;
    float *C[QGEMM_MAX_PEERS];        /* rank r's destination of this launch's slice (same logical offset everywhere) */
    uint32_t *flag[QGEMM_MAX_PEERS];  /* rank r's arrival counter */
    uint32_t *done;                   /* local scratch counter */
    const uint32_t *step;             /* local step counter */
    uint32_t launches_per_step, launch_index;
    uint32_t wait_index;              /* launches [0, wait_index) of this step (and all earlier steps) must have
                                         landed before this launch reads its activations; launch_index = strict
                                         chain, smaller = the producer of this launch's input finished earlier   */
    float *C_multicast;               /* optional (NULL: unused): an NVLS multicast mapping of the same slice that is
                                         bound on every rank (this one included).  The kernels then store each value
                                         once and the NVSwitch replicates it, instead of one store per rank.        */
} qgemm_peers;

QGEMM_API int qgemm_gemm_peers(int wtype, const void *act_q8_1, const void *weight, const qgemm_peers *peers, int T,
                               int F, int K, int64_t ldc_t, int64_t ldc_f, uint32_t flags, void *stream);
/* Grouped form (see qgemm_gemm_group): matrix m's slice goes to peers->C[r] + c_offsets[m] on every rank r. */
QGEMM_API int qgemm_gemm_group_peers(int wtype, const void *act_q8_1, int nmat, const void *const *weights,
                                     const int *Fs, const int64_t *c_offsets, const qgemm_peers *peers, int T, int K,
                                     int64_t ldc_t, int64_t ldc_f, uint32_t flags, void *stream);
QGEMM_API int qgemm_peer_step_advance(uint32_t *step, void *stream);
QGEMM_API int qgemm_peer_wait(const qgemm_peers *peers, void *stream);

/* ---- multi-GPU sharding helper (host arithmetic only) ------------------------ */
/*
 * Weight rows [0,F) split into `world` contiguous ranges whose sizes are
 * multiples of `align` (except the last): rank's range is [*f0, *f1).
 * New functionality; the reference has no multi-GPU code.
 */
QGEMM_API int qgemm_shard_range(int F, int world, int rank, int align, int *f0, int *f1);

#ifdef __cplusplus
}
#endif
#endif /* QGEMM_H */
