/*
 * include/llama_adapter.h -- ggml_tensor front-ends.
 *
 * The reference declares gemm_w4a8_from_ggml / gemm_w4a16_from_ggml / validate_tensor_types /
 * extract_dims_from_tensor (include/llama_adapter.h:52-120) but defines them nowhere; these are the definitions, with
 * the declared signatures: void return, `const char* kernel_type = "naive"` accepted for source compatibility (the
 * library picks its own path), work enqueued on stream 0 like the reference's launchers.  A failure (bad types, wrong
 * device ...) is reported on stderr and kept in qgemm_dropin_last_status(), see qgemm_dropin.h.
 *
 * Include ggml.h BEFORE this header: the definitions need the complete struct ggml_tensor (ne[], type, data); without
 * it (GGML_MAX_DIMS undefined) only the declarations are visible.  Activation tensor: Q8_1 (W4A8) or F32 (W4A16)
 * [K, M]; weight tensor: [K, N] of a supported type; output: F32 [N, M] in ggml order = row-major C[M, N], which is
 * the include/ convention of the C ABI.  Tensor data must be device pointers.
 *
 * The *_on_stream templates take any tensor-like type and a stream, and return the qgemm status.
 */
#ifndef LLAMA_ADAPTER_H
#define LLAMA_ADAPTER_H

#include "gemm_cuda_dp4a.cuh"
#include "gemm_cuda_naive.cuh"
#include "gemm_cuda_tiled.cuh"
#include "quant_types.h"

struct ggml_tensor;

#define QGEMM_GGML_TYPE_F32 0

template <typename Tensor>
inline void qgemm_tensor_dims(const Tensor* activation, const Tensor* weights, int* M, int* N, int* K) {
    *M = (int)activation->ne[1];
    *K = (int)activation->ne[0];
    *N = (int)weights->ne[1];
}
template <typename Tensor>
inline bool qgemm_tensor_types_ok(const Tensor* activation, const Tensor* weights, const Tensor* output, int expected_activation_type,
                                  int expected_weight_type, int expected_output_type) {
    return activation && weights && output && (int)activation->type == expected_activation_type &&
           (int)weights->type == expected_weight_type && (int)output->type == expected_output_type &&
           activation->ne[0] == weights->ne[0] && output->ne[0] == weights->ne[1] && output->ne[1] == activation->ne[1];
}
template <typename T, typename Tensor> inline T* get_tensor_data(Tensor* tensor) { return reinterpret_cast<T*>(tensor->data); }
template <typename T, typename Tensor> inline const T* get_tensor_data(const Tensor* tensor) { return reinterpret_cast<const T*>(tensor->data); }

/* any weight type of the path against Q8_1 activations; returns the qgemm status */
template <typename Tensor>
inline int gemm_w4a8_on_stream(const Tensor* activation, const Tensor* weights, Tensor* output, cudaStream_t stream) {
    if (!qgemm_tensor_types_ok(activation, weights, output, QGEMM_TYPE_Q8_1, weights ? (int)weights->type : -1, QGEMM_GGML_TYPE_F32))
        return QGEMM_E_BADARG;
    int M, N, K;
    qgemm_tensor_dims(activation, weights, &M, &N, &K);
    return qgemm_gemm((int)weights->type, activation->data, weights->data, (float*)output->data, M, N, K, N, 1, QGEMM_STREAM_ALLOC,
                      nullptr, 0, (void*)stream);
}
/* Q4_0 / Q8_0 weights against F32 activations, no activation quantization */
template <typename Tensor>
inline int gemm_w4a16_on_stream(const Tensor* activation, const Tensor* weights, Tensor* output, cudaStream_t stream) {
    if (!qgemm_tensor_types_ok(activation, weights, output, QGEMM_GGML_TYPE_F32, weights ? (int)weights->type : -1, QGEMM_GGML_TYPE_F32))
        return QGEMM_E_BADARG;
    int M, N, K;
    qgemm_tensor_dims(activation, weights, &M, &N, &K);
    return qgemm_gemm_a16((int)weights->type, (const float*)activation->data, weights->data, (float*)output->data, M, N, K, N, 1, 0,
                          (void*)stream);
}

#ifdef GGML_MAX_DIMS   /* ggml.h is in: struct ggml_tensor is complete */
inline void extract_dims_from_tensor(const struct ggml_tensor* activation, const struct ggml_tensor* weights, int* M, int* N, int* K) {
    qgemm_tensor_dims(activation, weights, M, N, K);
}
inline bool validate_tensor_types(const struct ggml_tensor* activation, const struct ggml_tensor* weights, const struct ggml_tensor* output,
                                  int expected_activation_type, int expected_weight_type, int expected_output_type) {
    return qgemm_tensor_types_ok(activation, weights, output, expected_activation_type, expected_weight_type, expected_output_type);
}
inline void gemm_w4a8_from_ggml(const struct ggml_tensor* activation, const struct ggml_tensor* weights, struct ggml_tensor* output,
                                const char* kernel_type = "naive") {
    (void)kernel_type;
    qgemm_dropin_status(gemm_w4a8_on_stream(activation, weights, output, 0), "gemm_w4a8_from_ggml");
}
inline void gemm_w4a16_from_ggml(const struct ggml_tensor* activation, const struct ggml_tensor* weights, struct ggml_tensor* output,
                                 const char* kernel_type = "naive") {
    (void)kernel_type;
    qgemm_dropin_status(gemm_w4a16_on_stream(activation, weights, output, 0), "gemm_w4a16_from_ggml");
}
#else
void extract_dims_from_tensor(const struct ggml_tensor* activation, const struct ggml_tensor* weights, int* M, int* N, int* K);
bool validate_tensor_types(const struct ggml_tensor* activation, const struct ggml_tensor* weights, const struct ggml_tensor* output,
                           int expected_activation_type, int expected_weight_type, int expected_output_type);
void gemm_w4a8_from_ggml(const struct ggml_tensor* activation, const struct ggml_tensor* weights, struct ggml_tensor* output,
                         const char* kernel_type = "naive");
void gemm_w4a16_from_ggml(const struct ggml_tensor* activation, const struct ggml_tensor* weights, struct ggml_tensor* output,
                          const char* kernel_type = "naive");
#endif
/* gemm_fp32_from_ggml (fp32 x fp32, include/llama_adapter.h:99-104) has no quantized operand and is outside this path */

#endif /* LLAMA_ADAPTER_H */
