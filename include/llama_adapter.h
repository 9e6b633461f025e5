/*
 * include/llama_adapter.h -- ggml_tensor front-ends.
 *
 * The reference declares gemm_w4a8_from_ggml / validate_tensor_types (include/llama_adapter.h:
 * 71-120) but defines them nowhere; these are the definitions.  Include ggml.h (struct
 * ggml_tensor with ne[], type, data) before this header.  Activation tensor: Q8_1 [K, M];
 * weight tensor: Q4_0 (or any supported weight type) [K, N]; output: F32 [N, M] in ggml
 * order = row-major C[M, N], which is exactly the include/ convention of the C ABI.
 * `kernel_type` ("naive" | "tiled" | "dp4a") is accepted for source compatibility; the library
 * picks its own path.
 */
#ifndef LLAMA_ADAPTER_H
#define LLAMA_ADAPTER_H

#include "gemm_cuda_dp4a.cuh"
#include "gemm_cuda_naive.cuh"
#include "gemm_cuda_tiled.cuh"
#include "quant_types.h"

/* a translation unit that has no ggml.h can still see the declarations */
struct ggml_tensor;

template <typename Tensor>
inline void extract_dims_from_tensor(const Tensor* activation, const Tensor* weights, int* M, int* N, int* K) {
    *M = (int)activation->ne[1];
    *K = (int)activation->ne[0];
    *N = (int)weights->ne[1];
}

template <typename Tensor>
inline bool validate_tensor_types(const Tensor* activation, const Tensor* weights, const Tensor* output,
                                  int expected_activation_type, int expected_weight_type, int expected_output_type) {
    return activation && weights && output && (int)activation->type == expected_activation_type &&
           (int)weights->type == expected_weight_type && (int)output->type == expected_output_type &&
           activation->ne[0] == weights->ne[0] && output->ne[0] == weights->ne[1] && output->ne[1] == activation->ne[1];
}

template <typename T, typename Tensor> inline T* get_tensor_data(Tensor* tensor) { return reinterpret_cast<T*>(tensor->data); }
template <typename T, typename Tensor> inline const T* get_tensor_data(const Tensor* tensor) { return reinterpret_cast<const T*>(tensor->data); }

/* returns the qgemm status (0 = ok); tensors' data must be device pointers */
template <typename Tensor>
inline int gemm_w4a8_from_ggml(const Tensor* activation, const Tensor* weights, Tensor* output,
                               const char* kernel_type = "naive", cudaStream_t stream = 0) {
    (void)kernel_type;
    if (!validate_tensor_types(activation, weights, output, QGEMM_TYPE_Q8_1, (int)weights->type, /*F32*/ 0)) return QGEMM_E_BADARG;
    int M, N, K;
    extract_dims_from_tensor(activation, weights, &M, &N, &K);
    return qgemm_gemm((int)weights->type, activation->data, weights->data, (float*)output->data, M, N, K, N, 1, QGEMM_STREAM_ALLOC,
                      nullptr, 0, (void*)stream);
}

#endif /* LLAMA_ADAPTER_H */
