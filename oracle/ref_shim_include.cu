/*
 * ref_shim_include.cu -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" doorway onto the UNMODIFIED reference sources that live in
 * /root/reference/include (quantize.h, gemm_reference.h, gemm_cuda_naive.cuh,
 * gemm_cuda_dp4a.cuh).  Nothing from the reference is copied into this repo:
 * the headers are #included where they lie and the result is written to
 * oracle/_ref/libqgemm_ref.so (git-ignored).  Used to pin oracle/qgemm_oracle.c,
 * to generate tests/golden/, as the GPU-vs-GPU parity partner, and as the
 * "reference" CPU baseline in bench.py.
 *
 * include/quant_types.h and compat/ggml_types.h cannot share a translation
 * unit (duplicate typedefs), hence two shims: this one and ref_shim_unit.cu.
 */
#include "quantize.h"
#include "gemm_reference.h"
#include "gemm_cuda_naive.cuh"
#include "gemm_cuda_dp4a.cuh"
#include "kernels/activation/silu.cuh"
#include "kernels/normalization/rms_norm.cuh"

extern "C" {

/* ---- CPU: include/quantize.h ---- */
void ref_quantize_row_q8_1_ref(const float* x, void* y, int64_t k) { quantize_row_q8_1_ref(x, (block_q8_1*)y, k); }
void ref_quantize_row_q4_0_ref(const float* x, void* y, int64_t k) { quantize_row_q4_0_ref(x, (block_q4_0*)y, k); }
void ref_quantize_row_q8_0_ref(const float* x, void* y, int64_t k) { quantize_row_q8_0_ref(x, (block_q8_0*)y, k); }
void ref_dequantize_row_q4_0(const void* x, float* y, int64_t k) { dequantize_row_q4_0((const block_q4_0*)x, y, k); }
void ref_dequantize_row_q8_0(const void* x, float* y, int64_t k) { dequantize_row_q8_0((const block_q8_0*)x, y, k); }
void ref_dequantize_row_q8_1(const void* x, float* y, int64_t k) { dequantize_row_q8_1((const block_q8_1*)x, y, k); }

/* ---- CPU: include/gemm_reference.h (include convention: A=q8_1 [M], B=w [N], C[M,N]) ---- */
void ref_gemm_w4a8_reference(const void* A, const void* B, float* C, int M, int N, int K) {
    gemm_w4a8_reference((const block_q8_1*)A, (const block_q4_0*)B, C, M, N, K);
}
void ref_gemm_w8a8_reference(const void* A, const void* B, float* C, int M, int N, int K) {
    gemm_w8a8_reference((const block_q8_1*)A, (const block_q8_0*)B, C, M, N, K);
}
/* fp32-activation references (next row): A fp32 [M,K], B weights [N], C[M,N] */
void ref_gemm_w4a16_reference(const float* A, const void* B, float* C, int M, int N, int K) {
    gemm_w4a16_reference(A, (const block_q4_0*)B, C, M, N, K);
}
void ref_gemm_w8a16_reference(const float* A, const void* B, float* C, int M, int N, int K) {
    gemm_w8a16_reference(A, (const block_q8_0*)B, C, M, N, K);
}
float ref_vec_dot_q4_0_q8_1(int n, const void* vx, const void* vy) { float s; vec_dot_q4_0_q8_1(n, &s, vx, vy); return s; }
float ref_vec_dot_q8_0_q8_1(int n, const void* vx, const void* vy) { float s; vec_dot_q8_0_q8_1(n, &s, vx, vy); return s; }

uint16_t ref_float2half_bits(float f) { __half h = __float2half(f); uint16_t u; memcpy(&u, &h, 2); return u; }
float ref_half_bits2float(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

/* ---- GPU: the reference's own kernels, device pointers, built for sm_100a ---- */
void ref_gpu_quantize_q8_1(const float* x, void* y, int64_t k, void* stream) {
    quantize_q8_1_cuda(x, (block_q8_1*)y, k, (cudaStream_t)stream);
}
void ref_gpu_gemm_w4a8_naive(const void* A, const void* B, float* C, int M, int N, int K, void* stream) {
    gemm_w4a8_naive((const block_q8_1*)A, (const block_q4_0*)B, C, M, N, K, (cudaStream_t)stream);
}
void ref_gpu_gemm_w8a8_naive(const void* A, const void* B, float* C, int M, int N, int K, void* stream) {
    gemm_w8a8_naive((const block_q8_1*)A, (const block_q8_0*)B, C, M, N, K, (cudaStream_t)stream);
}
void ref_gpu_gemm_w4a16_naive(const float* A, const void* B, float* C, int M, int N, int K, void* stream) {
    gemm_w4a16_naive(A, (const block_q4_0*)B, C, M, N, K, (cudaStream_t)stream);
}
void ref_gpu_gemm_w8a16_naive(const float* A, const void* B, float* C, int M, int N, int K, void* stream) {
    gemm_w8a16_naive(A, (const block_q8_0*)B, C, M, N, K, (cudaStream_t)stream);
}
void ref_gpu_silu_mul_f32(const float* x, const float* gate, float* y, int n, void* stream) {
    silu_mul_forward_f32(x, gate, y, n, (cudaStream_t)stream);
}
void ref_cpu_silu_f32(const float* x, float* y, int n) { silu_cpu_f32(x, y, n); }
void ref_cpu_rms_norm_f32(const float* x, const float* w, float* y, int rows, int cols, float eps) { rms_norm_cpu_f32(x, w, y, rows, cols, eps); }
void ref_gpu_rms_norm_f32(const float* x, const float* w, float* y, int rows, int cols, float eps, void* stream) {
    rms_norm_forward_f32(x, w, y, rows, cols, eps, (cudaStream_t)stream);
}
void ref_gpu_gemm_w4a8_tiled_dp4a(const void* A, const void* B, float* C, int M, int N, int K, void* stream) {
    gemm_w4a8_tiled_dp4a((const block_q8_1*)A, (const block_q4_0*)B, C, M, N, K, (cudaStream_t)stream);
}
void ref_gpu_gemm_w8a8_dp4a(const void* A, const void* B, float* C, int M, int N, int K, void* stream) {
    gemm_w8a8_dp4a((const block_q8_1*)A, (const block_q8_0*)B, C, M, N, K, (cudaStream_t)stream);
}

} /* extern "C" */
