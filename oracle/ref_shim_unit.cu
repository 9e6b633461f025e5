/*
 * ref_shim_unit.cu -- TEST INFRASTRUCTURE ONLY (see ref_shim_include.cu).
 *
 * Doorway onto the reference's all-formats unit test
 * (tests/unit/test_gemm_all_quants.cu: cpu_gemm_q{4_0,4_1,5_0,5_1,8_0}_q8_1, the
 * only CPU statement of q4_1/q5_0/q5_1), its test-framework quantizers
 * (tests/framework/test_framework.cuh: testing::quantize::to_q*), and the
 * five-format GPU kernels (kernels/gemm/gemm_quant_formats.cuh).  The test file
 * is included unmodified; only its main() is renamed out of the way.
 */
#include <thread>
#include <vector>
#include <algorithm>

#define main ref_all_quants_test_main
#include "tests/unit/test_gemm_all_quants.cu"
#undef main
#include "kernels/gemm/gemm_warp_optimized.cuh"

namespace {
typedef void (*cpu_gemm_fn)(const void*, const void*, float*, int, int, int);
template <typename BW, void (*FN)(const BW*, const block_q8_1*, float*, int, int, int)>
void call(const void* w, const void* a, float* o, int M, int N, int K) { FN((const BW*)w, (const block_q8_1*)a, o, M, N, K); }
cpu_gemm_fn pick(int wtype) {
    switch (wtype) {
    case QUANT_TYPE_Q4_0: return call<block_q4_0, cpu_gemm_q4_0_q8_1>;
    case QUANT_TYPE_Q4_1: return call<block_q4_1, cpu_gemm_q4_1_q8_1>;
    case QUANT_TYPE_Q5_0: return call<block_q5_0, cpu_gemm_q5_0_q8_1>;
    case QUANT_TYPE_Q5_1: return call<block_q5_1, cpu_gemm_q5_1_q8_1>;
    case QUANT_TYPE_Q8_0: return call<block_q8_0, cpu_gemm_q8_0_q8_1>;
    default: return nullptr;
    }
}
}

extern "C" {

/* ggml convention: weight [M rows], activation [N tokens], output[m*N+n]. */
int ref_cpu_gemm(int wtype, const void* weight, const void* act, float* out, int M, int N, int K) {
    cpu_gemm_fn fn = pick(wtype);
    if (!fn) return -1;
    fn(weight, act, out, M, N, K);
    return 0;
}

/* Same function, called unmodified on slabs of weight rows from nthreads host threads. */
int ref_cpu_gemm_threaded(int wtype, const void* weight, const void* act, float* out,
                          int M, int N, int K, int nthreads) {
    cpu_gemm_fn fn = pick(wtype);
    if (!fn) return -1;
    if (nthreads < 1) nthreads = 1;
    const size_t rowbytes = (size_t)(K / 32) * get_block_bytes((QuantType)wtype);
    std::vector<std::thread> th;
    const int slab = (M + nthreads - 1) / nthreads;
    for (int i = 0; i < nthreads; i++) {
        const int m0 = i * slab, m1 = std::min(M, m0 + slab);
        if (m0 >= m1) break;
        th.emplace_back([=] {
            fn((const char*)weight + (size_t)m0 * rowbytes, act, out + (size_t)m0 * N, m1 - m0, N, K);
        });
    }
    for (auto& t : th) t.join();
    return 0;
}

/* tests/framework/test_framework.cuh quantizers */
void ref_to_q4_0(const float* s, void* d, int n) { testing::quantize::to_q4_0(s, (block_q4_0*)d, n); }
void ref_to_q4_1(const float* s, void* d, int n) { testing::quantize::to_q4_1(s, (block_q4_1*)d, n); }
void ref_to_q5_0(const float* s, void* d, int n) { testing::quantize::to_q5_0(s, (block_q5_0*)d, n); }
void ref_to_q5_1(const float* s, void* d, int n) { testing::quantize::to_q5_1(s, (block_q5_1*)d, n); }
void ref_to_q8_0(const float* s, void* d, int n) { testing::quantize::to_q8_0(s, (block_q8_0*)d, n); }
void ref_to_q8_1(const float* s, void* d, int n) { testing::quantize::to_q8_1(s, (block_q8_1*)d, n); }

/* GPU: kernels/gemm/gemm_quant_formats.cuh launchers (ggml convention), device pointers */
int ref_gpu_gemm_quant(int wtype, const void* w, const void* a, float* o, int M, int N, int K, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    switch (wtype) {
    case QUANT_TYPE_Q4_0: gemm_q4_0_q8_1((const block_q4_0*)w, (const block_q8_1*)a, o, M, N, K, s); break;
    case QUANT_TYPE_Q4_1: gemm_q4_1_q8_1((const block_q4_1*)w, (const block_q8_1*)a, o, M, N, K, s); break;
    case QUANT_TYPE_Q5_0: gemm_q5_0_q8_1((const block_q5_0*)w, (const block_q8_1*)a, o, M, N, K, s); break;
    case QUANT_TYPE_Q5_1: gemm_q5_1_q8_1((const block_q5_1*)w, (const block_q8_1*)a, o, M, N, K, s); break;
    case QUANT_TYPE_Q8_0: gemm_q8_0_q8_1((const block_q8_0*)w, (const block_q8_1*)a, o, M, N, K, s); break;
    default: return -1;
    }
    return (int)cudaGetLastError();
}
/* the reference's best decode kernels, for the "beat the reference GPU path" timing */
int ref_gpu_gemm_q4_0_tile2d(const void* w, const void* a, float* o, int M, int N, int K, void* stream) {
    gemm_q4_0_q8_1_tile2d((const block_q4_0*)w, (const block_q8_1*)a, o, M, N, K, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}
int ref_gpu_gemm_q4_0_warp_multirow(const void* w, const void* a, float* o, int M, int N, int K, void* stream) {
    gemm_q4_0_q8_1_warp_multirow((const block_q4_0*)w, (const block_q8_1*)a, o, M, N, K, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}

} /* extern "C" */
