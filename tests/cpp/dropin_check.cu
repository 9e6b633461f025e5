// dropin_check.cu -- a translation unit written against the REFERENCE's C++ surface only
// (compat/ggml_types.h, include/*.cuh|h, kernels/gemm/*.cuh names and signatures), built against
// this repository's drop-in headers and linked with libqgemm_sm100.so.
//
//   dropin_check <dir>    reads  <dir>/x.f32 [T*K], <dir>/w_q4_0.bin, <dir>/w_q5_1.bin, <dir>/w_q8_0.bin, <dir>/dims.txt
//                         writes <dir>/a_q8_1.bin, c_inc_q4_0.f32 [T,F], c_inc_q8_0.f32 [T,F],
//                                c_ggml_q4_0.f32 [F,T], c_ggml_q5_1.f32 [F,T], c_tile2d.f32 [F,T], c_big.f32 [F,T],
//                                c_w4a16.f32 / c_w8a16.f32 / c_adapter16.f32 / c_hook16.f32 [T,F] (fp32 activations), c_hook.f32 [T,F]
#include <cstdio>
#include <string>
#include <vector>

#include "compat/ggml_types.h"
#include "include/quant_types.h"
#include "include/quantize.h"
#include "include/gemm_cuda_naive.cuh"
#include "include/gemm_cuda_tiled.cuh"
#include "include/gemm_cuda_dp4a.cuh"
#include "kernels/gemm/gemm_quant_formats.cuh"
#include "kernels/gemm/gemm_warp_optimized.cuh"
#include "kernels/gemm/gemm_async_copy.cuh"
#include "kernels/gemm/gemm_vectorized.cuh"
#include "kernels/gemm/gemm_fused.cuh"

// a stand-in for ggml.h (only what the adapter touches): the complete tensor type + the macro that says it is there
#define GGML_MAX_DIMS 4
struct ggml_tensor { int type; int64_t ne[GGML_MAX_DIMS]; void* data; };
#include "include/llama_adapter.h"
#include "compat/ggml_cuda_mul_mat.cuh"

static std::vector<char> slurp(const std::string& p) {
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", p.c_str()); exit(2); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<char> b(n);
    if (fread(b.data(), 1, n, f) != (size_t)n) exit(2);
    fclose(f);
    return b;
}
static void dump(const std::string& p, const void* d, size_t n) { FILE* f = fopen(p.c_str(), "wb"); fwrite(d, 1, n, f); fclose(f); }
template <typename T> static T* to_dev(const std::vector<char>& h) {
    T* d; CUDA_CHECK(cudaMalloc(&d, h.size())); CUDA_CHECK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice)); return d;
}

int main(int argc, char** argv) {
    static_assert(sizeof(block_q4_0) == 18 && sizeof(block_q4_1) == 20 && sizeof(block_q5_0) == 22 &&
                  sizeof(block_q5_1) == 24 && sizeof(block_q8_0) == 34 && sizeof(block_q8_1) == 36, "sizes");
    static_assert(QUANT_TYPE_Q4_0 == 2 && QUANT_TYPE_Q8_1 == 9 && quant_traits<block_q8_1>::has_sum, "enum / traits");
    if (get_block_bytes(QUANT_TYPE_Q5_0) != 22 || get_block_size(QUANT_TYPE_Q4_K) != 256) return 3;
    if (argc < 2) { printf("compiled against the drop-in headers: ok\n"); return 0; }
    const std::string dir = argv[1];
    int T, F, K;
    { FILE* f = fopen((dir + "/dims.txt").c_str(), "r"); if (!f || fscanf(f, "%d %d %d", &T, &F, &K) != 3) return 2; fclose(f); }
    const int nb = K / 32;
    float* x = to_dev<float>(slurp(dir + "/x.f32"));
    block_q4_0* w4 = to_dev<block_q4_0>(slurp(dir + "/w_q4_0.bin"));
    block_q5_1* w51 = to_dev<block_q5_1>(slurp(dir + "/w_q5_1.bin"));
    block_q8_0* w8 = to_dev<block_q8_0>(slurp(dir + "/w_q8_0.bin"));
    block_q8_1* a; CUDA_CHECK(cudaMalloc(&a, (size_t)T * nb * sizeof(block_q8_1)));
    float* c; CUDA_CHECK(cudaMalloc(&c, (size_t)T * F * sizeof(float)));
    std::vector<float> h((size_t)T * F);
    auto fetch = [&](const char* name) {
        KERNEL_CHECK();
        CUDA_CHECK(cudaMemcpy(h.data(), c, h.size() * 4, cudaMemcpyDeviceToHost));
        dump(dir + "/" + name, h.data(), h.size() * 4);
    };
    cudaStream_t st; CUDA_CHECK(cudaStreamCreate(&st));

    quantize_q8_1_cuda(x, a, (int64_t)T * K, st);                 // include/quantize.h
    CUDA_CHECK(cudaStreamSynchronize(st));
    { std::vector<char> ha((size_t)T * nb * 36); CUDA_CHECK(cudaMemcpy(ha.data(), a, ha.size(), cudaMemcpyDeviceToHost)); dump(dir + "/a_q8_1.bin", ha.data(), ha.size()); }

    gemm_w4a8_naive(a, w4, c, T, F, K, st);  fetch("c_inc_q4_0.f32");            // include/ convention, C[T,F]
    gemm_w4a8_tiled_dp4a(a, w4, c, T, F, K); fetch("c_inc_q4_0_b.f32");          // default stream
    gemm_w8a8_dp4a(a, w8, c, T, F, K, st);   fetch("c_inc_q8_0.f32");
    gemm_q4_0_q8_1(w4, a, c, F, T, K, st);   fetch("c_ggml_q4_0.f32");           // kernels/gemm convention, out[F,T]
    gemm_q5_1_q8_1(w51, a, c, F, T, K, st);  fetch("c_ggml_q5_1.f32");
    gemm_q4_0_q8_1_tile2d(w4, a, c, F, T, K, st); fetch("c_tile2d.f32");
    gemm_q4_0_q8_1_async(w4, a, c, F, T, K, st);  fetch("c_async.f32");

    // ggml_tensor adapter: activation [K, M], weights [K, N], output [N, M]
    ggml_tensor ta{QUANT_TYPE_Q8_1, {K, T, 1, 1}, a}, tw{QUANT_TYPE_Q4_0, {K, F, 1, 1}, w4}, to{QUANT_TYPE_F32, {F, T, 1, 1}, c};
    gemm_w4a8_from_ggml(&ta, &tw, &to, "dp4a");                                   // the reference's declared signature
    if (qgemm_dropin_last_status() != 0) return 4;
    fetch("c_adapter.f32");
    int dm, dn, dk;
    extract_dims_from_tensor(&ta, &tw, &dm, &dn, &dk);
    if (dm != T || dn != F || dk != K || !validate_tensor_types(&ta, &tw, &to, QUANT_TYPE_Q8_1, QUANT_TYPE_Q4_0, QUANT_TYPE_F32)) return 6;
    ggml_tensor bad{QUANT_TYPE_Q4_0, {K, T, 1, 1}, a};                           // wrong activation type: reported, not silently dropped
    gemm_w4a8_from_ggml(&bad, &tw, &to);
    if (qgemm_dropin_last_status() != QGEMM_E_BADARG) return 7;

    // fp32 activations, no quantization: W4A16 / W8A16 launchers, adapter and the mul_mat-shaped hooks
    gemm_w4a16_naive(x, w4, c, T, F, K, st);  fetch("c_w4a16.f32");
    gemm_w8a16_naive(x, w8, c, T, F, K, st);  fetch("c_w8a16.f32");
    ggml_tensor tx{QUANT_TYPE_F32, {K, T, 1, 1}, x};
    gemm_w4a16_from_ggml(&tx, &tw, &to);
    if (qgemm_dropin_last_status() != 0) return 8;
    fetch("c_adapter16.f32");
    if (qgemm_ggml_cuda_op_mul_mat_f32act(QUANT_TYPE_Q4_0, (const char*)w4, x, c, K, 0, F, T, F, st) != 0) return 9;
    fetch("c_hook16.f32");
    if (qgemm_ggml_cuda_op_mul_mat_q(QUANT_TYPE_Q4_0, (const char*)w4, (const char*)a, c, K, 0, F, T, K, F, st) != 0) return 10;
    fetch("c_hook.f32");

    // fp16 activations through the reference's fused launcher name: quantized like its in-kernel quantizer, then the q8_1 GEMM
    {
        std::vector<char> hx = slurp(dir + "/x.f16");
        half* xh = to_dev<half>(hx);
        gemm_q4_0_fp16_fused(w4, xh, c, F, T, K, st);
        if (qgemm_dropin_last_status() != 0) return 11;
        fetch("c_fused16.f32");
    }

    // the unchanged launcher reaches the tensor-core path on its own (scratch from the stream's pool) ...
    printf("launcher path 0x%x\n", qgemm_last_path());
    // ... or with scratch the caller registered
    size_t wsb = qgemm_workspace_bytes(QGEMM_TYPE_Q4_0, T, F, K, QGEMM_PATH_TCGEN05);
    void* ws; CUDA_CHECK(cudaMalloc(&ws, wsb));
    if (qgemm_set_default_workspace(ws, wsb) != 0) return 5;
    gemm_q4_0_q8_1(w4, a, c, F, T, K, st);   fetch("c_big.f32");
    printf("last path 0x%x\n", qgemm_last_path());
    qgemm_set_default_workspace(nullptr, 0);
    printf("ok\n");
    return 0;
}
