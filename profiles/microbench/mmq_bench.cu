// mmq_bench.cu -- prefill-path harness over the C ABI: times qgemm_gemm(QGEMM_PATH_TCGEN05) on random blocks
// and checks sampled outputs against the sequential kernel (QGEMM_SEQUENTIAL, bit-identical to the reference GPU
// kernel) on the same device.  A development aid: parity proper lives in tests/ (oracle on the CPU).
//
//   mmq_bench <wtype> <T> <F> <K> [reps] [flags-hex] [layout: 0 = C[F,T] (ggml), 1 = C[T,F] (include/)]
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../include mmq_bench.cu -o mmq_bench \
//        -L../../llama.cpp-quant-gemm_b200/lib -lqgemm_sm100 -Xlinker -rpath -Xlinker '$ORIGIN/../../llama.cpp-quant-gemm_b200/lib'
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "qgemm.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#define QK(x) do { int r_ = (x); if (r_ != 0) { printf("qgemm error %d (%s; %s) at %s:%d\n", r_, qgemm_strerror(r_), qgemm_last_error_detail(), __FILE__, __LINE__); exit(1); } } while (0)

static uint32_t rng_state = 12345u;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

int main(int argc, char** argv) {
    if (argc < 5) { printf("usage: mmq_bench wtype T F K [reps] [flags] [layout]\n"); return 2; }
    const int wt = atoi(argv[1]), T = atoi(argv[2]), F = atoi(argv[3]), K = atoi(argv[4]);
    const int reps = argc > 5 ? atoi(argv[5]) : 5;
    const uint32_t flags = argc > 6 ? (uint32_t)strtoul(argv[6], nullptr, 16) : 0u;
    const int layout = argc > 7 ? atoi(argv[7]) : 0;
    const int nb = K / 32, bs = (int)qgemm_block_bytes(wt);
    const int64_t ldc_t = layout ? F : 1, ldc_f = layout ? 1 : T;

    // random blocks: every nibble / byte value, small positive d, m in [-0.5, 0.5], activations with random d and s
    std::vector<uint8_t> w((size_t)F * nb * bs), a((size_t)T * nb * 36);
    for (auto& v : w) v = (uint8_t)rnd();
    for (auto& v : a) v = (uint8_t)rnd();
    for (size_t i = 0; i < (size_t)F * nb; i++) {
        const __half d = __float2half(0.001f + (rnd() % 1000) * 2e-5f);
        memcpy(&w[i * bs], &d, 2);
        if (wt == QGEMM_TYPE_Q4_1 || wt == QGEMM_TYPE_Q5_1) { const __half m = __float2half(((int)(rnd() % 1001) - 500) * 1e-3f); memcpy(&w[i * bs + 2], &m, 2); }
    }
    for (size_t i = 0; i < (size_t)T * nb; i++) {
        const __half d = __float2half(0.002f + (rnd() % 1000) * 1e-4f), s = __float2half(((int)(rnd() % 8001) - 4000) * 1e-3f);
        memcpy(&a[i * 36], &d, 2);
        memcpy(&a[i * 36 + 2], &s, 2);
    }
    uint8_t *dw, *da; float *dc, *dref; void* ws; uint8_t* flush;
    CK(cudaMalloc(&dw, w.size())); CK(cudaMalloc(&da, a.size()));
    CK(cudaMalloc(&dc, (size_t)T * F * 4)); CK(cudaMalloc(&dref, (size_t)T * F * 4));
    CK(cudaMemcpy(dw, w.data(), w.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice));
    const size_t wsb = qgemm_workspace_bytes(wt, T, F, K, QGEMM_PATH_TCGEN05 | flags);
    CK(cudaMalloc(&ws, wsb ? wsb : 256));
    const size_t flush_bytes = 256u << 20;
    CK(cudaMalloc(&flush, flush_bytes));
    CK(cudaMemset(dc, 0xff, (size_t)T * F * 4));

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0;
    for (int i = 0; i < reps + 1; i++) {
        CK(cudaMemsetAsync(flush, i, flush_bytes));
        CK(cudaEventRecord(e0));
        QK(qgemm_gemm(wt, da, dw, dc, T, F, K, ldc_t, ldc_f, QGEMM_PATH_TCGEN05 | flags, ws, wsb, nullptr));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (i) { best = std::min(best, ms); sum += ms; }
    }
    const double ops = 2.0 * T * F * K;
    printf("wtype %d T %d F %d K %d flags 0x%x: best %.4f ms (%.1f TOPS), mean %.4f ms (%.1f TOPS), path 0x%x, ws %.1f MB\n", wt, T, F, K,
           flags, best, ops / best / 1e9, sum / reps, ops / (sum / reps) / 1e9, qgemm_last_path(), wsb / 1048576.0);

    // check: sampled weight rows x all tokens against the sequential kernel
    const int nrows = std::min(F, 96);
    std::vector<int> rows(nrows);
    for (int i = 0; i < nrows; i++) rows[i] = (i < 32) ? i : (i < 64 ? F - 64 + i : (int)(rnd() % F));
    std::vector<uint8_t> wsub((size_t)nrows * nb * bs);
    for (int i = 0; i < nrows; i++) memcpy(&wsub[(size_t)i * nb * bs], &w[(size_t)rows[i] * nb * bs], (size_t)nb * bs);
    uint8_t* dwsub; CK(cudaMalloc(&dwsub, wsub.size()));
    CK(cudaMemcpy(dwsub, wsub.data(), wsub.size(), cudaMemcpyHostToDevice));
    QK(qgemm_gemm(wt, da, dwsub, dref, T, nrows, K, 1, T, QGEMM_SEQUENTIAL | (flags & QGEMM_MS_EXACT), nullptr, 0, nullptr));   // ref[f][t]
    CK(cudaDeviceSynchronize());
    std::vector<float> hc((size_t)T * F), href((size_t)T * nrows);
    CK(cudaMemcpy(hc.data(), dc, hc.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));
    double maxc = 0, maxd = 0; long nbits = 0, n = 0, nonfinite = 0;
    for (int i = 0; i < nrows; i++)
        for (int t = 0; t < T; t++) {
            const float r = href[(size_t)i * T + t], c = hc[(size_t)t * ldc_t + (size_t)rows[i] * ldc_f];
            if (!std::isfinite(c)) nonfinite++;
            maxc = std::max(maxc, (double)fabsf(r));
            maxd = std::max(maxd, (double)fabsf(r - c));
            uint32_t rb, cb; memcpy(&rb, &r, 4); memcpy(&cb, &c, 4);
            nbits += (rb != cb); n++;
        }
    printf("check vs sequential kernel: max|dC|/max|C| = %.3e, %ld of %ld sampled outputs differ in bits, %ld non-finite -> %s\n",
           maxd / maxc, nbits, n, nonfinite, (maxd / maxc <= 1e-5 && nonfinite == 0) ? "OK" : "FAIL");
    return (maxd / maxc <= 1e-5 && nonfinite == 0) ? 0 : 1;
}
