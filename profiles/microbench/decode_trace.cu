// decode_trace.cu -- timeline of the bench's decode launch chain, from inside the kernels.  Needs the library built with
// -DQGEMM_GEMV_TRACE (gemv.cu stamps %globaltimer per CTA: entry, dependency wait passed, activations in registers, last
// store) preloaded in front of the product library; a development aid, not a bench.
//   build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../include decode_trace.cu -o decode_trace \
//          -L../../llama.cpp-quant-gemm_b200/lib -lqgemm_sm100 -Xlinker -rpath -Xlinker '$ORIGIN/../../llama.cpp-quant-gemm_b200/lib'
//   run:   LD_PRELOAD=<trace build of libqgemm_sm100.so> ./decode_trace
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <vector>
#include <cuda_runtime.h>
#include "qgemm.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
#define QK(x) do { int r_ = (x); if (r_ != 0) { printf("qgemm error %d (%s) at line %d\n", r_, qgemm_strerror(r_), __LINE__); return 1; } } while (0)

int main() {
    typedef int (*read_fn)(unsigned long long*, size_t);
    read_fn rd = (read_fn)dlsym(RTLD_DEFAULT, "qgemm_debug_read_gemv_trace");
    if (!rd) { printf("preload a -DQGEMM_GEMV_TRACE build of the library\n"); return 2; }
    const int layers = 32, per = 7;
    const int Fs[per] = {4096, 4096, 4096, 4096, 11008, 11008, 4096}, Ks[per] = {4096, 4096, 4096, 4096, 4096, 4096, 11008};
    std::vector<size_t> off(layers * per + 1, 0);
    for (int i = 0; i < layers * per; i++) off[i + 1] = off[i] + ((size_t)Fs[i % per] * (Ks[i % per] / 32) * 18 + 255) / 256 * 256;
    uint8_t* arena; CK(cudaMalloc(&arena, off.back()));
    CK(cudaMemset(arena, 0x11, off.back()));
    uint8_t *a4k, *a11k; float* out;
    CK(cudaMalloc(&a4k, 128 * 36)); CK(cudaMalloc(&a11k, 344 * 36)); CK(cudaMalloc(&out, 4 * 11008 * sizeof(float)));
    CK(cudaMemset(a4k, 0, 128 * 36)); CK(cudaMemset(a11k, 0, 344 * 36));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    const size_t window = 12u << 20;
    auto hint = [&](int idx) { const size_t o = off[idx % (layers * per)]; qgemm_hint_next_weights(arena + o, std::min(window, off.back() - o)); };
    auto step = [&]() -> int {
        for (int l = 0; l < layers; l++) {
            const int b = l * per;
            const void* g3[3] = {arena + off[b], arena + off[b + 1], arena + off[b + 2]};
            float* c3[3] = {out, out + 4096, out + 8192};
            const int f3[3] = {4096, 4096, 4096};
            hint(b + 3);
            QK(qgemm_gemm_group(QGEMM_TYPE_Q4_0, a4k, 3, g3, c3, f3, 1, 4096, 1, 1, QGEMM_WEIGHTS_STATIC, st));
            hint(b + 4);
            QK(qgemm_gemm(QGEMM_TYPE_Q4_0, a4k, arena + off[b + 3], out, 1, 4096, 4096, 1, 1, QGEMM_WEIGHTS_STATIC, nullptr, 0, st));
            const void* g2[2] = {arena + off[b + 4], arena + off[b + 5]};
            float* c2[2] = {out, out + 11008};
            const int f2[2] = {11008, 11008};
            hint(b + 6);
            QK(qgemm_gemm_group(QGEMM_TYPE_Q4_0, a4k, 2, g2, c2, f2, 1, 4096, 1, 1, QGEMM_WEIGHTS_STATIC, st));
            hint(b + 7);
            QK(qgemm_gemm(QGEMM_TYPE_Q4_0, a11k, arena + off[b + 6], out, 1, 4096, 11008, 1, 1, QGEMM_WEIGHTS_STATIC, nullptr, 0, st));
        }
        return 0;
    };
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    if (step()) return 1;
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int r = 0; r < 3; r++) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e0, st));
    CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("one step (128 launches): %.1f us\n", ms * 1e3);
    const int slots = 256, ctas = 304, nl = layers * 4;
    std::vector<unsigned long long> tr((size_t)slots * ctas * 6);
    if (rd(tr.data(), tr.size() * 8)) { printf("trace read failed\n"); return 1; }
    // the capture pass used slots 0..127, the graph nodes keep those slot numbers: the last replay is what is in them
    auto T = [&](int l, int c, int i) { return tr[((size_t)(l % slots) * ctas + c) * 6 + i]; };
    const int grid = 296;
    double sum_wait2bulk = 0, sum_wait2bulk_mean = 0, sum_gap = 0, sum_skew = 0, sum_startskew = 0, sum_wait2act = 0, sum_len = 0, sum_period = 0;
    int cnt = 0;
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < grid; c++) t0 = std::min(t0, T(0, c, 0));
    for (int l = 0; l < nl; l++) {
        unsigned long long s_min = ~0ull, s_max = 0, w_min = ~0ull, w_max = 0, a_max = 0, e_min = ~0ull, e_max = 0;
        for (int c = 0; c < grid; c++) {
            s_min = std::min(s_min, T(l, c, 0)); s_max = std::max(s_max, T(l, c, 0));
            w_min = std::min(w_min, T(l, c, 1)); w_max = std::max(w_max, T(l, c, 1));
            a_max = std::max(a_max, T(l, c, 2));
            e_min = std::min(e_min, T(l, c, 3)); e_max = std::max(e_max, T(l, c, 3));
        }
        if (l >= 8 && l < 24)
            printf("launch %3d (%s): first CTA in %7.2f  last CTA in %7.2f | wait passed %7.2f .. %7.2f | acts ready (last) %7.2f | first CTA done %7.2f  last CTA done %7.2f us\n",
                   l, (const char*[]){"qkv", "wo ", "g/u", "dwn"}[l % 4], (s_min - t0) * 1e-3, (s_max - t0) * 1e-3, (w_min - t0) * 1e-3, (w_max - t0) * 1e-3,
                   (a_max - t0) * 1e-3, (e_min - t0) * 1e-3, (e_max - t0) * 1e-3);
        if (l >= 4 && l + 1 < nl) {
            unsigned long long nw_min = ~0ull, ne_max = 0;
            for (int c = 0; c < grid; c++) { nw_min = std::min(nw_min, T(l + 1, c, 1)); ne_max = std::max(ne_max, T(l + 1, c, 3)); }
            sum_gap += (double)nw_min - (double)e_max;          // last store of launch l -> first CTA of l+1 past its wait
            sum_skew += (double)e_max - (double)e_min;
            sum_startskew += (double)w_max - (double)w_min;
            sum_wait2act += (double)a_max - (double)w_max;
            { unsigned long long b_max = 0; double m = 0; for (int c = 0; c < grid; c++) { b_max = std::max(b_max, T(l, c, 4)); m += (double)T(l, c, 4) - (double)T(l, c, 1); } sum_wait2bulk += (double)b_max - (double)w_max; sum_wait2bulk_mean += m / grid; }
            sum_len += (double)e_max - (double)w_min;
            sum_period += (double)ne_max - (double)e_max;
            cnt++;
        }
    }
    // do the CTAs that come in late (their slot's predecessor finished late) also finish late?
    for (int l = 8; l < 12; l++) {
        std::vector<std::pair<unsigned long long, unsigned long long>> se;
        for (int c = 0; c < grid; c++) se.push_back({T(l, c, 0), T(l, c, 3)});
        std::sort(se.begin(), se.end());
        double q[4] = {0, 0, 0, 0};
        for (int c = 0; c < grid; c++) q[c * 4 / grid] += (double)(se[c].second - t0) * 1e-3 / (grid / 4);
        int lo = 0, hi = 0;   // CTAs 0..147 vs 148..295 by index: mean end time
        double elo = 0, ehi = 0;
        for (int c = 0; c < grid; c++) { if (c < grid / 2) { elo += (double)(T(l, c, 3) - t0) * 1e-3; lo++; } else { ehi += (double)(T(l, c, 3) - t0) * 1e-3; hi++; } }
        printf("launch %d: mean end time by start-time quartile (earliest starters first): %.2f %.2f %.2f %.2f us | by block index half: %.2f %.2f\n", l, q[0], q[1], q[2], q[3],
               elo / lo, ehi / hi);
    }
    printf("wait passed -> activation bulk copy complete: last CTA %.2f us, mean over CTAs %.2f us\n", sum_wait2bulk / cnt * 1e-3, sum_wait2bulk_mean / cnt * 1e-3);
    printf("averages over %d launches (us): period %.2f | wait passed -> last store %.2f | last store -> next launch past its wait %.2f | "
           "skew of the waits %.2f | last wait -> activations in registers %.2f | skew of the last stores %.2f\n",
           cnt, sum_period / cnt * 1e-3, sum_len / cnt * 1e-3, sum_gap / cnt * 1e-3, sum_startskew / cnt * 1e-3, sum_wait2act / cnt * 1e-3, sum_skew / cnt * 1e-3);
    return 0;
}
