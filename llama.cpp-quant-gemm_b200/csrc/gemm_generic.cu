// gemm_generic.cu -- the always-correct path and the sumi test hook.
//
// QGEMM_PATH_GENERIC / QGEMM_SEQUENTIAL: one thread per output, K-blocks
// accumulated in order b = 0..nb-1 with the reference GPU kernel's exact FMA
// sequence, so C is bit-identical to kernels/gemm/gemm_quant_formats.cuh:312-334
// (and to include/gemm_cuda_naive.cuh:158-249) as nvcc builds them.  It accepts
// any 2-byte aligned weight pointer and any K % 32 == 0, like the reference, and
// is the landing spot for shapes the fast paths decline.  Threads of a warp
// walk consecutive weight rows of one token, so activation loads broadcast.
#include "qgemm_common.cuh"

namespace qgemm {

template <int WT, bool kMsExact>
__global__ void __launch_bounds__(256)
gemm_sequential_kernel(const uint8_t* __restrict__ act, const uint8_t* __restrict__ wgt, float* __restrict__ C,
                       int T, int F, int nb, int64_t ldc_t, int64_t ldc_f) {
    using Fm = Fmt<WT>;
    const int f = blockIdx.x * 32 + (threadIdx.x & 31);
    const int t = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (f >= F || t >= T) return;
    const uint8_t* wrow = wgt + (size_t)f * nb * Fm::bytes;
    const uint8_t* arow = act + (size_t)t * nb * kQ81Bytes;
    float acc = 0.0f;
    for (int b = 0; b < nb; b++) {
        const uint8_t* wb = wrow + (size_t)b * Fm::bytes;
        const uint8_t* ab = arow + (size_t)b * kQ81Bytes;
        uint32_t w[8];
        int a[8];
        unpack_block<WT>(wb, w);
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = *reinterpret_cast<const int*>(ab + 4 + 4 * i);
        const int sumi = block_sumi<WT>(w, a);
        const uint32_t ds = *reinterpret_cast<const uint32_t*>(ab);
        ActScale as{half_bits_to_float(ds), half_bits_to_float(ds >> 16)};
        acc = fold_block<WT, kMsExact>(acc, sumi, load_wscale<WT>(wb), as);
    }
    C[(int64_t)t * ldc_t + (int64_t)f * ldc_f] = acc;
}

template <int WT>
static cudaError_t launch_seq_t(const void* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                                int64_t ldc_f, bool ms_exact, cudaStream_t st) {
    dim3 grid((F + 31) / 32, (T + 7) / 8);
    const uint8_t* a = (const uint8_t*)act;
    const uint8_t* w = (const uint8_t*)wgt;
    if (ms_exact) gemm_sequential_kernel<WT, true><<<grid, 256, 0, st>>>(a, w, C, T, F, K / 32, ldc_t, ldc_f);
    else gemm_sequential_kernel<WT, false><<<grid, 256, 0, st>>>(a, w, C, T, F, K / 32, ldc_t, ldc_f);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_gemm_sequential(int wtype, const void* act, const void* wgt, float* C, int T, int F, int K,
                                   int64_t ldc_t, int64_t ldc_f, uint32_t flags, cudaStream_t st) {
    const bool ms = flags & QGEMM_MS_EXACT;
    switch (wtype) {
    case QGEMM_TYPE_Q4_0: return launch_seq_t<QGEMM_TYPE_Q4_0>(act, wgt, C, T, F, K, ldc_t, ldc_f, ms, st);
    case QGEMM_TYPE_Q4_1: return launch_seq_t<QGEMM_TYPE_Q4_1>(act, wgt, C, T, F, K, ldc_t, ldc_f, ms, st);
    case QGEMM_TYPE_Q5_0: return launch_seq_t<QGEMM_TYPE_Q5_0>(act, wgt, C, T, F, K, ldc_t, ldc_f, ms, st);
    case QGEMM_TYPE_Q5_1: return launch_seq_t<QGEMM_TYPE_Q5_1>(act, wgt, C, T, F, K, ldc_t, ldc_f, ms, st);
    case QGEMM_TYPE_Q8_0: return launch_seq_t<QGEMM_TYPE_Q8_0>(act, wgt, C, T, F, K, ldc_t, ldc_f, ms, st);
    default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------------------
// sumi dump (dp4a integers): sumi[(t*F + f)*nb + b]
// ---------------------------------------------------------------------------
template <int WT>
__global__ void __launch_bounds__(256)
sumi_kernel(const uint8_t* __restrict__ act, const uint8_t* __restrict__ wgt, int32_t* __restrict__ out, int T, int F,
            int nb) {
    using Fm = Fmt<WT>;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)T * F * nb;
    if (gid >= total) return;
    const int b = (int)(gid % nb);
    const int f = (int)((gid / nb) % F);
    const int t = (int)(gid / ((int64_t)nb * F));
    const uint8_t* wb = wgt + ((size_t)f * nb + b) * Fm::bytes;
    const uint8_t* ab = act + ((size_t)t * nb + b) * kQ81Bytes;
    uint32_t w[8];
    int a[8];
    unpack_block<WT>(wb, w);
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = *reinterpret_cast<const int*>(ab + 4 + 4 * i);
    out[gid] = block_sumi<WT>(w, a);
}

cudaError_t launch_sumi_generic(int wtype, const void* act, const void* wgt, int32_t* out, int T, int F, int K,
                                cudaStream_t st) {
    const int nb = K / 32;
    const int64_t total = (int64_t)T * F * nb;
    if (total == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((total + 255) / 256);
    const uint8_t* a = (const uint8_t*)act;
    const uint8_t* w = (const uint8_t*)wgt;
    switch (wtype) {
    case QGEMM_TYPE_Q4_0: sumi_kernel<QGEMM_TYPE_Q4_0><<<grid, 256, 0, st>>>(a, w, out, T, F, nb); break;
    case QGEMM_TYPE_Q4_1: sumi_kernel<QGEMM_TYPE_Q4_1><<<grid, 256, 0, st>>>(a, w, out, T, F, nb); break;
    case QGEMM_TYPE_Q5_0: sumi_kernel<QGEMM_TYPE_Q5_0><<<grid, 256, 0, st>>>(a, w, out, T, F, nb); break;
    case QGEMM_TYPE_Q5_1: sumi_kernel<QGEMM_TYPE_Q5_1><<<grid, 256, 0, st>>>(a, w, out, T, F, nb); break;
    case QGEMM_TYPE_Q8_0: sumi_kernel<QGEMM_TYPE_Q8_0><<<grid, 256, 0, st>>>(a, w, out, T, F, nb); break;
    default: return cudaErrorInvalidValue;
    }
    note_launch();
    return cudaGetLastError();
}

}  // namespace qgemm
