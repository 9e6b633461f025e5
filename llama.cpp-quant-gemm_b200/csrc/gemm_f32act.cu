// gemm_f32act.cu -- W4A16 / W8A16: fp32 activations against block-quantized weights, NO activation quantization.
//
//     C[t, f] = sum_k A[t, k] * dequant(W[f, k])          dequant = d * (q - 8) (q4_0),  d * q (q8_0)
//
// Replaces gemm_w4a16_naive / gemm_w8a16_naive (include/gemm_cuda_naive.cuh:66-143) and the python extension's
// gemm_q4_0_fp32 (python/quant_gemm/csrc/gemm_ops.cu:271-463), all one-thread-per-output kernels that re-read a whole
// weight row and activation row per output.  Arithmetic: every product and sum in fp32 like the reference; the
// per-block scale is applied once per block, d * sum_k a_k * (q_k - 8), and K is summed in a different (parallel)
// order, so results agree with the reference's sequential FMA chain to ~1e-6 of max|C| (tests: <= 1e-5), not bit for bit.
//
//   T <= 8  (decode): weight stream against activations held in shared memory (f32act_gemv_smem_kernel), or, when they do
//            not fit, re-read coalesced through L1 (f32act_gemv_kernel).  fp32 activations cost 4 bytes of on-chip traffic
//            per weight element and token, so unlike the q8_1 path this one is bound by the load/store pipe, not by HBM.
//   T  > 8  (prefill): 64 rows x 64 tokens register-tiled fp32 GEMM; weights are dequantized into shared memory one
//            block column (32 k) at a time.  CUDA cores only: fp32 x fp32 has no exact tensor-core form short of a
//            3-way split; this path is a neighbour of the hot path, not the hot path.
#include <algorithm>

#include "qgemm_common.cuh"

namespace qgemm {

// packed fp32x2 helpers
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2r(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2r(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// byte i of `w` (an unsigned value n < 256) as the float 2^23 + n; adding -(2^23 + off) afterwards leaves n - off exactly
template <int I>
__device__ __forceinline__ float byte_as_biased_float(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | I));
}

// One weight block -> 32 fp32 values q - off (element order 0..31), as 16 packed pairs (k, k + 1).
template <int WT>
__device__ __forceinline__ void block_to_float(const uint32_t (&w)[8], uint64_t (&v)[16]) {
    // w: 8 words of 4 unsigned bytes each (q4_0: nibbles 0..15; q8_0: the s8 bytes XOR 0x80)
    constexpr float kOff = (WT == QGEMM_TYPE_Q4_0) ? 8.0f : 128.0f;
    const uint64_t bias = pk2(-(8388608.0f + kOff), -(8388608.0f + kOff));
#pragma unroll
    for (int i = 0; i < 8; i++) {
        v[2 * i] = fadd2r(pk2(byte_as_biased_float<0>(w[i]), byte_as_biased_float<1>(w[i])), bias);
        v[2 * i + 1] = fadd2r(pk2(byte_as_biased_float<2>(w[i]), byte_as_biased_float<3>(w[i])), bias);
    }
}

// words of one block from a pair's word run x (block 0 at byte 0, block 1 at byte Fmt::bytes): unsigned bytes + d
template <int WT, int NW>
__device__ __forceinline__ void pair_block_words(const uint32_t (&x)[NW], int which, uint32_t (&w)[8], float& d) {
    auto fs = [](uint32_t lo, uint32_t hi) { return __funnelshift_r(lo, hi, 16); };
    if constexpr (WT == QGEMM_TYPE_Q4_0) {
        uint32_t q[4];
        if (which == 0) {
            d = half_bits_to_float(x[0]);
            q[0] = fs(x[0], x[1]); q[1] = fs(x[1], x[2]); q[2] = fs(x[2], x[3]); q[3] = fs(x[3], x[4]);
        } else {
            d = half_bits_to_float(x[4] >> 16);
            q[0] = x[5]; q[1] = x[6]; q[2] = x[7]; q[3] = x[8];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            w[i] = q[i] & 0x0f0f0f0fu;
            w[i + 4] = (q[i] >> 4) & 0x0f0f0f0fu;
        }
    } else {   // q8_0: 34-byte blocks
        if (which == 0) {
            d = half_bits_to_float(x[0]);
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = fs(x[i], x[i + 1]) ^ 0x80808080u;
        } else {
            d = half_bits_to_float(x[8] >> 16);
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = x[9 + i] ^ 0x80808080u;
        }
    }
}

// ---------------------------------------------------------------------------
// decode: T <= 8
// ---------------------------------------------------------------------------
constexpr int kFaWarps = 8;

template <int WT, int TT>
__global__ void __launch_bounds__(kFaWarps * 32) f32act_gemv_kernel(const float* __restrict__ act, const uint8_t* __restrict__ wgt,
                                                                   float* __restrict__ C, int F, int K, int64_t ldc_t, int64_t ldc_f) {
    // Activations too large for shared memory (T * K * 4 > ~200 KB).  One warp per weight row.  A warp step covers four
    // adjacent blocks (128 k): lane = (block b of the four, float4 j of its 32 elements), so the activations are read as one
    // coalesced 512-byte run per token and the 72 / 136 weight bytes as adjacent half-words.  q4_0: elements 4j .. 4j+3 are
    // the low (j < 4) or high (j >= 4) nibbles of bytes 4(j & 3) .. of the block.  (A first version gave every lane whole
    // blocks and read their activations from global memory: 128-byte private reads, 32 L1 wavefronts per instruction,
    // 0.56 TB/s.)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = lane & 7, bsub = lane >> 3;
    const int nb = K / 32;
    const size_t rowbytes = (size_t)nb * Fmt<WT>::bytes;
    constexpr float kOff = (WT == QGEMM_TYPE_Q4_0) ? 8.0f : 128.0f;
    const uint64_t bias = pk2(-(8388608.0f + kOff), -(8388608.0f + kOff));
    for (int f = blockIdx.x * kFaWarps + warp; f < F; f += gridDim.x * kFaWarps) {
        const uint8_t* row = wgt + (size_t)f * rowbytes;
        float acc[TT];
#pragma unroll
        for (int t = 0; t < TT; t++) acc[t] = 0.f;
#pragma unroll 4
        for (int b = bsub; b < nb; b += 4) {
            const uint8_t* blk = row + (size_t)b * Fmt<WT>::bytes;
            const float d = ld_half(blk);
            uint32_t u;
            if constexpr (WT == QGEMM_TYPE_Q4_0) {
                const uint8_t* q = blk + 2 + 4 * (j & 3);
                u = ld_u16(q) | (ld_u16(q + 2) << 16);
                u = ((j & 4) ? (u >> 4) : u) & 0x0f0f0f0fu;
            } else {
                const uint8_t* q = blk + 2 + 4 * j;
                u = (ld_u16(q) | (ld_u16(q + 2) << 16)) ^ 0x80808080u;
            }
            const uint64_t w01 = fadd2r(pk2(byte_as_biased_float<0>(u), byte_as_biased_float<1>(u)), bias);
            const uint64_t w23 = fadd2r(pk2(byte_as_biased_float<2>(u), byte_as_biased_float<3>(u)), bias);
            const int k0 = b * 32 + 4 * j;
#pragma unroll
            for (int t = 0; t < TT; t++) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(act + (size_t)t * K + k0));
                uint64_t s2 = ffma2r(pk2(a.x, a.y), w01, 0ull);   // (even-k partial, odd-k partial)
                s2 = ffma2r(pk2(a.z, a.w), w23, s2);
                float s0, s1;
                unpk2(s2, s0, s1);
                acc[t] = __fmaf_rn(d, s0 + s1, acc[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < TT; t++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
        }
        if (lane < TT) {
            float v = acc[0];
#pragma unroll
            for (int t = 1; t < TT; t++) v = (lane == t) ? acc[t] : v;
            C[(int64_t)lane * ldc_t + (int64_t)f * ldc_f] = v;
        }
    }
}

// Decode with the activations resident in shared memory (T * K * 4 bytes <= ~200 KB: every Llama shape up to T = 4, and
// K = 4096 up to T = 8).  Persistent CTAs, one warp per weight row, a lane owns PAIRS of adjacent blocks (a pair starts
// 4-byte aligned: 9 / 17 words), so a warp instruction dequantizes and multiplies 32 blocks at once -- five times fewer
// instructions per element than the kernel above.  Weights go nibble -> fp32 with one byte-permute and one packed add per
// two values (0x4B000000 | n is the float 2^23 + n).  The activations are stored float4-wise with the float4 index inside
// a block XORed by the pair index, so that the lanes of a quarter warp (8 consecutive pairs, same float4 index) read 8
// different bank groups.  The words of a row's first two pair groups are requested before either is used.
constexpr int kFsWarps = 16;

template <int WT, int TT>
__global__ void __launch_bounds__(kFsWarps * 32) f32act_gemv_smem_kernel(const float* __restrict__ act, const uint8_t* __restrict__ wgt,
                                                                        float* __restrict__ C, int F, int K, int64_t ldc_t, int64_t ldc_f) {
    extern __shared__ float4 sA[];   // [TT][K / 4], swizzled inside every block of 8
    constexpr int kPairWords = Fmt<WT>::bytes / 2;   // 9 (q4_0) or 17 (q8_0) words per pair of blocks
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = K / 32, np = nb / 2, ng = K / 4;
    for (int g = threadIdx.x; g < TT * ng; g += blockDim.x) {
        const int t = g / ng, q = g - t * ng, b = q >> 3, i = q & 7;
        sA[t * ng + (b << 3) + (i ^ ((b >> 1) & 7))] = __ldg(reinterpret_cast<const float4*>(act) + g);
    }
    __syncthreads();
    const size_t rowbytes = (size_t)nb * Fmt<WT>::bytes;
    float acc[TT];
    auto fold_pair_words = [&](const uint32_t (&x)[kPairWords], int pg) {
#pragma unroll
        for (int which = 0; which < 2; which++) {
            uint32_t w[8];
            float d;
            pair_block_words<WT, kPairWords>(x, which, w, d);
            uint64_t v[16];
            block_to_float<WT>(w, v);
            const int b = 2 * pg + which, sw = pg & 7;
#pragma unroll
            for (int t = 0; t < TT; t++) {
                const float4* a4 = sA + t * ng + (b << 3);
                uint64_t s2 = 0ull;   // (even-k partial, odd-k partial)
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float4 a = a4[i ^ sw];
                    s2 = ffma2r(pk2(a.x, a.y), v[2 * i], s2);
                    s2 = ffma2r(pk2(a.z, a.w), v[2 * i + 1], s2);
                }
                float s0, s1;
                unpk2(s2, s0, s1);
                acc[t] = __fmaf_rn(d, s0 + s1, acc[t]);
            }
        }
    };
    for (int f = blockIdx.x * kFsWarps + warp; f < F; f += gridDim.x * kFsWarps) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(wgt + (size_t)f * rowbytes);   // 4-byte aligned: nb even
#pragma unroll
        for (int t = 0; t < TT; t++) acc[t] = 0.f;
        constexpr bool kTwoDeep = kPairWords <= 9;   // q8_0's 17-word pairs would spill
        for (int pg = lane; pg < np; pg += (kTwoDeep ? 64 : 32)) {
            uint32_t x0[kPairWords], x1[kTwoDeep ? kPairWords : 1];
            const bool two = kTwoDeep && pg + 32 < np;
#pragma unroll
            for (int i = 0; i < kPairWords; i++) x0[i] = __ldcs(row + (size_t)pg * kPairWords + i);   // streamed once
            if constexpr (kTwoDeep) {
                if (two) {
#pragma unroll
                    for (int i = 0; i < kPairWords; i++) x1[i] = __ldcs(row + (size_t)(pg + 32) * kPairWords + i);
                }
            }
            fold_pair_words(x0, pg);
            if constexpr (kTwoDeep) {
                if (two) fold_pair_words(x1, pg + 32);
            }
        }
#pragma unroll
        for (int t = 0; t < TT; t++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
        }
        if (lane < TT) {
            float v = acc[0];
#pragma unroll
            for (int t = 1; t < TT; t++) v = (lane == t) ? acc[t] : v;
            C[(int64_t)lane * ldc_t + (int64_t)f * ldc_f] = v;
        }
    }
}

// ---------------------------------------------------------------------------
// prefill: T > 8.  64 x 64 output tile, K in steps of one block (32)
// ---------------------------------------------------------------------------
constexpr int kGaTile = 64, kGaPitch = kGaTile + 4;

template <int WT>
__global__ void __launch_bounds__(256) f32act_gemm_kernel(const float* __restrict__ act, const uint8_t* __restrict__ wgt,
                                                         float* __restrict__ C, int T, int F, int K, int64_t ldc_t, int64_t ldc_f) {
    __shared__ __align__(16) float sW[2][32][kGaPitch];   // [k][row]   dequantized q - off times d
    __shared__ __align__(16) float sA[2][32][kGaPitch];   // [k][token]
    const int tid = threadIdx.x;
    const int f0 = blockIdx.x * kGaTile, t0 = blockIdx.y * kGaTile;
    const int nb = K / 32;
    const size_t rowbytes = (size_t)nb * Fmt<WT>::bytes;
    const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads, 4 rows x 4 tokens each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    auto load_stage = [&](int b, int buf) {
        // weights: threads 0..63 dequantize one block each (row f0 + tid)
        if (tid < kGaTile) {
            const int f = f0 + tid;
            float vals[32];
            if (f < F) {
                const uint8_t* blk = wgt + (size_t)f * rowbytes + (size_t)b * Fmt<WT>::bytes;
                const float d = ld_half(blk);
                if constexpr (WT == QGEMM_TYPE_Q4_0) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t q = ld_u16(blk + 2 + 2 * i);
                        vals[2 * i] = (float)((int)(q & 15) - 8) * d;
                        vals[2 * i + 1] = (float)((int)((q >> 8) & 15) - 8) * d;
                        vals[2 * i + 16] = (float)((int)((q >> 4) & 15) - 8) * d;
                        vals[2 * i + 17] = (float)((int)((q >> 12) & 15) - 8) * d;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const uint32_t q = ld_u16(blk + 2 + 2 * i);
                        vals[2 * i] = (float)(int)(int8_t)(q & 0xff) * d;
                        vals[2 * i + 1] = (float)(int)(int8_t)(q >> 8) * d;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 32; k++) vals[k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 32; k++) sW[buf][k][tid] = vals[k];
        }
        // activations: 64 tokens x 32 floats, 8 floats per thread (two float4), stored k-major
        {
            const int t = tid >> 2, kq = (tid & 3) * 8;
            const int tt = t0 + t;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            if (tt < T) {
                const float4* src = reinterpret_cast<const float4*>(act + (size_t)tt * K + (size_t)b * 32 + kq);
                a0 = __ldg(src);
                a1 = __ldg(src + 1);
            }
            sA[buf][kq + 0][t] = a0.x; sA[buf][kq + 1][t] = a0.y; sA[buf][kq + 2][t] = a0.z; sA[buf][kq + 3][t] = a0.w;
            sA[buf][kq + 4][t] = a1.x; sA[buf][kq + 5][t] = a1.y; sA[buf][kq + 6][t] = a1.z; sA[buf][kq + 7][t] = a1.w;
        }
    };

    load_stage(0, 0);
    __syncthreads();
    for (int b = 0; b < nb; b++) {
        const int buf = b & 1;
        if (b + 1 < nb) load_stage(b + 1, buf ^ 1);
#pragma unroll
        for (int k = 0; k < 32; k++) {
            const float4 wv = *reinterpret_cast<const float4*>(&sW[buf][k][ty * 4]);
            const float4 av = *reinterpret_cast<const float4*>(&sA[buf][k][tx * 4]);
            const float wr[4] = {wv.x, wv.y, wv.z, wv.w}, ar[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = __fmaf_rn(ar[j], wr[i], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int f = f0 + ty * 4 + i;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int t = t0 + tx * 4 + j;
            if (f < F && t < T) C[(int64_t)t * ldc_t + (int64_t)f * ldc_f] = acc[i][j];
        }
    }
}

// ---------------------------------------------------------------------------
// sequential: one thread per output, k in the reference's order with its FMA contraction -- bit-identical to
// gemm_w4a16_naive_kernel / gemm_w8a16_naive_kernel built by nvcc (include/gemm_cuda_naive.cuh:66-143).  Also the
// landing spot for shapes the fast kernels do not take (K % 64 != 0, unaligned pointers).
// ---------------------------------------------------------------------------
template <int WT>
__global__ void __launch_bounds__(256) f32act_seq_kernel(const float* __restrict__ act, const uint8_t* __restrict__ wgt,
                                                        float* __restrict__ C, int T, int F, int K, int64_t ldc_t, int64_t ldc_f) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)T * F) return;
    const int f = (int)(i % F), t = (int)(i / F);
    const int nb = K / 32;
    const float* a = act + (size_t)t * K;
    const uint8_t* row = wgt + (size_t)f * nb * Fmt<WT>::bytes;
    float sum = 0.0f;
    for (int b = 0; b < nb; b++) {
        const uint8_t* blk = row + (size_t)b * Fmt<WT>::bytes;
        const float d = ld_half(blk);
        const float* ab = a + b * 32;
        if constexpr (WT == QGEMM_TYPE_Q4_0) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t packed = blk[2 + k];
                const float w0 = __fmul_rn((float)((int)(packed & 0x0F) - 8), d);
                const float w1 = __fmul_rn((float)((int)(packed >> 4) - 8), d);
                sum = __fmaf_rn(ab[k], w0, sum);
                sum = __fmaf_rn(ab[k + 16], w1, sum);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 32; k++) {
                const float w = __fmul_rn((float)(int)(int8_t)blk[2 + k], d);
                sum = __fmaf_rn(ab[k], w, sum);
            }
        }
    }
    C[(int64_t)t * ldc_t + (int64_t)f * ldc_f] = sum;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
bool f32act_supported(int wtype, const void* act, const void* wgt, int K) {
    if (wtype != QGEMM_TYPE_Q4_0 && wtype != QGEMM_TYPE_Q8_0) return false;   // the formats the reference has for A16
    if (K < 64 || K % 64 != 0) return false;                                  // pairs of blocks, float4 activation rows
    return reinterpret_cast<uintptr_t>(act) % 16 == 0 && reinterpret_cast<uintptr_t>(wgt) % 4 == 0;
}

template <int WT>
static cudaError_t launch_f32act_t(const float* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t, int64_t ldc_f,
                                   int num_sms, cudaStream_t st) {
    const uint8_t* w = (const uint8_t*)wgt;
    if (T <= 8) {
        const size_t smem = (size_t)T * K * sizeof(float);
        if (smem <= 200 * 1024 && !QGEMM_ENV("QGEMM_A16_NO_SMEM")) {
            const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / (kFsWarps * 32), (220 * 1024) / (smem + 1024)));
            const int grid = min((F + kFsWarps - 1) / kFsWarps, num_sms * per_sm);
#define QG_FS(TTv)                                                                                                   \
    case TTv:                                                                                                        \
        if (cudaError_t e = smem_optin(reinterpret_cast<const void*>(f32act_gemv_smem_kernel<WT, TTv>), smem)) return e; \
        f32act_gemv_smem_kernel<WT, TTv><<<grid, kFsWarps * 32, smem, st>>>(act, w, C, F, K, ldc_t, ldc_f);             \
        break;
            switch (T) { QG_FS(1) QG_FS(2) QG_FS(3) QG_FS(4) QG_FS(5) QG_FS(6) QG_FS(7) QG_FS(8) default: return cudaErrorInvalidValue; }
#undef QG_FS
            note_launch();
            return cudaGetLastError();
        }
        const int grid = min((F + kFaWarps - 1) / kFaWarps, num_sms * 8);
#define QG_FA(TTv) case TTv: f32act_gemv_kernel<WT, TTv><<<grid, kFaWarps * 32, 0, st>>>(act, w, C, F, K, ldc_t, ldc_f); break;
        switch (T) { QG_FA(1) QG_FA(2) QG_FA(3) QG_FA(4) QG_FA(5) QG_FA(6) QG_FA(7) QG_FA(8) default: return cudaErrorInvalidValue; }
#undef QG_FA
    } else {
        const dim3 grid((F + kGaTile - 1) / kGaTile, (T + kGaTile - 1) / kGaTile);
        f32act_gemm_kernel<WT><<<grid, 256, 0, st>>>(act, w, C, T, F, K, ldc_t, ldc_f);
    }
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_gemm_f32act_sequential(int wtype, const float* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                                          int64_t ldc_f, cudaStream_t st) {
    const int64_t n = (int64_t)T * F;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (wtype == QGEMM_TYPE_Q4_0) f32act_seq_kernel<QGEMM_TYPE_Q4_0><<<grid, 256, 0, st>>>(act, (const uint8_t*)wgt, C, T, F, K, ldc_t, ldc_f);
    else if (wtype == QGEMM_TYPE_Q8_0) f32act_seq_kernel<QGEMM_TYPE_Q8_0><<<grid, 256, 0, st>>>(act, (const uint8_t*)wgt, C, T, F, K, ldc_t, ldc_f);
    else return cudaErrorInvalidValue;
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_gemm_f32act_dequant(int wtype, const float* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                                       int64_t ldc_f, int num_sms, cudaStream_t st) {
    switch (wtype) {
    case QGEMM_TYPE_Q4_0: return launch_f32act_t<QGEMM_TYPE_Q4_0>(act, wgt, C, T, F, K, ldc_t, ldc_f, num_sms, st);
    case QGEMM_TYPE_Q8_0: return launch_f32act_t<QGEMM_TYPE_Q8_0>(act, wgt, C, T, F, K, ldc_t, ldc_f, num_sms, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace qgemm
