// gemv.cu -- decode path (T <= 8 tokens per pass): an HBM-bound weight stream.
//
// Replaces every decode-shaped kernel of the reference
// (kernels/gemm/gemm_warp_optimized.cuh:108-1070, gemm_async_copy.cuh:65-232,
// gemm_vectorized.cuh:66-230): those re-derive one idea -- keep the q8_1
// activations close and stream the weight rows once -- and top out at 42 % of
// their GPU's DRAM bandwidth.
//
// Design (B200):
//   * one persistent CTA per SM; weight rows are grouped R at a time and a
//     row group is cut into K-chunks of <= 128 blocks; (group, chunk) = one tile
//   * a producer warp streams tiles HBM -> smem with 1-D bulk async copies
//     (cp.async.bulk, the TMA engine; one copy per row segment, all landing on
//     one mbarrier) through a ring of `stages` buffers -- tens of KB in flight
//     per SM with no registers or L1 involved
//   * the T x K activations are staged once per CTA in smem, re-laid out so
//     that consecutive lanes read consecutive 16-byte quads (conflict-free)
//   * 8 consumer warps: warp -> (row, K-slice) of the tile, lane -> pairs of
//     adjacent blocks (a pair starts 4-byte aligned in every format, which a
//     single 18/22/34-byte block does not); integer dot by dp4a on UN-offset
//     weights, then the reference's exact per-block fold (qgemm_common.cuh)
//   * per-row partial sums: lane-sequential over K, butterfly across lanes,
//     fixed-order across the warps of a row -> deterministic
#include "ptx.cuh"
#include "qgemm_common.cuh"

namespace qgemm {

constexpr int kGemvWarps = 8;                       // consumer warps
constexpr int kGemvThreads = (kGemvWarps + 1) * 32; // + 1 producer warp
constexpr int kGemvMaxStages = 8;
constexpr int kGemvChunkBlocks = 128;               // K-chunk, in 32-element blocks
constexpr int kGemvSmemBudget = 200 * 1024;

// ---- a pair of adjacent weight blocks as 32-bit words --------------------------
template <int WT> struct Pair { static constexpr int words = Fmt<WT>::bytes / 2; };

template <int WT>
__device__ __forceinline__ void load_pair(const uint8_t* p, uint32_t (&x)[Pair<WT>::words]) {
    if constexpr (WT == QGEMM_TYPE_Q5_1) {  // 48 B, 16-byte aligned
        const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const uint4 v = q[i];
            x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
    } else if constexpr (WT == QGEMM_TYPE_Q4_1) {  // 40 B, 8-byte aligned
        const uint2* q = reinterpret_cast<const uint2*>(p);
#pragma unroll
        for (int i = 0; i < 5; i++) {
            const uint2 v = q[i];
            x[2 * i] = v.x; x[2 * i + 1] = v.y;
        }
    } else {  // 36 / 44 / 68 B, 4-byte aligned, odd word stride across lanes
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
        for (int i = 0; i < Pair<WT>::words; i++) x[i] = q[i];
    }
}

__device__ __forceinline__ void expand4(const uint32_t (&q)[4], uint32_t (&w)[8]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        w[i] = q[i] & 0x0f0f0f0fu;
        w[i + 4] = (q[i] >> 4) & 0x0f0f0f0fu;
    }
}
__device__ __forceinline__ void expand5(const uint32_t (&q)[4], uint32_t qh, uint32_t (&w)[8]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        w[i] = (q[i] & 0x0f0f0f0fu) | spread_qh4(qh, 4 * i);
        w[i + 4] = ((q[i] >> 4) & 0x0f0f0f0fu) | spread_qh4(qh, 16 + 4 * i);
    }
}

// x = the pair's words; block 0 starts at byte 0, block 1 at byte Fmt::bytes.
template <int WT>
__device__ __forceinline__ void expand_pair(const uint32_t (&x)[Pair<WT>::words], uint32_t (&w)[2][8], WScale (&ws)[2]) {
    auto fs = [](uint32_t lo, uint32_t hi) { return __funnelshift_r(lo, hi, 16); };
    if constexpr (WT == QGEMM_TYPE_Q4_0) {
        ws[0] = {half_bits_to_float(x[0]), 0.f};
        ws[1] = {half_bits_to_float(x[4] >> 16), 0.f};
        const uint32_t q0[4] = {fs(x[0], x[1]), fs(x[1], x[2]), fs(x[2], x[3]), fs(x[3], x[4])};
        const uint32_t q1[4] = {x[5], x[6], x[7], x[8]};
        expand4(q0, w[0]);
        expand4(q1, w[1]);
    } else if constexpr (WT == QGEMM_TYPE_Q4_1) {
        ws[0] = {half_bits_to_float(x[0]), half_bits_to_float(x[0] >> 16)};
        ws[1] = {half_bits_to_float(x[5]), half_bits_to_float(x[5] >> 16)};
        const uint32_t q0[4] = {x[1], x[2], x[3], x[4]};
        const uint32_t q1[4] = {x[6], x[7], x[8], x[9]};
        expand4(q0, w[0]);
        expand4(q1, w[1]);
    } else if constexpr (WT == QGEMM_TYPE_Q5_0) {
        ws[0] = {half_bits_to_float(x[0]), 0.f};
        ws[1] = {half_bits_to_float(x[5] >> 16), 0.f};
        const uint32_t q0[4] = {fs(x[1], x[2]), fs(x[2], x[3]), fs(x[3], x[4]), fs(x[4], x[5])};
        const uint32_t q1[4] = {x[7], x[8], x[9], x[10]};
        expand5(q0, fs(x[0], x[1]), w[0]);
        expand5(q1, x[6], w[1]);
    } else if constexpr (WT == QGEMM_TYPE_Q5_1) {
        ws[0] = {half_bits_to_float(x[0]), half_bits_to_float(x[0] >> 16)};
        ws[1] = {half_bits_to_float(x[6]), half_bits_to_float(x[6] >> 16)};
        const uint32_t q0[4] = {x[2], x[3], x[4], x[5]};
        const uint32_t q1[4] = {x[8], x[9], x[10], x[11]};
        expand5(q0, x[1], w[0]);
        expand5(q1, x[7], w[1]);
    } else {  // q8_0
        ws[0] = {half_bits_to_float(x[0]), 0.f};
        ws[1] = {half_bits_to_float(x[8] >> 16), 0.f};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            w[0][i] = fs(x[i], x[i + 1]);
            w[1][i] = x[9 + i];
        }
    }
}

struct GemvParams {
    const uint8_t* act;   // q8_1, first token of this pass
    const uint8_t* wgt;
    float* C;             // already offset to the first token of this pass
    int F, nb;
    int64_t ldc_t, ldc_f;
    int R;                // weight rows per tile: 1, 2, 4 or 8
    int stages;
    int stage_bytes;      // 128-byte multiple
};

template <int WT, int TT, bool kMsExact>
__global__ void __launch_bounds__(kGemvThreads, 1) gemv_kernel(const GemvParams p) {
    using Fm = Fmt<WT>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = p.nb, np = nb >> 1;

    // ---- carve shared memory
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [kGemvMaxStages]
    uint64_t* empty = full + kGemvMaxStages;                     // [kGemvMaxStages]
    float* red = reinterpret_cast<float*>(smem + 128);           // [2][kGemvWarps][8]
    float4* a_scale = reinterpret_cast<float4*>(smem + 1024);    // [TT][np]   (d0,s0,d1,s1) of a pair
    uint4* a_qs = reinterpret_cast<uint4*>(a_scale + (size_t)TT * np);  // [TT][4][np] 16-byte quads
    uint8_t* stage0 = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(a_qs + (size_t)TT * 4 * np) + 127) & ~uintptr_t(127));

    if (tid == 0) {
        for (int s = 0; s < p.stages; s++) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], kGemvWarps);
        }
        ptx::fence_mbar_init();
    }
    // ---- stage the activations: q8_1 AoS -> quad-interleaved SoA + fp32 scales
    {
        const uint32_t* a32 = reinterpret_cast<const uint32_t*>(p.act);
        float* sc = reinterpret_cast<float*>(a_scale);
        uint32_t* qs = reinterpret_cast<uint32_t*>(a_qs);
        const int total = TT * nb * 9;
        for (int i = tid; i < total; i += kGemvThreads) {
            const int blk = i / 9, wd = i - blk * 9;
            const int t = blk / nb, b = blk - t * nb;
            const uint32_t v = __ldg(a32 + i);
            if (wd == 0) {
                float* d = sc + ((size_t)t * np + (b >> 1)) * 4 + (b & 1) * 2;
                d[0] = half_bits_to_float(v);
                d[1] = half_bits_to_float(v >> 16);
            } else {
                const int e = wd - 1;                       // word 0..7 of the block's qs
                const int quad = (b & 1) * 2 + (e >> 2);    // 4 quads per pair
                qs[(((size_t)t * 4 + quad) * np + (b >> 1)) * 4 + (e & 3)] = v;
            }
        }
    }
    __syncthreads();

    const int R = p.R;
    const int G = (p.F + R - 1) / R;
    const int nchunks = (nb + kGemvChunkBlocks - 1) / kGemvChunkBlocks;
    const size_t rowbytes = (size_t)nb * Fm::bytes;

    if (warp == kGemvWarps) {
        // ================= producer warp =================
        int it = 0;
        for (int g = blockIdx.x; g < G; g += gridDim.x) {
            for (int c = 0; c < nchunks; c++, it++) {
                const int s = it % p.stages;
                const uint32_t ph = (it / p.stages) & 1;
                const int cb = min(kGemvChunkBlocks, nb - c * kGemvChunkBlocks);
                const uint32_t seg = (uint32_t)cb * Fm::bytes;
                ptx::mbar_wait(&empty[s], ph ^ 1);
                if (lane == 0) ptx::mbar_arrive_expect_tx(&full[s], seg * R);
                __syncwarp();
                if (lane < R) {
                    const int f = min(g * R + lane, p.F - 1);
                    ptx::bulk_g2s(stage0 + (size_t)s * p.stage_bytes + (size_t)lane * seg,
                                  p.wgt + (size_t)f * rowbytes + (size_t)c * kGemvChunkBlocks * Fm::bytes, seg,
                                  &full[s]);
                }
            }
        }
        return;
    }

    // ================= consumer warps =================
    const int WPR = kGemvWarps / R;          // warps sharing one row
    const int row = warp / WPR, sub = warp - row * WPR;
    float acc[TT];
#pragma unroll
    for (int t = 0; t < TT; t++) acc[t] = 0.f;

    int it = 0, gpar = 0;
    for (int g = blockIdx.x; g < G; g += gridDim.x, gpar ^= 1) {
        for (int c = 0; c < nchunks; c++, it++) {
            const int s = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1;
            const int cb = min(kGemvChunkBlocks, nb - c * kGemvChunkBlocks);
            const int npc = cb >> 1;
            const int pair0 = (c * kGemvChunkBlocks) >> 1;
            ptx::mbar_wait(&full[s], ph);
            const uint8_t* rowp = stage0 + (size_t)s * p.stage_bytes + (size_t)row * cb * Fm::bytes;
            for (int pp = sub * 32 + lane; pp < npc; pp += WPR * 32) {
                uint32_t x[Pair<WT>::words];
                uint32_t w[2][8];
                WScale ws[2];
                load_pair<WT>(rowp + (size_t)pp * (2 * Fm::bytes), x);
                expand_pair<WT>(x, w, ws);
                const int pg = pair0 + pp;
#pragma unroll
                for (int t = 0; t < TT; t++) {
                    const float4 sc = a_scale[(size_t)t * np + pg];
                    int a0[8], a1[8];
                    {
                        const uint4 q0 = a_qs[((size_t)t * 4 + 0) * np + pg];
                        const uint4 q1 = a_qs[((size_t)t * 4 + 1) * np + pg];
                        const uint4 q2 = a_qs[((size_t)t * 4 + 2) * np + pg];
                        const uint4 q3 = a_qs[((size_t)t * 4 + 3) * np + pg];
                        a0[0] = q0.x; a0[1] = q0.y; a0[2] = q0.z; a0[3] = q0.w;
                        a0[4] = q1.x; a0[5] = q1.y; a0[6] = q1.z; a0[7] = q1.w;
                        a1[0] = q2.x; a1[1] = q2.y; a1[2] = q2.z; a1[3] = q2.w;
                        a1[4] = q3.x; a1[5] = q3.y; a1[6] = q3.z; a1[7] = q3.w;
                    }
                    const int s0 = block_sumi<WT>(w[0], a0);
                    const int s1 = block_sumi<WT>(w[1], a1);
                    acc[t] = fold_block<WT, kMsExact>(acc[t], s0, ws[0], ActScale{sc.x, sc.y});
                    acc[t] = fold_block<WT, kMsExact>(acc[t], s1, ws[1], ActScale{sc.z, sc.w});
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty[s]);
        }
        // ---- row group finished: combine lanes, then the warps of each row
        float* rbuf = red + gpar * (kGemvWarps * 8);
#pragma unroll
        for (int t = 0; t < TT; t++) {
            float v = acc[t];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) rbuf[warp * 8 + t] = v;
            acc[t] = 0.f;
        }
        ptx::bar_sync(1, kGemvWarps * 32);
        if (tid < R * TT) {
            const int r = tid / TT, t = tid - r * TT;
            float v = 0.f;
            for (int k = 0; k < WPR; k++) v += rbuf[(r * WPR + k) * 8 + t];
            const int f = g * R + r;
            if (f < p.F) p.C[(int64_t)t * p.ldc_t + (int64_t)f * p.ldc_f] = v;
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static size_t gemv_act_bytes(int tt, int nb) { return (size_t)tt * nb * 40; }

// Can the fast path take this problem at all?  Rows must be bulk-copyable.
bool gemv_supported(int wtype, const void* act, const void* wgt, int F, int K) {
    const int nb = K / 32;
    const size_t rowbytes = (size_t)nb * block_bytes(wtype);
    if (F < 1 || nb < 2 || (nb & 1)) return false;
    if (rowbytes % 16 != 0) return false;
    if (((size_t)kGemvChunkBlocks * block_bytes(wtype)) % 16 != 0) return false;
    if (reinterpret_cast<uintptr_t>(wgt) % 16 != 0) return false;
    if (reinterpret_cast<uintptr_t>(act) % 4 != 0) return false;
    // at least one token's activations plus two stages must fit
    const size_t tile = (size_t)min(nb, kGemvChunkBlocks) * block_bytes(wtype);
    return 1024 + gemv_act_bytes(1, nb) + 128 + 2 * ((tile + 127) / 128 * 128) <= (size_t)kGemvSmemBudget;
}

// Tokens per pass (<= 8) that still leave room for a useful ring.
int gemv_tokens_per_pass(int wtype, int T, int K) {
    const int nb = K / 32;
    const size_t tile = (size_t)min(nb, kGemvChunkBlocks) * block_bytes(wtype) * 4;
    int tt = min(T, 8);
    while (tt > 1 && 1024 + gemv_act_bytes(tt, nb) + 128 + 3 * tile > (size_t)kGemvSmemBudget) tt--;
    return tt;
}

template <int WT, int TT>
static cudaError_t launch_gemv_tt(const GemvParams& p, size_t smem, int grid, bool ms_exact, cudaStream_t st) {
    cudaError_t e;
    if (ms_exact) {
        auto k = gemv_kernel<WT, TT, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<grid, kGemvThreads, smem, st>>>(p);
    } else {
        auto k = gemv_kernel<WT, TT, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<grid, kGemvThreads, smem, st>>>(p);
    }
    note_launch();
    return cudaGetLastError();
}

template <int WT>
static cudaError_t launch_gemv_wt(int tt, const GemvParams& p, size_t smem, int grid, bool ms, cudaStream_t st) {
    switch (tt) {
    case 1: return launch_gemv_tt<WT, 1>(p, smem, grid, ms, st);
    case 2: return launch_gemv_tt<WT, 2>(p, smem, grid, ms, st);
    case 3: return launch_gemv_tt<WT, 3>(p, smem, grid, ms, st);
    case 4: return launch_gemv_tt<WT, 4>(p, smem, grid, ms, st);
    case 5: return launch_gemv_tt<WT, 5>(p, smem, grid, ms, st);
    case 6: return launch_gemv_tt<WT, 6>(p, smem, grid, ms, st);
    case 7: return launch_gemv_tt<WT, 7>(p, smem, grid, ms, st);
    case 8: return launch_gemv_tt<WT, 8>(p, smem, grid, ms, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_gemv(int wtype, const void* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                        int64_t ldc_f, uint32_t flags, int num_sms, cudaStream_t st) {
    const int nb = K / 32;
    const int bs = block_bytes(wtype);
    const bool ms = flags & QGEMM_MS_EXACT;
    const int tpp = gemv_tokens_per_pass(wtype, T, K);

    // rows per tile: as many as keeps >= 3 row groups per SM (load balance beats tile size)
    int R = 4;
    while (R > 1 && (F + R - 1) / R < 3 * num_sms) R >>= 1;
    const int cb = min(nb, kGemvChunkBlocks);
    const int stage_bytes = (int)(((size_t)R * cb * bs + 127) / 128 * 128);
    const int G = (F + R - 1) / R;
    const int grid = min(G, num_sms);

    for (int t0 = 0; t0 < T; t0 += tpp) {
        const int tt = min(tpp, T - t0);
        const size_t fixed = 1024 + gemv_act_bytes(tt, nb) + 128;
        int stages = (int)(((size_t)kGemvSmemBudget - fixed) / stage_bytes);
        stages = max(2, min(kGemvMaxStages, stages));
        GemvParams p;
        p.act = (const uint8_t*)act + (size_t)t0 * nb * kQ81Bytes;
        p.wgt = (const uint8_t*)wgt;
        p.C = C + (int64_t)t0 * ldc_t;
        p.F = F; p.nb = nb; p.ldc_t = ldc_t; p.ldc_f = ldc_f;
        p.R = R; p.stages = stages; p.stage_bytes = stage_bytes;
        const size_t smem = fixed + (size_t)stages * stage_bytes;
        cudaError_t e;
        switch (wtype) {
        case QGEMM_TYPE_Q4_0: e = launch_gemv_wt<QGEMM_TYPE_Q4_0>(tt, p, smem, grid, ms, st); break;
        case QGEMM_TYPE_Q4_1: e = launch_gemv_wt<QGEMM_TYPE_Q4_1>(tt, p, smem, grid, ms, st); break;
        case QGEMM_TYPE_Q5_0: e = launch_gemv_wt<QGEMM_TYPE_Q5_0>(tt, p, smem, grid, ms, st); break;
        case QGEMM_TYPE_Q5_1: e = launch_gemv_wt<QGEMM_TYPE_Q5_1>(tt, p, smem, grid, ms, st); break;
        case QGEMM_TYPE_Q8_0: e = launch_gemv_wt<QGEMM_TYPE_Q8_0>(tt, p, smem, grid, ms, st); break;
        default: e = cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace qgemm
