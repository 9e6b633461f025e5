// quantize.cu -- fp32 -> block_q8_1 (the activation quantizer on the hot path),
// the weight quantizers used to make test data, and the dequantizers.
//
// Reference semantics: include/quantize.h:165-193 (CPU), :302-337 (GPU),
// tests/framework/test_framework.cuh:195-225, python ext gemm_ops.cu:75-110.
#include "qgemm_common.cuh"

namespace qgemm {

// roundf() for |v| < 2^23 without the libdevice slow path: truncate, then add
// +-1 when the (exactly representable) remainder is at least one half.
__device__ __forceinline__ int round_half_away(float v) {
    const float t = truncf(v);
    const float r = v - t;  // exact
    int q = __float2int_rz(t);
    if (fabsf(r) >= 0.5f) q += (v < 0.0f) ? -1 : 1;
    return q;
}

// ---------------------------------------------------------------------------
// quantize_q8_1: one warp owns 32 consecutive blocks (1024 floats).
//   phase 1: 8 coalesced float4 loads per lane -> smem tile [32][33]
//   phase 2: lane b walks row b IN ORDER (j = 0..31), so `sum` is the same
//            sequential fp32 sum the reference computes (s must be bit-equal)
//   phase 3: 36-byte blocks staged in smem, written back as coalesced words
// HBM-bound: 4 B read + 1.125 B written per element.
// ---------------------------------------------------------------------------
constexpr int kQWarps = 8;

// kTiles: instead of 36-byte blocks the kernel writes what the tensor-core GEMM kernels read -- the s8 values in the
// 128-byte-swizzled K-major operand tiles a8[K/128][Tpad][128] and (d_a, coef * s_a) in the slabs as[Tpad/128][nb][128]
// (layout of mmq.cu's activation prepass) -- so that the fp32-activation GEMM is two launches with no q8_1 round trip
// through HBM.  d, s and q are the same values, rounded the same way, as in the block_q8_1 the plain kernel emits.
struct Q81Tiles {
    uint8_t* a8;
    float2* as;
    int T, Tpad, nb;   // nb % 4 == 0
    float coef;
    unsigned* zero;    // split-K arrival counters of the GEMM behind this kernel: cleared here
    int nzero;
    const float* gate; // kPro == 1: the second operand of silu(x) * gate;  kPro == 2: the rms_norm weight [K]
    const float* inv_rms; // kPro == 2: 1 / rms of every row (row_inv_rms_kernel)
    int pro_nb;        // kPro == 2: blocks per row
};

// kSiluMul: the value that is quantized is silu(x) * gate, computed with the operation sequence of the reference's
// silu_mul_f32_kernel (kernels/activation/silu.cuh:97-108): val / (1.0f + expf(-val)), then times gate -- the SwiGLU
// neighbour of the FFN down projection folded into its quantizer, 9.1 instead of 17.1 bytes per element.
// kPro == 2: the value that is quantized is x * inv_rms[row] * weight[col], the operation order of the reference's
// rms_norm kernels (kernels/normalization/rms_norm.cuh:54-56, 134-136), with 1 / rms from row_inv_rms_kernel below.
template <bool kAlignedX, bool kTiles, int kPro = 0>
__global__ void __launch_bounds__(kQWarps * 32)
quantize_q8_1_kernel(const float* __restrict__ x, uint32_t* __restrict__ y, int64_t nblocks, uint32_t flags, const Q81Tiles tl) {
    __shared__ float tile[kQWarps][32][33];
    __shared__ uint32_t stage[kQWarps][32 * 9];

    if constexpr (kTiles) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the GEMM behind us may set itself up
        const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (gid < tl.nzero) tl.zero[gid] = 0u;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b0 = ((int64_t)blockIdx.x * kQWarps + warp) * 32;
    if (b0 >= nblocks) return;
    const int nvalid = (int)min((int64_t)32, nblocks - b0);
    const float* xb = x + b0 * 32;
    float(*tw)[33] = tile[warp];

#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int v4 = it * 32 + lane;   // float4 index inside the 1024-float group
        const int blk = v4 >> 3, j = (v4 & 7) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (blk < nvalid) {
            if constexpr (kAlignedX) {
                v = __ldcs(reinterpret_cast<const float4*>(xb) + v4);
            } else {
                v.x = xb[v4 * 4 + 0]; v.y = xb[v4 * 4 + 1]; v.z = xb[v4 * 4 + 2]; v.w = xb[v4 * 4 + 3];
            }
            if constexpr (kPro == 1) {
                const float* gb = tl.gate + b0 * 32;
                float4 g;
                if constexpr (kAlignedX) {
                    g = __ldcs(reinterpret_cast<const float4*>(gb) + v4);
                } else {
                    g.x = gb[v4 * 4 + 0]; g.y = gb[v4 * 4 + 1]; g.z = gb[v4 * 4 + 2]; g.w = gb[v4 * 4 + 3];
                }
                v.x = __fmul_rn(__fdiv_rn(v.x, __fadd_rn(1.0f, expf(-v.x))), g.x);
                v.y = __fmul_rn(__fdiv_rn(v.y, __fadd_rn(1.0f, expf(-v.y))), g.y);
                v.z = __fmul_rn(__fdiv_rn(v.z, __fadd_rn(1.0f, expf(-v.z))), g.z);
                v.w = __fmul_rn(__fdiv_rn(v.w, __fadd_rn(1.0f, expf(-v.w))), g.w);
            }
            if constexpr (kPro == 2) {
                const int64_t g = b0 + blk;                       // block index: row g / nb, columns (g % nb) * 32 + j ..
                const int64_t r = g / tl.pro_nb;
                const int col = (int)(g - r * tl.pro_nb) * 32 + j;
                const float ir = __ldg(tl.inv_rms + r);
                const float4 w = __ldg(reinterpret_cast<const float4*>(tl.gate + col));   // K % 32 == 0, weight 16-byte aligned
                v.x = __fmul_rn(__fmul_rn(v.x, ir), w.x);
                v.y = __fmul_rn(__fmul_rn(v.y, ir), w.y);
                v.z = __fmul_rn(__fmul_rn(v.z, ir), w.z);
                v.w = __fmul_rn(__fmul_rn(v.w, ir), w.w);
            }
        }
        tw[blk][j + 0] = v.x; tw[blk][j + 1] = v.y; tw[blk][j + 2] = v.z; tw[blk][j + 3] = v.w;
    }
    __syncwarp();

    float amax = 0.0f, sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const float v = tw[lane][j];
        amax = fmaxf(amax, fabsf(v));
        sum = __fadd_rn(sum, v);
    }
    const float d = __fdiv_rn(amax, 127.0f);
    const float id = (d > 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
    const bool even = flags & QGEMM_Q81_ROUND_EVEN;
    const int lo = (flags & QGEMM_Q81_CLAMP127) ? -127 : -128;

    uint32_t* sw = stage[warp] + lane * 9;
    int sum_q = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        uint32_t packed = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float v = __fmul_rn(tw[lane][w * 4 + c], id);
            int q = even ? __float2int_rn(v) : round_half_away(v);
            q = max(lo, min(127, q));
            sum_q += q;
            packed |= (uint32_t)(q & 0xff) << (8 * c);
        }
        sw[1 + w] = packed;
    }
    const float s = (flags & QGEMM_Q81_S_FROM_QSUM) ? __fmul_rn(__int2float_rn(sum_q), d) : sum;
    if constexpr (kTiles) {
        // lane = block g = t * nb + b; rows past T (tile padding) are quantized zeros: d = s = 0, q = 0
        const int64_t g = b0 + lane;
        if (lane < nvalid) {
            const int t = (int)(g / tl.nb), b = (int)(g - (int64_t)t * tl.nb);
            const int kc = b >> 2, c = (b & 3) * 2;
            uint8_t* row = tl.a8 + ((size_t)kc * tl.Tpad + t) * 128;
            *reinterpret_cast<uint4*>(row + ((c ^ (t & 7)) << 4)) = make_uint4(sw[1], sw[2], sw[3], sw[4]);
            *reinterpret_cast<uint4*>(row + (((c + 1) ^ (t & 7)) << 4)) = make_uint4(sw[5], sw[6], sw[7], sw[8]);
            const float dh = __half2float(__float2half_rn(d)), sh = __half2float(__float2half_rn(s));
            tl.as[((size_t)(t >> 7) * tl.nb + b) * 128 + (t & 127)] = make_float2(dh, __fmul_rn(tl.coef, sh));
        }
        return;
    }
    sw[0] = (uint32_t)__half_as_ushort(__float2half_rn(d)) | ((uint32_t)__half_as_ushort(__float2half_rn(s)) << 16);
    __syncwarp();

    uint32_t* yb = y + b0 * 9;
    const uint32_t* st = stage[warp];
    const int nwords = nvalid * 9;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int wi = i * 32 + lane;
        if (wi < nwords) __stcs(yb + wi, st[wi]);
    }
}

// ---------------------------------------------------------------------------
// The remaining quantize_q8_1 flavours of the reference (SURVEY 8 row A2'), none of them on the hot path: one warp per
// block, lane = element.
//   QGEMM_Q81_TREE_SUM       s = pairwise tree sum (i, i+16), (i, i+8), (i, i+4), (i, i+2), (0, 1): the shared-memory
//                            reduction of quantize_fp16_to_q8_1_smem (kernels/gemm/gemm_fused.cuh:96-127)
//   QGEMM_Q81_ID_FROM_HALF_D 1/d from the fp16-rounded d, and q narrowed to int8 before the clamp (gemm_fused.cuh:131-140)
//   QGEMM_Q81_ZERO_D1        an all-zero block stores d = 1.0 (schemas/definitions/quantization/quantize_q8_1.json)
// kHalfIn: the input is fp16 (the reference's fused kernel reads half activations).
// ---------------------------------------------------------------------------
template <bool kHalfIn>
__global__ void __launch_bounds__(256) quantize_q8_1_lanes_kernel(const void* __restrict__ xin, uint8_t* __restrict__ y, int64_t nblocks,
                                                                  uint32_t flags) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    float v;
    if constexpr (kHalfIn) v = __half2float(reinterpret_cast<const __half*>(xin)[b * 32 + lane]);
    else v = reinterpret_cast<const float*>(xin)[b * 32 + lane];
    float amax = fabsf(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    float sum;
    if (flags & QGEMM_Q81_TREE_SUM) {
        sum = v;   // lane i < w adds lane i + w: the reference's tree, level by level
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum = __fadd_rn(sum, __shfl_down_sync(0xffffffffu, sum, o));
        sum = __shfl_sync(0xffffffffu, sum, 0);
    } else {
        sum = 0.0f;   // element order, like quantize_row_q8_1_ref
#pragma unroll
        for (int j = 0; j < 32; j++) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, v, j));
    }
    float d = __fdiv_rn(amax, 127.0f);
    float id = (d > 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
    if (flags & QGEMM_Q81_ID_FROM_HALF_D) {
        const float dh = __half2float(__float2half_rn(d));
        id = (dh != 0.0f) ? __fdiv_rn(1.0f, dh) : 0.0f;
    }
    if ((flags & QGEMM_Q81_ZERO_D1) && amax == 0.0f) d = 1.0f;
    const float sv = __fmul_rn(v, id);
    int q = (flags & QGEMM_Q81_ROUND_EVEN) ? __float2int_rn(sv) : round_half_away(sv);
    if (flags & QGEMM_Q81_ID_FROM_HALF_D) q = (int)(int8_t)(q & 0xff);   // the reference's `(int8_t)roundf(..)` before its clamp
    q = max((flags & QGEMM_Q81_CLAMP127) ? -127 : -128, min(127, q));
    int sum_q = q;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum_q += __shfl_xor_sync(0xffffffffu, sum_q, o);
    uint8_t* dst = y + b * 36;
    dst[4 + lane] = (uint8_t)(q & 0xff);
    if (lane == 0) {
        const float s = (flags & QGEMM_Q81_S_FROM_QSUM) ? __fmul_rn(__int2float_rn(sum_q), d) : sum;
        *reinterpret_cast<uint32_t*>(dst) =
            (uint32_t)__half_as_ushort(__float2half_rn(d)) | ((uint32_t)__half_as_ushort(__float2half_rn(s)) << 16);
    }
}

constexpr uint32_t kQ81LaneFlags = QGEMM_Q81_TREE_SUM | QGEMM_Q81_ID_FROM_HALF_D | QGEMM_Q81_ZERO_D1;

cudaError_t launch_quantize_q8_1_f16(const void* x_f16, void* y, int64_t nblocks, uint32_t flags, cudaStream_t st) {
    if (nblocks == 0) return cudaSuccess;
    quantize_q8_1_lanes_kernel<true><<<(unsigned)((nblocks + 7) / 8), 256, 0, st>>>(x_f16, (uint8_t*)y, nblocks, flags);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_quantize_q8_1(const float* x, void* y, int64_t nblocks, uint32_t flags, cudaStream_t st) {
    if (nblocks == 0) return cudaSuccess;
    if (flags & kQ81LaneFlags) {   // the flavours of SURVEY row A2': not hot, own kernel
        quantize_q8_1_lanes_kernel<false><<<(unsigned)((nblocks + 7) / 8), 256, 0, st>>>(x, (uint8_t*)y, nblocks, flags);
        note_launch();
        return cudaGetLastError();
    }
    const int64_t per_cta = (int64_t)kQWarps * 32;
    const unsigned grid = (unsigned)((nblocks + per_cta - 1) / per_cta);
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0)
        quantize_q8_1_kernel<true, false><<<grid, kQWarps * 32, 0, st>>>(x, (uint32_t*)y, nblocks, flags, Q81Tiles{});
    else
        quantize_q8_1_kernel<false, false><<<grid, kQWarps * 32, 0, st>>>(x, (uint32_t*)y, nblocks, flags, Q81Tiles{});
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_quantize_q8_1_silu_mul(const float* x, const float* gate, void* y, int64_t nblocks, uint32_t flags, cudaStream_t st) {
    if (nblocks == 0) return cudaSuccess;
    const int64_t per_cta = (int64_t)kQWarps * 32;
    const unsigned grid = (unsigned)((nblocks + per_cta - 1) / per_cta);
    Q81Tiles tl{};
    tl.gate = gate;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gate)) & 15) == 0)
        quantize_q8_1_kernel<true, false, 1><<<grid, kQWarps * 32, 0, st>>>(x, (uint32_t*)y, nblocks, flags, tl);
    else
        quantize_q8_1_kernel<false, false, 1><<<grid, kQWarps * 32, 0, st>>>(x, (uint32_t*)y, nblocks, flags, tl);
    note_launch();
    return cudaGetLastError();
}

// 1 / rms of every row: inv_rms[r] = 1 / sqrtf((float)(sum_k x^2 / K) + eps), the sum of squares in double like
// rms_norm_cpu_f32 (kernels/normalization/rms_norm.cuh:43-51).  One CTA per row, partial sums combined in a fixed order.
__global__ void __launch_bounds__(256) row_inv_rms_kernel(const float* __restrict__ x, float* __restrict__ inv_rms, int K, float eps) {
    __shared__ double part[256];
    const float* xr = x + (size_t)blockIdx.x * K;
    double acc = 0.0;
    for (int i = threadIdx.x; i < K; i += 256) {
        const double v = (double)xr[i];
        acc += v * v;
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float rms = sqrtf(__fadd_rn((float)(part[0] / (double)K), eps));
        inv_rms[blockIdx.x] = __fdiv_rn(1.0f, rms);
    }
}

// y = quantize_q8_1(x * inv_rms[row] * weight): two launches (row statistics, then the quantizer reads x again, from L2
// for the sizes a decode or prefill step has).  inv_rms: `rows` floats of scratch.
cudaError_t launch_quantize_q8_1_rms_norm(const float* x, const float* weight, float* inv_rms, void* y, int64_t rows, int K, float eps,
                                          uint32_t flags, cudaStream_t st) {
    if (rows == 0 || K == 0) return cudaSuccess;
    row_inv_rms_kernel<<<(unsigned)rows, 256, 0, st>>>(x, inv_rms, K, eps);
    note_launch();
    if (cudaError_t e = cudaGetLastError()) return e;
    const int64_t nblocks = rows * (K / 32);
    const int64_t per_cta = (int64_t)kQWarps * 32;
    const unsigned grid = (unsigned)((nblocks + per_cta - 1) / per_cta);
    Q81Tiles tl{};
    tl.gate = weight;
    tl.inv_rms = inv_rms;
    tl.pro_nb = K / 32;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0)
        quantize_q8_1_kernel<true, false, 2><<<grid, kQWarps * 32, 0, st>>>(x, (uint32_t*)y, nblocks, flags, tl);
    else
        quantize_q8_1_kernel<false, false, 2><<<grid, kQWarps * 32, 0, st>>>(x, (uint32_t*)y, nblocks, flags, tl);
    note_launch();
    return cudaGetLastError();
}

// fp32 x[T][K] -> operand tiles + slabs for the tensor-core kernels (K % 128 == 0).  The source has T rows; the tiles
// have Tpad: the kernel runs over Tpad * nb blocks and reads zeros past row T.
__global__ void zero_pad_rows_kernel(uint8_t* a8, float2* as, int T, int Tpad, int nb) {
    // rows [T, Tpad) of every K chunk and slab: quantized zeros
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int pad = Tpad - T;
    if (i >= (int64_t)pad * nb) return;
    const int t = T + (int)(i / nb), b = (int)(i % nb);
    const int kc = b >> 2, c = (b & 3) * 2;
    uint8_t* row = a8 + ((size_t)kc * Tpad + t) * 128;
    *reinterpret_cast<uint4*>(row + ((c ^ (t & 7)) << 4)) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(row + (((c + 1) ^ (t & 7)) << 4)) = make_uint4(0, 0, 0, 0);
    as[((size_t)(t >> 7) * nb + b) * 128 + (t & 127)] = make_float2(0.f, 0.f);
}

cudaError_t launch_quantize_q8_1_tiles(const float* x, uint8_t* a8, float2* as, int T, int Tpad, int K, float coef, uint32_t flags,
                                       cudaStream_t st, unsigned* zero, int nzero, const float* gate) {
    const int nb = K / 32;
    if (T < 1 || nb < 4 || (nb & 3)) return cudaErrorInvalidValue;
    if (Tpad > T) {   // at most 127 rows: a sliver in front of the main pass (same stream, so ordered before the GEMM)
        const int64_t n = (int64_t)(Tpad - T) * nb;
        zero_pad_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a8, as, T, Tpad, nb);
        note_launch();
    }
    const int64_t nblocks = (int64_t)T * nb;
    const int64_t per_cta = (int64_t)kQWarps * 32;
    const unsigned grid = (unsigned)((nblocks + per_cta - 1) / per_cta);
    const Q81Tiles tl{a8, as, T, Tpad, nb, coef, zero, nzero, gate, nullptr, 0};
    const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gate)) & 15) == 0;
    if (gate) {   // silu(x) * gate is what gets quantized (the FFN down projection's input)
        if (al) quantize_q8_1_kernel<true, true, 1><<<grid, kQWarps * 32, 0, st>>>(x, nullptr, nblocks, flags, tl);
        else quantize_q8_1_kernel<false, true, 1><<<grid, kQWarps * 32, 0, st>>>(x, nullptr, nblocks, flags, tl);
    } else if (al) {
        quantize_q8_1_kernel<true, true><<<grid, kQWarps * 32, 0, st>>>(x, nullptr, nblocks, flags, tl);
    } else {
        quantize_q8_1_kernel<false, true><<<grid, kQWarps * 32, 0, st>>>(x, nullptr, nblocks, flags, tl);
    }
    note_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Weight quantizers: test-data producers, one thread per block (not hot).
//   q4_0: include/quantize.h:35-70 / test_framework.cuh:162-192
//   q8_0: include/quantize.h:111-135 (clamp -128..127)
//   q4_1/q5_0/q5_1: test_framework.cuh:256-367
// ---------------------------------------------------------------------------
__device__ __forceinline__ void st_u16(uint8_t* p, uint32_t v) { *reinterpret_cast<uint16_t*>(p) = (uint16_t)v; }
__device__ __forceinline__ uint32_t f2h_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }

template <int WT>
__global__ void quantize_weight_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, int64_t nblocks,
                                       uint32_t flags) {
    using F = Fmt<WT>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    const float* src = x + i * 32;
    uint8_t* dst = y + i * F::bytes;
    const bool even = flags & QGEMM_Q81_ROUND_EVEN;
    auto rnd = [&](float v) { return even ? __float2int_rn(v) : round_half_away(v); };

    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = src[j];

    if constexpr (WT == QGEMM_TYPE_Q8_0) {
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j++) amax = fmaxf(amax, fabsf(v[j]));
        const float d = __fdiv_rn(amax, 127.0f);
        const float id = (d > 0.f) ? __fdiv_rn(1.0f, d) : 0.f;
        st_u16(dst, f2h_bits(d));
#pragma unroll
        for (int j = 0; j < 32; j++) dst[2 + j] = (uint8_t)(int8_t)max(-128, min(127, rnd(__fmul_rn(v[j], id))));
    } else if constexpr (F::m < 0) {  // symmetric q4_0 / q5_0
        constexpr int half_range = (F::bits == 4) ? 8 : 16;
        constexpr float div = (F::bits == 4) ? 7.0f : 15.0f;
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j++) amax = fmaxf(amax, fabsf(v[j]));
        const float d = __fdiv_rn(amax, div);
        const float id = (d > 0.f) ? __fdiv_rn(1.0f, d) : 0.f;
        st_u16(dst, f2h_bits(d));
        uint32_t qh = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            int q0 = rnd(__fmul_rn(v[j], id)) + half_range;
            int q1 = rnd(__fmul_rn(v[j + 16], id)) + half_range;
            q0 = max(0, min(2 * half_range - 1, q0));
            q1 = max(0, min(2 * half_range - 1, q1));
            dst[F::qs + j] = (uint8_t)(((q1 & 0xf) << 4) | (q0 & 0xf));
            qh |= (uint32_t)((q0 >> 4) & 1) << j;
            qh |= (uint32_t)((q1 >> 4) & 1) << (j + 16);
        }
        if constexpr (F::bits == 5) { st_u16(dst + F::qh, qh & 0xffff); st_u16(dst + F::qh + 2, qh >> 16); }
    } else {  // asymmetric q4_1 / q5_1
        constexpr int qmax = (F::bits == 4) ? 15 : 31;
        float mn = v[0], mx = v[0];
#pragma unroll
        for (int j = 1; j < 32; j++) { mn = fminf(mn, v[j]); mx = fmaxf(mx, v[j]); }
        const float d = __fdiv_rn(__fsub_rn(mx, mn), (float)qmax);
        const float id = (d > 0.f) ? __fdiv_rn(1.0f, d) : 0.f;
        st_u16(dst, f2h_bits(d));
        st_u16(dst + F::m, f2h_bits(mn));
        uint32_t qh = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            int q0 = rnd(__fmul_rn(__fsub_rn(v[j], mn), id));
            int q1 = rnd(__fmul_rn(__fsub_rn(v[j + 16], mn), id));
            q0 = max(0, min(qmax, q0));
            q1 = max(0, min(qmax, q1));
            dst[F::qs + j] = (uint8_t)(((q1 & 0xf) << 4) | (q0 & 0xf));
            qh |= (uint32_t)((q0 >> 4) & 1) << j;
            qh |= (uint32_t)((q1 >> 4) & 1) << (j + 16);
        }
        if constexpr (F::bits == 5) { st_u16(dst + F::qh, qh & 0xffff); st_u16(dst + F::qh + 2, qh >> 16); }
    }
}

cudaError_t launch_quantize_weight(int wtype, const float* x, void* y, int64_t nblocks, uint32_t flags,
                                   cudaStream_t st) {
    if (nblocks == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((nblocks + 127) / 128);
    uint8_t* yb = (uint8_t*)y;
    switch (wtype) {
    case QGEMM_TYPE_Q4_0: quantize_weight_kernel<QGEMM_TYPE_Q4_0><<<grid, 128, 0, st>>>(x, yb, nblocks, flags); break;
    case QGEMM_TYPE_Q4_1: quantize_weight_kernel<QGEMM_TYPE_Q4_1><<<grid, 128, 0, st>>>(x, yb, nblocks, flags); break;
    case QGEMM_TYPE_Q5_0: quantize_weight_kernel<QGEMM_TYPE_Q5_0><<<grid, 128, 0, st>>>(x, yb, nblocks, flags); break;
    case QGEMM_TYPE_Q5_1: quantize_weight_kernel<QGEMM_TYPE_Q5_1><<<grid, 128, 0, st>>>(x, yb, nblocks, flags); break;
    case QGEMM_TYPE_Q8_0: quantize_weight_kernel<QGEMM_TYPE_Q8_0><<<grid, 128, 0, st>>>(x, yb, nblocks, flags); break;
    default: return cudaErrorInvalidValue;
    }
    note_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Dequantize: 8 threads per block, 4 consecutive elements each, so a warp
// writes 512 contiguous bytes.  include/quantize.h:84-102,140-153,198-211.
// ---------------------------------------------------------------------------
template <int TYPE>
__global__ void dequantize_kernel(const uint8_t* __restrict__ x, float* __restrict__ y, int64_t nblocks) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = gid >> 3;
    const int part = (int)(gid & 7);  // elements 4*part .. 4*part+3
    if (b >= nblocks) return;
    float out[4];
    if constexpr (TYPE == QGEMM_TYPE_Q8_1 || TYPE == QGEMM_TYPE_Q8_0) {
        constexpr int bytes = (TYPE == QGEMM_TYPE_Q8_1) ? 36 : 34;
        constexpr int qs = (TYPE == QGEMM_TYPE_Q8_1) ? 4 : 2;
        const uint8_t* blk = x + b * bytes;
        const float d = ld_half(blk);
#pragma unroll
        for (int c = 0; c < 4; c++) out[c] = __fmul_rn((float)(int8_t)blk[qs + part * 4 + c], d);
    } else {
        using F = Fmt<TYPE>;
        const uint8_t* blk = x + b * F::bytes;
        const float d = ld_half(blk);
        float m = 0.f;
        if constexpr (F::m >= 0) m = ld_half(blk + F::m);
        uint32_t qh = 0;
        if constexpr (F::bits == 5) qh = ld_u32_a2(blk + F::qh);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int e = part * 4 + c;
            const int byte = blk[F::qs + (e & 15)];
            int q = (e < 16) ? (byte & 0xf) : (byte >> 4);
            if constexpr (F::bits == 5) q |= ((qh >> e) & 1) << 4;
            if constexpr (F::m >= 0) out[c] = __fadd_rn(__fmul_rn((float)q, d), m);
            else out[c] = __fmul_rn((float)(q - (F::bits == 4 ? 8 : 16)), d);
        }
    }
    reinterpret_cast<float4*>(y)[gid] = make_float4(out[0], out[1], out[2], out[3]);
}

cudaError_t launch_dequantize(int type, const void* x, float* y, int64_t nblocks, cudaStream_t st) {
    if (nblocks == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((nblocks * 8 + 255) / 256);
    const uint8_t* xb = (const uint8_t*)x;
    switch (type) {
    case QGEMM_TYPE_Q4_0: dequantize_kernel<QGEMM_TYPE_Q4_0><<<grid, 256, 0, st>>>(xb, y, nblocks); break;
    case QGEMM_TYPE_Q4_1: dequantize_kernel<QGEMM_TYPE_Q4_1><<<grid, 256, 0, st>>>(xb, y, nblocks); break;
    case QGEMM_TYPE_Q5_0: dequantize_kernel<QGEMM_TYPE_Q5_0><<<grid, 256, 0, st>>>(xb, y, nblocks); break;
    case QGEMM_TYPE_Q5_1: dequantize_kernel<QGEMM_TYPE_Q5_1><<<grid, 256, 0, st>>>(xb, y, nblocks); break;
    case QGEMM_TYPE_Q8_0: dequantize_kernel<QGEMM_TYPE_Q8_0><<<grid, 256, 0, st>>>(xb, y, nblocks); break;
    case QGEMM_TYPE_Q8_1: dequantize_kernel<QGEMM_TYPE_Q8_1><<<grid, 256, 0, st>>>(xb, y, nblocks); break;
    default: return cudaErrorInvalidValue;
    }
    note_launch();
    return cudaGetLastError();
}

}  // namespace qgemm
