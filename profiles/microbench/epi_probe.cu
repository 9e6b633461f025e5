// epi_probe.cu -- round-2 microbenchmarks behind the prefill epilogue design.
//
//  (1) fold_rate: the per-block epilogue of the tcgen05 kernel in isolation -- 16 (or 8) warps per CTA read a
//      128 x 128 s32 tile out of TMEM and fold it into fp32 register accumulators, no MMA, no operand traffic --
//      for the candidate instruction sequences.  Reports cycles per quantization block per SM; the int8 MMA needs
//      64 of them, so 64 / cycles is the fraction of the tensor peak that epilogue can sustain.
//  (2) mma_probe: functional check of the operand layouts the new kernel relies on: u8 x s8 + (s8 const) x s8
//      accumulated into one s32 tile (the offset hoist), and kind::tf32 with the no-swizzle K-major layout
//      (the scale-term GEMM), against host arithmetic.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo epi_probe.cu -o epi_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#define LD16_ARGS(v, o) "=r"(v[o+0]), "=r"(v[o+1]), "=r"(v[o+2]), "=r"(v[o+3]), "=r"(v[o+4]), "=r"(v[o+5]), "=r"(v[o+6]), "=r"(v[o+7]), \
                        "=r"(v[o+8]), "=r"(v[o+9]), "=r"(v[o+10]), "=r"(v[o+11]), "=r"(v[o+12]), "=r"(v[o+13]), "=r"(v[o+14]), "=r"(v[o+15])
template <int O>
__device__ __forceinline__ void ld16(uint32_t taddr, int (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : LD16_ARGS(v, O) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : LD16_ARGS(v, 0), LD16_ARGS(v, 16) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const int (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                 "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

__device__ __forceinline__ uint64_t pk(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int kMagic = 0x4B400000;           // bits of 12582912.0f = 1.5 * 2^23
constexpr float kMagicF = 12582912.0f;

// ---------------------------------------------------------------------------------------------------
// (1) fold-rate kernel.  MODE:
//   0  round-1 sequence: x32 load, wait, I2FP, ffma2(da, f, ca) [da per thread], ffma2(dw_i, t, acc)
//   1  magic IADD instead of I2FP, first FMA carries -d*magic; x32 load, no overlap
//   2  as 1, two x16 half loads, the next half in flight while one half is folded
//   3  as 2 without any conversion (what a float-producing MMA would leave): lower bound
//   4  as 2 with scalar FFMA instead of FFMA2
//   5  as 2 with I2FP (separates "overlap" from "conversion")
// The per-thread scale (d of the lane's row) is re-read from shared memory every block, the 32 per-column scales
// come as 8 broadcast LDS.128 -- the traffic of the real kernel.
// ---------------------------------------------------------------------------------------------------
template <int MODE, int COLS>
__device__ __forceinline__ void fold_half(uint64_t (&acc)[16], const int (&x)[32], int off, int h, float dl, float nb,
                                          const float4* dcol4) {
    // 16 columns [off, off+16) of this thread's 32
#pragma unroll
    for (int i4 = 0; i4 < 4; i4++) {
        const float4 dc = dcol4[(off >> 2) + i4];   // per-column scales, broadcast
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int c = off + 4 * i4 + 2 * j;
            const float d0 = j ? dc.z : dc.x, d1 = j ? dc.w : dc.y;
            uint64_t& a = acc[(h * 16 + 4 * i4 + 2 * j) >> 1];
            if constexpr (MODE == 0 || MODE == 5) {
                const uint64_t f = pk(__int2float_rn(x[c]), __int2float_rn(x[c + 1]));
                a = ffma2(pk(d0, d1), ffma2(pk(dl, dl), f, pk(nb, nb)), a);
            } else if constexpr (MODE == 3) {
                const uint64_t f = pk(__int_as_float(x[c]), __int_as_float(x[c + 1]));
                a = ffma2(pk(d0, d1), ffma2(pk(dl, dl), f, pk(nb, nb)), a);
            } else if constexpr (MODE == 4) {
                float lo, hi;
                unpk(a, lo, hi);
                const float u0 = __fmaf_rn(dl, __int_as_float(x[c] + kMagic), nb);
                const float u1 = __fmaf_rn(dl, __int_as_float(x[c + 1] + kMagic), nb);
                a = pk(__fmaf_rn(d0, u0, lo), __fmaf_rn(d1, u1, hi));
            } else {
                const uint64_t f = pk(__int_as_float(x[c] + kMagic), __int_as_float(x[c + 1] + kMagic));
                a = ffma2(pk(d0, d1), ffma2(pk(dl, dl), f, pk(nb, nb)), a);
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(640, 1) fold_rate(float* out, long long* cyc, int nblk) {
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_lane[8][128];   // per-row scale of 8 blocks
    __shared__ __align__(16) float s_col[8][128];    // per-column scale of 8 blocks
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * 128; i += blockDim.x) {
        (&s_lane[0][0])[i] = 1.0f + (i % 7) * 0.125f;
        (&s_col[0][0])[i] = 0.5f + (i % 5) * 0.0625f;
    }
    if (warp == 2) tmem_alloc(&tmem_slot, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp >= 4) {
        const int ew = warp - 4, quarter = warp & 3, cgrp = ew >> 2;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        // small integers into every buffer so the arithmetic stays finite
        int v[16];
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = (lane * 3 + i * 7 + warp) % 2001 - 1000;
        for (int b = 0; b < 4; b++) {
            st16(tmem + lane_addr + b * 128 + cgrp * 32, v);
            st16(tmem + lane_addr + b * 128 + cgrp * 32 + 16, v);
        }
        wait_st();
        fence_before();
        asm volatile("bar.sync 1, 512;" ::: "memory");
        fence_after();
        uint64_t acc[16];
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = 0ull;
        const int row = quarter * 32 + lane;
        int x[32];
        const long long t0 = clock64();
        if constexpr (MODE == 7 || MODE == 8 || MODE == 9) {
#pragma unroll 1
            for (int b = 0; b < nblk; b++) {
                const int buf = b & 3, j = b & 7;
                if (MODE != 9) asm volatile("bar.sync 1, 512;" ::: "memory");   // all 16 warps start the block together
                ld32(tmem + lane_addr + buf * 128 + cgrp * 32, x);
                wait_ld();
                const float dl = s_lane[j][row];
                const float nb = dl * -8.0f;
                const float4* dc = reinterpret_cast<const float4*>(&s_col[j][cgrp * 32]);
                if constexpr (MODE == 7) {
                    fold_half<0, 32>(acc, x, 0, 0, dl, nb, dc);
                    fold_half<0, 32>(acc, x, 16, 1, dl, nb, dc);
                } else {
                    // conversions and FMAs alternate, FMAs two pairs behind, order pinned with asm volatile
                    float4 d4[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) d4[i] = dc[i];
                    auto fma_pair = [&](int q) {
                        const float4 d = d4[q >> 1];
                        const uint64_t dd = (q & 1) ? pk(d.z, d.w) : pk(d.x, d.y);
                        uint64_t t;
                        const uint64_t f = pk(__int_as_float(x[2 * q]), __int_as_float(x[2 * q + 1]));
                        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(pk(dl, dl)), "l"(f), "l"(pk(nb, nb)));
                        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[q]) : "l"(dd), "l"(t));
                    };
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(x[2 * i]));
                        asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(x[2 * i + 1]));
                        if (i >= 2) fma_pair(i - 2);
                    }
                    fma_pair(14);
                    fma_pair(15);
                }
            }
        } else if constexpr (MODE == 0 || MODE == 1) {
#pragma unroll 1
            for (int b = 0; b < nblk; b++) {
                const int buf = b & 3, j = b & 7;
                ld32(tmem + lane_addr + buf * 128 + cgrp * 32, x);
                wait_ld();
                const float dl = s_lane[j][row];
                const float nb = (MODE == 0) ? dl * -8.0f : dl * -kMagicF;
                const float4* dc = reinterpret_cast<const float4*>(&s_col[j][cgrp * 32]);
                fold_half<MODE, 32>(acc, x, 0, 0, dl, nb, dc);
                fold_half<MODE, 32>(acc, x, 16, 1, dl, nb, dc);
            }
        } else if constexpr (MODE == 6) {   // MODE 2 arithmetic, block loop unrolled by 4: buffer and slab addresses are immediates
            ld16<0>(tmem + lane_addr + cgrp * 32, x);
            wait_ld();
            const uint32_t tb = tmem + lane_addr + cgrp * 32;
#pragma unroll 1
            for (int b4 = 0; b4 < nblk; b4 += 4) {
                const int jb = b4 & 4;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    ld16<16>(tb + q * 128 + 16, x);
                    const float dl = s_lane[jb + q][row];
                    const float nb = dl * -kMagicF;
                    const float4* dc = reinterpret_cast<const float4*>(&s_col[jb + q][cgrp * 32]);
                    fold_half<2, 32>(acc, x, 0, 0, dl, nb, dc);
                    wait_ld();
                    ld16<0>(tb + ((q + 1) & 3) * 128, x);
                    fold_half<2, 32>(acc, x, 16, 1, dl, nb, dc);
                    wait_ld();
                }
            }
        } else {
            ld16<0>(tmem + lane_addr + cgrp * 32, x);
            wait_ld();
#pragma unroll 1
            for (int b = 0; b < nblk; b++) {
                const int buf = b & 3, j = b & 7, nbuf = (b + 1) & 3;
                ld16<16>(tmem + lane_addr + buf * 128 + cgrp * 32 + 16, x);   // second half of this block in flight
                const float dl = s_lane[j][row];
                const float nb = (MODE == 5) ? dl * -8.0f : dl * -kMagicF;
                const float4* dc = reinterpret_cast<const float4*>(&s_col[j][cgrp * 32]);
                fold_half<MODE, 32>(acc, x, 0, 0, dl, nb, dc);
                wait_ld();
                ld16<0>(tmem + lane_addr + nbuf * 128 + cgrp * 32, x);        // first half of the next block in flight
                fold_half<MODE, 32>(acc, x, 16, 1, dl, nb, dc);
                wait_ld();
            }
        }
        const long long t1 = clock64();
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++) { float lo, hi; unpk(acc[i], lo, hi); s += lo + hi; }
        out[(size_t)blockIdx.x * 512 + (threadIdx.x - 128)] = s;
        if (threadIdx.x == 128) cyc[blockIdx.x] = t1 - t0;
    }
    fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
}

// 8 epilogue warps x 64 columns, halves of 16 columns pipelined (MODE 2 arithmetic)
__global__ void __launch_bounds__(384, 1) fold_rate_8w(float* out, long long* cyc, int nblk) {
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_lane[8][128];
    __shared__ __align__(16) float s_col[8][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * 128; i += blockDim.x) {
        (&s_lane[0][0])[i] = 1.0f + (i % 7) * 0.125f;
        (&s_col[0][0])[i] = 0.5f + (i % 5) * 0.0625f;
    }
    if (warp == 2) tmem_alloc(&tmem_slot, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp >= 4) {
        const int ew = warp - 4, quarter = warp & 3, cgrp = ew >> 2;   // cgrp 0..1, 64 columns each
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        int v[16];
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = (lane * 3 + i * 7 + warp) % 2001 - 1000;
        for (int b = 0; b < 4; b++)
            for (int q = 0; q < 4; q++) st16(tmem + lane_addr + b * 128 + cgrp * 64 + q * 16, v);
        wait_st();
        fence_before();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        fence_after();
        uint64_t acc0[16], acc1[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { acc0[i] = 0ull; acc1[i] = 0ull; }
        const int row = quarter * 32 + lane;
        int x[32];
        const long long t0 = clock64();
        ld16<0>(tmem + lane_addr + cgrp * 64, x);
        wait_ld();
#pragma unroll 1
        for (int b = 0; b < nblk; b++) {
            const int buf = b & 3, j = b & 7, nbuf = (b + 1) & 3;
            const uint32_t base = tmem + lane_addr + buf * 128 + cgrp * 64;
            const float dl = s_lane[j][row];
            const float nb = dl * -kMagicF;
            const float4* dc = reinterpret_cast<const float4*>(&s_col[j][cgrp * 64]);
            ld16<16>(base + 16, x);
            fold_half<2, 64>(acc0, x, 0, 0, dl, nb, dc);
            wait_ld();
            ld16<0>(base + 32, x);
            fold_half<2, 64>(acc0, x, 16, 1, dl, nb, dc);
            wait_ld();
            ld16<16>(base + 48, x);
            fold_half<2, 64>(acc1, x, 0, 0, dl, nb, dc + 8);
            wait_ld();
            ld16<0>(tmem + lane_addr + nbuf * 128 + cgrp * 64, x);
            fold_half<2, 64>(acc1, x, 16, 1, dl, nb, dc + 8);
            wait_ld();
        }
        const long long t1 = clock64();
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++) { float lo, hi; unpk(acc0[i], lo, hi); s += lo + hi; unpk(acc1[i], lo, hi); s += lo + hi; }
        out[(size_t)blockIdx.x * 512 + (threadIdx.x - 128)] = s;
        if (threadIdx.x == 128) cyc[blockIdx.x] = t1 - t0;
    }
    fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------
// (2) functional probe of the MMA operand layouts
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(accum) : "memory");
}
// K-major, 128-byte rows, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// K-major, no swizzle: 8 x 16-byte core matrices; lbo = bytes between the two 16-byte K chunks, sbo = bytes between 8-row groups
__device__ __forceinline__ uint64_t desc_interleave(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}

// w [128][32] u8 (values 0..15), a [128][32] s8, sw [128][16] f32 (tf32-exact), sa [128][16] f32; outputs [128][128]
__global__ void __launch_bounds__(128, 1) mma_probe(const uint8_t* w, const int8_t* a, const float* sw, const float* sa, int* d_int,
                                                    float* d_f32, int offset) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint8_t* tW = smem;                    // 16 KB, sw128 (first 32 bytes of each row used)
    uint8_t* tA = smem + 16384;            // 16 KB
    uint8_t* tC = smem + 32768;            // 4 KB constant tile (every byte = -offset)
    float* tSW = reinterpret_cast<float*>(smem + 36864);   // 8 KB: [chunk 4][group 16][8 rows][4 floats]
    float* tSA = reinterpret_cast<float*>(smem + 45056);   // 8 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 53248);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 53248 + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {   // thread = row
        const int r = tid;
        for (int c = 0; c < 2; c++) {
            *reinterpret_cast<uint4*>(tW + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(w + r * 32 + c * 16);
            *reinterpret_cast<uint4*>(tA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a + r * 32 + c * 16);
        }
        for (int i = 0; i < 8; i++) reinterpret_cast<uint32_t*>(tC)[r * 8 + i] = 0x01010101u * (uint32_t)((-offset) & 0xff);
        for (int c = 0; c < 4; c++) {
            const size_t o = ((size_t)c * 16 + (r >> 3)) * 32 + (r & 7) * 4;
            *reinterpret_cast<float4*>(tSW + o) = *reinterpret_cast<const float4*>(sw + r * 16 + c * 4);
            *reinterpret_cast<float4*>(tSA + o) = *reinterpret_cast<const float4*>(sa + r * 16 + c * 4);
        }
    }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc(slot, 256);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t id_u8s8 = (2u << 4) | (0u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t id_s8s8 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t id_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        mma_i8(tmem, desc_sw128(smem_u32(tW)), desc_sw128(smem_u32(tA)), id_u8s8, 0u);
        if (offset) mma_i8(tmem, desc_interleave(smem_u32(tC), 2048, 128), desc_sw128(smem_u32(tA)), id_s8s8, 1u);
        // two K = 8 steps: chunks {0,1} then {2,3}; chunks are 2048 bytes apart, 8-row groups 128 bytes apart
        mma_tf32(tmem + 128, desc_interleave(smem_u32(tSW), 2048, 128), desc_interleave(smem_u32(tSA), 2048, 128), id_tf32, 0u);
        mma_tf32(tmem + 128, desc_interleave(smem_u32(tSW) + 4096, 2048, 128), desc_interleave(smem_u32(tSA) + 4096, 2048, 128), id_tf32, 1u);
        commit(bar);
    }
    while (!mbar_try_wait(bar, 0)) {}
    fence_after();
    int x[32];
    for (int c0 = 0; c0 < 256; c0 += 32) {
        ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, x);
        wait_ld();
        for (int i = 0; i < 32; i++) {
            if (c0 < 128) d_int[(warp * 32 + lane) * 128 + c0 + i] = x[i];
            else d_f32[(warp * 32 + lane) * 128 + c0 - 128 + i] = __int_as_float(x[i]);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// Is the fp32 accumulation of kind::f8f6f4 exact on small integers?  w8 [128][32] e4m3 (q - 8), alo / ahi [128][32] e4m3
// (a = 16 * hi + lo: lo in 0..15, 16 * hi in -128..112); out [128][128] f32 = sum_k w * (lo + 16 hi)
__global__ void __launch_bounds__(128, 1) f8_probe(const uint8_t* w8, const uint8_t* alo, const uint8_t* ahi, float* d_f32) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint8_t* tW = smem;
    uint8_t* tL = smem + 16384;
    uint8_t* tH = smem + 32768;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 49152 + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int c = 0; c < 2; c++) {
        const int r = tid;
        *reinterpret_cast<uint4*>(tW + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(w8 + r * 32 + c * 16);
        *reinterpret_cast<uint4*>(tL + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(alo + r * 32 + c * 16);
        *reinterpret_cast<uint4*>(tH + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(ahi + r * 32 + c * 16);
    }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc(slot, 128);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t id = (1u << 4) | (0u << 7) | (0u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // e4m3 x e4m3 -> f32
        mma_f8(tmem, desc_sw128(smem_u32(tW)), desc_sw128(smem_u32(tL)), id, 0u);
        mma_f8(tmem, desc_sw128(smem_u32(tW)), desc_sw128(smem_u32(tH)), id, 1u);
        commit(bar);
    }
    while (!mbar_try_wait(bar, 0)) {}
    fence_after();
    int x[32];
    for (int c0 = 0; c0 < 128; c0 += 32) {
        ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, x);
        wait_ld();
        for (int i = 0; i < 32; i++) d_f32[(warp * 32 + lane) * 128 + c0 + i] = __int_as_float(x[i]);
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// exact e4m3 encoding of an integer that is representable (|n| <= 448, few significant bits); aborts otherwise
static uint8_t e4m3_of_int(int n) {
    if (n == 0) return 0;
    const uint8_t sign = n < 0 ? 0x80 : 0;
    int m = abs(n), e = 0;
    while (m >= 16) { if (m & 1) { printf("e4m3: %d not representable\n", n); exit(1); } m >>= 1; e++; }
    while (m < 8) { m <<= 1; e--; }          // m in [8, 15] = 1.mmm * 8
    const int exp = e + 3 + 7;                // value = (m / 8) * 2^(e + 3)
    if (exp < 1 || exp > 15) { printf("e4m3: %d out of range\n", n); exit(1); }
    return sign | (uint8_t)(exp << 3) | (uint8_t)(m & 7);
}

int main(int argc, char** argv) {
    const int nblk = argc > 1 ? atoi(argv[1]) : 4096;
    int dev = 0, sms = 0, clk = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    printf("device: %d SMs, %d kHz\n", sms, clk);

    // ---- (2) functional probe first: a wrong layout should be known before anything is timed
    {
        std::vector<uint8_t> w(128 * 32);
        std::vector<int8_t> a(128 * 32);
        std::vector<float> sw(128 * 16), sa(128 * 16);
        srand(7);
        for (auto& v : w) v = rand() % 16;
        for (auto& v : a) v = (int8_t)(rand() % 256 - 128);
        for (auto& v : sw) v = (float)(rand() % 2047 - 1023) / 64.0f;    // 11 significant bits: exact in tf32
        for (auto& v : sa) v = (float)(rand() % 2047 - 1023) / 8.0f;
        uint8_t* dw; int8_t* da; float *dsw, *dsa, *df; int* di;
        CK(cudaMalloc(&dw, w.size())); CK(cudaMalloc(&da, a.size())); CK(cudaMalloc(&dsw, sw.size() * 4)); CK(cudaMalloc(&dsa, sa.size() * 4));
        CK(cudaMalloc(&di, 128 * 128 * 4)); CK(cudaMalloc(&df, 128 * 128 * 4));
        CK(cudaMemcpy(dw, w.data(), w.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dsw, sw.data(), sw.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dsa, sa.data(), sa.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024));
        for (int offset : {0, 8, 16}) {
            mma_probe<<<1, 128, 56 * 1024>>>(dw, da, dsw, dsa, di, df, offset);
            CK(cudaDeviceSynchronize());
            std::vector<int> hi(128 * 128);
            std::vector<float> hf(128 * 128);
            CK(cudaMemcpy(hi.data(), di, hi.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hf.data(), df, hf.size() * 4, cudaMemcpyDeviceToHost));
            long bad_i = 0, bad_f = 0;
            double maxrel = 0;
            for (int f = 0; f < 128; f++)
                for (int t = 0; t < 128; t++) {
                    int s = 0;
                    for (int k = 0; k < 32; k++) s += ((int)w[f * 32 + k] - offset) * (int)a[t * 32 + k];
                    if (s != hi[f * 128 + t]) bad_i++;
                    double r = 0;
                    for (int k = 0; k < 16; k++) r += (double)sw[f * 16 + k] * (double)sa[t * 16 + k];
                    const double e = fabs(r - hf[f * 128 + t]) / (fabs(r) + 1e-3);
                    if (e > 1e-6) bad_f++;
                    if (e > maxrel) maxrel = e;
                }
            printf("mma_probe offset %2d: int mismatches %ld / 16384, tf32 mismatches %ld / 16384 (max rel %.3g)  [D rows = weights, columns = tokens]\n",
                   offset, bad_i, bad_f, maxrel);
        }
    }

    // ---- (2b) fp8 accumulation exactness
    {
        uint8_t *dw, *dl, *dh; float* df;
        CK(cudaMalloc(&dw, 4096)); CK(cudaMalloc(&dl, 4096)); CK(cudaMalloc(&dh, 4096)); CK(cudaMalloc(&df, 128 * 128 * 4));
        CK(cudaFuncSetAttribute(f8_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024));
        for (int tc = 0; tc < 3; tc++) {
            std::vector<int> w(128 * 32), a(128 * 32);
            srand(11 + tc);
            for (int i = 0; i < 128 * 32; i++) {
                if (tc == 0) { w[i] = rand() % 16 - 8; a[i] = rand() % 256 - 128; }
                else if (tc == 1) { w[i] = (i / 32) % 2 ? -8 : 7; a[i] = (i / 32) % 3 ? -128 : 127; }          // |sum| at its bound
                else { w[i] = (i % 2) ? -8 : 7; a[i] = (i % 32) < 16 ? 127 - (i % 3) : -128 + (i % 5); }       // large partial sums that cancel
            }
            std::vector<uint8_t> ew(4096), el(4096), eh(4096);
            for (int i = 0; i < 4096; i++) {
                const int lo = a[i] & 15, hi16 = a[i] - lo;
                ew[i] = e4m3_of_int(w[i]); el[i] = e4m3_of_int(lo); eh[i] = e4m3_of_int(hi16);
            }
            CK(cudaMemcpy(dw, ew.data(), 4096, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dl, el.data(), 4096, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dh, eh.data(), 4096, cudaMemcpyHostToDevice));
            f8_probe<<<1, 128, 52 * 1024>>>(dw, dl, dh, df);
            CK(cudaDeviceSynchronize());
            std::vector<float> hf(128 * 128);
            CK(cudaMemcpy(hf.data(), df, hf.size() * 4, cudaMemcpyDeviceToHost));
            long bad = 0; double maxabs = 0; int maxsum = 0;
            for (int f = 0; f < 128; f++)
                for (int t = 0; t < 128; t++) {
                    int sum = 0;
                    for (int k = 0; k < 32; k++) sum += w[f * 32 + k] * a[t * 32 + k];
                    if (abs(sum) > maxsum) maxsum = abs(sum);
                    const double e = fabs((double)sum - hf[f * 128 + t]);
                    if (e != 0) bad++;
                    if (e > maxabs) maxabs = e;
                }
            printf("f8_probe case %d: inexact %ld / 16384, max |err| %.3g, max |sum| %d\n", tc, bad, maxabs, maxsum);
        }
    }

    // ---- (1) fold rates
    float* out; long long* cyc;
    CK(cudaMalloc(&out, (size_t)sms * 512 * 4)); CK(cudaMalloc(&cyc, sms * 8));
    std::vector<long long> h(sms);
    auto report = [&](const char* name) {
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost));
        double sum = 0; long long mx = 0;
        for (auto v : h) { sum += v; if (v > mx) mx = v; }
        const double per = sum / sms / nblk;
        printf("%-58s %7.1f cycles/block (max CTA %7.1f) -> %5.1f %% of the int8 MMA rate\n", name, per, (double)mx / nblk, 6400.0 / per);
    };
    for (int rep = 0; rep < 2; rep++) {
        fold_rate<0><<<sms, 640>>>(out, cyc, nblk); if (rep) report("0: x32, I2FP, 2 FFMA2 (round 1)");
        fold_rate<1><<<sms, 640>>>(out, cyc, nblk); if (rep) report("1: x32, magic IADD, 2 FFMA2");
        fold_rate<2><<<sms, 640>>>(out, cyc, nblk); if (rep) report("2: 2 x x16 pipelined, magic IADD, 2 FFMA2");
        fold_rate<3><<<sms, 640>>>(out, cyc, nblk); if (rep) report("3: 2 x x16 pipelined, no conversion, 2 FFMA2");
        fold_rate<4><<<sms, 640>>>(out, cyc, nblk); if (rep) report("4: 2 x x16 pipelined, magic IADD, 2 scalar FFMA");
        fold_rate<5><<<sms, 640>>>(out, cyc, nblk); if (rep) report("5: 2 x x16 pipelined, I2FP, 2 FFMA2");
        fold_rate<6><<<sms, 640>>>(out, cyc, nblk); if (rep) report("6: as 2, block loop unrolled by 4");
        fold_rate<7><<<sms, 640>>>(out, cyc, nblk); if (rep) report("7: as 0, all 16 warps barrier-aligned at every block");
        fold_rate<8><<<sms, 640>>>(out, cyc, nblk); if (rep) report("8: as 7, conversions and FMAs interleaved (pinned order)");
        fold_rate<9><<<sms, 640>>>(out, cyc, nblk); if (rep) report("9: as 8 without the barrier");
        fold_rate_8w<<<sms, 384>>>(out, cyc, nblk); if (rep) report("8 warps x 64 columns, x16 pipelined, magic IADD, 2 FFMA2");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
