// tc05.cuh -- tcgen05 / TMEM wrappers, packed fp32x2 arithmetic and the per-block scale fold on register pairs,
// shared by the two prefill kernels (mmq.cu: operand tiles from a prepass; mmq_native.cu: native blocks unpacked in
// shared memory).
#pragma once
#include "ptx.cuh"
#include "qgemm_common.cuh"

namespace qgemm {

// ---------------------------------------------------------------------------
// tcgen05 wrappers
// ---------------------------------------------------------------------------
namespace t5 {
__device__ __forceinline__ void alloc(uint32_t* smem_slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_slot)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     ptx::smem_u32(bar))
                 : "memory");
}
// D[tmem] = A[smem] . B[smem]^T, 8-bit integer operands, s32 accumulate; overwrite (no accumulate)
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive columns -> 32 registers per thread
__device__ __forceinline__ void ld32(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// narrower forms of the same load: 32 lanes x 16 / 8 consecutive columns
__device__ __forceinline__ void ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ld8(uint32_t taddr, int (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
template <int N>
__device__ __forceinline__ void ldn(uint32_t taddr, int (&v)[N]) {
    static_assert(N == 8 || N == 16 || N == 32, "tcgen05.ld widths in use");
    if constexpr (N == 32) ld32(taddr, v);
    else if constexpr (N == 16) ld16(taddr, v);
    else ld8(taddr, v);
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
}  // namespace t5

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2: one issue slot, two IEEE results) ----
__device__ __forceinline__ uint64_t pk(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// s32 -> f32, exact for |x| < 2^22 (|sumi| <= 524288): integer add on the ALU pipe builds the
// bits of 12582912 + x, one packed FADD removes the bias exactly.  kCvtMagic = 0 uses I2FP.
#ifndef QGEMM_MMQ_CVT_MAGIC
#define QGEMM_MMQ_CVT_MAGIC 0
#endif
__device__ __forceinline__ uint64_t cvt2(int x0, int x1) {
#if QGEMM_MMQ_CVT_MAGIC == 2
    return pk(__int_as_float(x0), __int_as_float(x1));  // timing experiment only: what the fold costs without a conversion
#elif QGEMM_MMQ_CVT_MAGIC
    const uint64_t biased = pk(__int_as_float(x0 + 0x4B400000), __int_as_float(x1 + 0x4B400000));
    return fadd2(biased, pk(-12582912.0f, -12582912.0f));
#else
    return pk(__int2float_rn(x0), __int2float_rn(x1));
#endif
}
// accumulate in place: the tied operand keeps the accumulator pair in one register pair for the whole K loop
// (with a separate destination the register allocator rotated the pairs and paid a move per pair and block)
__device__ __forceinline__ void ffma2_acc(uint64_t& acc, uint64_t a, uint64_t b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void fadd2_acc(uint64_t& acc, uint64_t a) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(a)); }

// Two outputs of one token.  kRefSeq: the rounding sequence per element of fold_block_pre() (qgemm_common.cuh), i.e. of
// the reference GPU kernel -- bit-identical results.  Without it q4_1 / q5_1 use three FMA-pipe operations per pair
// instead of four, acc = fma(m_w, c_a, fma(d_w, d_a * sumi, acc)): the same terms, associated differently (agreement
// with the reference order ~1e-7 of max|C|); the other formats have only one sequence.
template <int WT, bool kRefSeq = true>
__device__ __forceinline__ uint64_t fold_pair(uint64_t acc, int x0, int x1, uint64_t dw, uint64_t mw, uint64_t da, uint64_t ca) {
    const uint64_t f = cvt2(x0, x1);
    if constexpr (WT == QGEMM_TYPE_Q4_0 || WT == QGEMM_TYPE_Q5_0) {
        ffma2_acc(acc, dw, ffma2(da, f, ca));
    } else if constexpr (WT == QGEMM_TYPE_Q4_1 || WT == QGEMM_TYPE_Q5_1) {
        if constexpr (kRefSeq) {
            fadd2_acc(acc, ffma2(fmul2(dw, da), f, fmul2(mw, ca)));
        } else {
            ffma2_acc(acc, dw, fmul2(da, f));
            ffma2_acc(acc, mw, ca);
        }
    } else {
        ffma2_acc(acc, fmul2(dw, da), f);
    }
    return acc;
}

}  // namespace qgemm
