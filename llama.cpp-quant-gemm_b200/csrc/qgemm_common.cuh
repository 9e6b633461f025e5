// qgemm_common.cuh -- shared device helpers: block formats, unaligned loads,
// nibble/5-bit expansion, the per-block scale fold.
//
// Formats are llama.cpp's (reference: compat/ggml_types.h:62-191); blocks are
// addressed as raw bytes here because weight blocks are only 2-byte aligned.
#pragma once
#include <cstdlib>

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qgemm.h"

namespace qgemm {

constexpr int kQK = 32;  // elements per block, every supported format

// ---------------------------------------------------------------------------
// Format traits.  Byte offsets inside one weight block.
// ---------------------------------------------------------------------------
template <int WT> struct Fmt;
template <> struct Fmt<QGEMM_TYPE_Q4_0> { static constexpr int bytes = 18, qs = 2, qh = -1, m = -1, bits = 4; };
template <> struct Fmt<QGEMM_TYPE_Q4_1> { static constexpr int bytes = 20, qs = 4, qh = -1, m = 2, bits = 4; };
template <> struct Fmt<QGEMM_TYPE_Q5_0> { static constexpr int bytes = 22, qs = 6, qh = 2, m = -1, bits = 5; };
template <> struct Fmt<QGEMM_TYPE_Q5_1> { static constexpr int bytes = 24, qs = 8, qh = 4, m = 2, bits = 5; };
template <> struct Fmt<QGEMM_TYPE_Q8_0> { static constexpr int bytes = 34, qs = 2, qh = -1, m = -1, bits = 8; };

constexpr int kQ81Bytes = 36;

__host__ __device__ inline int block_bytes(int type) {
    switch (type) {
    case QGEMM_TYPE_Q4_0: return 18;
    case QGEMM_TYPE_Q4_1: return 20;
    case QGEMM_TYPE_Q5_0: return 22;
    case QGEMM_TYPE_Q5_1: return 24;
    case QGEMM_TYPE_Q8_0: return 34;
    case QGEMM_TYPE_Q8_1: return 36;
    default: return 0;
    }
}

// ---------------------------------------------------------------------------
// Loads from 2-byte aligned memory.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_u16(const uint8_t* p) { return *reinterpret_cast<const uint16_t*>(p); }
__device__ __forceinline__ uint32_t ld_u32_a2(const uint8_t* p) { return ld_u16(p) | (ld_u16(p + 2) << 16); }
__device__ __forceinline__ float ld_half(const uint8_t* p) {
    return __half2float(__ushort_as_half((unsigned short)ld_u16(p)));
}
__device__ __forceinline__ float half_bits_to_float(uint32_t h) {
    return __half2float(__ushort_as_half((unsigned short)(h & 0xffffu)));
}

// ---------------------------------------------------------------------------
// Integer dot product of 4 packed bytes.  Weights enter UN-offset (0..15,
// 0..31) as unsigned bytes, activations as signed bytes -- the same integers
// the reference sums (include/gemm_reference.h:199-212).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int dp4a_us(uint32_t w_u8x4, int a_s8x4, int acc) {
    int r;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(w_u8x4), "r"(a_s8x4), "r"(acc));
    return r;
}
__device__ __forceinline__ int dp4a_ss(int w_s8x4, int a_s8x4, int acc) { return __dp4a(w_s8x4, a_s8x4, acc); }

// Spread 4 consecutive bits of qh (starting at bit `pos`) to bit 4 of each byte.
__device__ __forceinline__ uint32_t spread_qh4(uint32_t qh, int pos) {
    const uint32_t n = (qh >> pos) & 0xfu;
    // bit0->bit4, bit1->bit12, bit2->bit20, bit3->bit28
    return ((n * 0x00204081u) & 0x01010101u) << 4;
}

// One weight block expanded to 8 words of 4 x u8 (or s8 for q8_0):
// w[0..3] = elements 0..15, w[4..7] = elements 16..31.
template <int WT>
__device__ __forceinline__ void unpack_block(const uint8_t* blk, uint32_t (&w)[8]) {
    using F = Fmt<WT>;
    if constexpr (F::bits == 8) {
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = ld_u32_a2(blk + F::qs + 4 * i);
    } else {
        uint32_t qh = 0;
        if constexpr (F::bits == 5) qh = ld_u32_a2(blk + F::qh);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t v = ld_u32_a2(blk + F::qs + 4 * i);
            w[i] = v & 0x0f0f0f0fu;
            w[i + 4] = (v >> 4) & 0x0f0f0f0fu;
            if constexpr (F::bits == 5) {
                w[i] |= spread_qh4(qh, 4 * i);
                w[i + 4] |= spread_qh4(qh, 16 + 4 * i);
            }
        }
    }
}

template <int WT>
__device__ __forceinline__ int block_sumi(const uint32_t (&w)[8], const int (&a)[8]) {
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if constexpr (Fmt<WT>::bits == 8) s = dp4a_ss((int)w[i], a[i], s);
        else s = dp4a_us(w[i], a[i], s);
    }
    return s;
}

// ---------------------------------------------------------------------------
// The per-block scale fold.  Written with explicit fma/mul intrinsics so the
// rounding sequence is exactly what nvcc emits for the reference's GPU kernels
// (kernels/gemm/gemm_quant_formats.cuh:102,148,207,266,295 -- read from the
// SASS of that file built for sm_100a):
//   q4_0/q5_0: t = fma(d_a, sumi, -(off*s_a)); acc = fma(d_w, t, acc)
//   q4_1/q5_1: r = fma(d_w*d_a, sumi, (m_w*s_a)/4); acc = acc + r
//   q8_0:      acc = fma(d_w*d_a, sumi, acc)
// ---------------------------------------------------------------------------
struct ActScale { float d, s; };   // d_a, s_a of one q8_1 block
struct WScale { float d, m; };     // d_w, m_w of one weight block

template <int WT, bool kMsExact>
__device__ __forceinline__ float fold_block(float acc, int sumi, WScale w, ActScale a) {
    const float fs = __int2float_rn(sumi);
    if constexpr (WT == QGEMM_TYPE_Q4_0 || WT == QGEMM_TYPE_Q5_0) {
        const float off = (WT == QGEMM_TYPE_Q4_0) ? 8.0f : 16.0f;
        const float t = __fmaf_rn(a.d, fs, -__fmul_rn(off, a.s));
        return __fmaf_rn(w.d, t, acc);
    } else if constexpr (WT == QGEMM_TYPE_Q4_1 || WT == QGEMM_TYPE_Q5_1) {
        float ms = __fmul_rn(w.m, a.s);
        if constexpr (!kMsExact) ms = __fmul_rn(ms, 0.25f);
        const float r = __fmaf_rn(__fmul_rn(w.d, a.d), fs, ms);
        return __fadd_rn(acc, r);
    } else {
        return __fmaf_rn(__fmul_rn(w.d, a.d), fs, acc);
    }
}

// Same fold with the activation-only factor hoisted out of the row loop: the decode path
// keeps one activation block in registers for thousands of weight rows, so
//   q4_0/q5_0:  a.s := -(off * s_a)         (exact: power-of-two times a half)
//   q4_1/q5_1:  a.s := s_a / 4  (or s_a with QGEMM_MS_EXACT; (m*s)/4 == m*(s/4) exactly)
// is computed once by prep_act_scale().  Rounding sequence per block is unchanged.
template <int WT, bool kMsExact>
__device__ __forceinline__ ActScale prep_act_scale(float d_a, float s_a) {
    if constexpr (WT == QGEMM_TYPE_Q4_0) return {d_a, -__fmul_rn(8.0f, s_a)};
    else if constexpr (WT == QGEMM_TYPE_Q5_0) return {d_a, -__fmul_rn(16.0f, s_a)};
    else if constexpr (WT == QGEMM_TYPE_Q4_1 || WT == QGEMM_TYPE_Q5_1) return {d_a, kMsExact ? s_a : __fmul_rn(s_a, 0.25f)};
    else return {d_a, 0.0f};
}
template <int WT>
__device__ __forceinline__ float fold_block_pre(float acc, int sumi, WScale w, ActScale a) {
    const float fs = __int2float_rn(sumi);
    if constexpr (WT == QGEMM_TYPE_Q4_0 || WT == QGEMM_TYPE_Q5_0) {
        return __fmaf_rn(w.d, __fmaf_rn(a.d, fs, a.s), acc);
    } else if constexpr (WT == QGEMM_TYPE_Q4_1 || WT == QGEMM_TYPE_Q5_1) {
        return __fadd_rn(acc, __fmaf_rn(__fmul_rn(w.d, a.d), fs, __fmul_rn(w.m, a.s)));
    } else {
        return __fmaf_rn(__fmul_rn(w.d, a.d), fs, acc);
    }
}

template <int WT>
__device__ __forceinline__ WScale load_wscale(const uint8_t* blk) {
    WScale s;
    s.d = ld_half(blk);
    if constexpr (Fmt<WT>::m >= 0) s.m = ld_half(blk + Fmt<WT>::m);
    else s.m = 0.0f;
    return s;
}

// ---------------------------------------------------------------------------
// Fused all-gather: the decode kernels can store their slice of C straight into every
// peer GPU's copy of the gathered buffer (NVLink peer stores) and signal completion with
// system-scope counters, instead of a separate NCCL all-gather per GEMV.
//   * launch q = step * launches_per_step + launch_index (same sequence on every rank)
//   * before touching its activations, launch q waits until flag[rank] >= q * world:
//     every rank's launches < q have landed here (the activations may derive from them)
//   * the last CTA of a launch adds 1 to flag[r] on every rank r (release, system scope)
// ---------------------------------------------------------------------------
constexpr int kMaxPeers = 8;
struct PeerOut {
    int world;                    // 0 or 1: plain store to C
    int rank;
    float* C[kMaxPeers];          // rank r's destination for this launch's slice (same logical offset)
    uint32_t* flag[kMaxPeers];    // rank r's arrival counter
    uint32_t* done;               // local CTA-completion counter (returns to 0 after each launch)
    const uint32_t* step;         // local step counter (qgemm_peer_step_advance)
    uint32_t lps, li;             // launches per step, launches of this step that must have landed first
    float* mc;                    // NVLS multicast mapping of the same slice on every rank (nullptr: store per rank)
    long long moff[kMaxPeers];    // grouped launch: element offset of matrix m's slice from C[r]
    int dbg;                      // tuning aid: 1 skip wait, 2 local store only, 4 skip per-thread fence
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_sys_add(uint32_t* p, uint32_t v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// one thread: block until every earlier launch of every rank has landed in local memory
__device__ __forceinline__ void peer_wait_prior(const PeerOut& po) {
    if (po.world > 1 && !(po.dbg & 1)) {
        const uint32_t target = ((*po.step) * po.lps + po.li) * (uint32_t)po.world;
        while ((int32_t)(ld_acquire_sys(po.flag[po.rank]) - target) < 0) __nanosleep(64);
    }
}
// every storing thread calls this after its last store; then, after a CTA-wide barrier, one
// thread calls peer_signal_done()
__device__ __forceinline__ void peer_store(const PeerOut& po, float* C, int64_t idx, float v, int m = 0) {
    if (po.world > 1) {
        idx += po.moff[m];
        if (po.dbg & 2) { po.C[po.rank][idx] = v; return; }
        if (po.mc) { po.mc[idx] = v; return; }   // one store, replicated to every rank by the switch
#pragma unroll 1
        for (int q = 0; q < po.world; q++) {   // staggered start: ranks do not all hit the same receiver first
            int r = po.rank + 1 + q;
            if (r >= po.world) r -= po.world;
            po.C[r][idx] = v;
        }
    } else {
        C[idx] = v;
    }
}
__device__ __forceinline__ void peer_signal_done(const PeerOut& po, unsigned grid) {
    if (po.world > 1) {
        // CTA-local barrier already ordered this CTA's peer stores before this thread.  A device-scope
        // fence per CTA + one system-scope fence by the last CTA (release cumulativity) replaces a
        // system-scope fence per CTA, which serialises chip-wide (measured: +12 us per launch).
        if (!(po.dbg & 8)) __threadfence();
        if (atomicAdd(po.done, 1u) == grid - 1) {   // last CTA of this launch on this GPU
            *po.done = 0u;
            if (!(po.dbg & 16)) __threadfence_system();
            // fence.sys + relaxed system-scope atomics = one release pattern for all peers
            // (a red.release.sys per peer costs a system fence each: measured +1.6 us per peer)
            for (int r = 0; r < po.world; r++) atomicAdd_system(po.flag[r], 1u);
        }
    }
}

// ---------------------------------------------------------------------------
// Tuning aids (QGEMM_* environment overrides used by profiles/ and the experiments DESIGN.md cites): read once
// per process and call site, never on the launch path after the first call.
// ---------------------------------------------------------------------------
#define QGEMM_ENV(name) ([]() -> const char* { static const char* const v = getenv(name); return v; }())

// ---------------------------------------------------------------------------
// Host-side launch bookkeeping (defined in qgemm_abi.cu)
// ---------------------------------------------------------------------------
void note_launch(int n = 1);
// Opt a kernel in to `smem` bytes of dynamic shared memory; remembered per (device, kernel), thread-safe.
cudaError_t smem_optin(const void* kernel, size_t smem);

}  // namespace qgemm
