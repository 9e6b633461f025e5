// mmq.cu -- prefill path (T >= 64): block-quantized GEMM on the 5th-gen tensor cores.
//
// Replaces the reference's many-token kernels (include/gemm_cuda_naive.cuh:158-249,
// gemm_cuda_tiled.cuh:191-269, gemm_cuda_dp4a.cuh:158-403,
// kernels/gemm/gemm_quant_formats.cuh:312-334), all CUDA-core dp4a/scalar code.
//
// One tcgen05.mma kind::i8 instruction has K = 32 for 8-bit operands = exactly one
// quantization block, so sumi[t,f,b] falls out of each instruction unmodified:
//   D(TMEM, s32)[128 tokens x BN rows] = A(smem, s8 activations) . B(smem, u8/s8 weights)^T
// with the weights fed UN-offset (0..15 / 0..31), like the reference's integer sum
// (include/gemm_reference.h:199-212).  The per-block scale fold
//   acc[t,f] = fma(d_w, fma(d_a, float(sumi), -8*s_a), acc)       (q4_0; see qgemm_common.cuh)
// runs on the CUDA cores from registers, in block order b = 0..nb-1, with the exact
// FMA sequence of the reference GPU kernel -- C is bit-identical to it.
//
// Pipeline (one persistent CTA per SM, 12 warps):
//   warp 0      producer: 1-D bulk async copies (TMA engine) of pre-swizzled operand tiles
//               + scale slabs into a 4-stage smem ring (stage = 128 K-elements = 4 blocks)
//   warp 1      one elected thread issues the MMAs; each block gets its own TMEM
//               accumulator buffer out of a ring of 512/BN, tcgen05.commit -> mbarrier
//   warp 2      TMEM allocator
//   warps 4-11  epilogue: tcgen05.ld the s32 tile (lane = token, 64 columns per thread),
//               release the TMEM buffer at once, fold into 64 fp32 register accumulators
//
// Operands come from a prepass (qgemm workspace): q8_1 AoS -> s8 K-major tiles + (d, c)
// slabs; weight blocks -> u8 K-major tiles + d (+m) slabs, both already in the 128-byte
// swizzle image the UMMA smem descriptor expects, so the producer needs no tensor map.
#include <cstdlib>

#include "ptx.cuh"
#include "qgemm_common.cuh"
#include "tc05.cuh"

#ifdef QGEMM_MMQ_PROFILE
#include <cstdio>
#define PROF_DECL long long pf_wait = 0, pf_wait2 = 0, pf_t0 = clock64(), pf_n = 0
#define PROF_WAIT(acc, stmt) do { const long long c0_ = clock64(); stmt; acc += clock64() - c0_; } while (0)
#else
#define PROF_DECL
#define PROF_WAIT(acc, stmt) stmt
#endif

namespace qgemm {

constexpr int kBM = 128;        // tokens per tile = TMEM lanes
constexpr int kBN = 128;        // weight rows per tile = TMEM columns per buffer
constexpr int kTmemCols = 512;
constexpr int kTmemBufs = kTmemCols / kBN;
constexpr int kKC = 128;        // K elements per smem stage (4 blocks), one 128-byte swizzle row
constexpr int kBlocksPerStage = kKC / 32;
constexpr int kMmqStages = 4;
constexpr int kEpiWarps = 16;       // 4 per TMEM lane quarter, 32 columns each
constexpr int kEpiCols = kBN / (kEpiWarps / 4);
constexpr int kMmqThreads = (4 + kEpiWarps) * 32;

// stage layout (bytes); operand tiles 1024-byte aligned for SWIZZLE_128B
constexpr int kStageA = 0;                                   // [128 tokens][128 B]
constexpr int kStageW = kStageA + kBM * kKC;                 // [kBN rows][128 B]
constexpr int kStageAS = kStageW + kBN * kKC;                // [4 blocks][128 tokens] float2 (d_a, c_a)
constexpr int kStageWS = kStageAS + kBlocksPerStage * kBM * 8;   // [4 blocks][kBN] float d_w
constexpr int kStageWM = kStageWS + kBlocksPerStage * kBN * 4;   // [4 blocks][kBN] float m_w
constexpr int kStageBytes = kStageWM + kBlocksPerStage * kBN * 4;
static_assert(kStageBytes % 1024 == 0, "stage must keep 1024-byte alignment");
constexpr int kMmqSmem = 1024 /*align slack*/ + kMmqStages * kStageBytes + 256 /*barriers*/;
constexpr int kOutTileBytes = kBM * kBN * 4;           // one finished fp32 tile, [f][t]
constexpr int kMmqSmemOut = kMmqSmem + kOutTileBytes;  // peer mode: + the staging tile for the bulk stores
static_assert(kMmqSmemOut <= 227 * 1024, "staging tile must fit next to the operand ring");

// ---------------------------------------------------------------------------
// workspace layout (all offsets 1024-byte aligned)
// ---------------------------------------------------------------------------
struct MmqWs {
    size_t a8, as, w8, ws, wm, total;
    int Tpad, Fpad, nkc;
};
static MmqWs mmq_layout(int T, int F, int K) {
    MmqWs L;
    L.Tpad = (T + kBM - 1) / kBM * kBM;
    L.Fpad = (F + kBN - 1) / kBN * kBN;
    L.nkc = (K + kKC - 1) / kKC;
    const int nbp = L.nkc * kBlocksPerStage;  // blocks, padded to whole stages
    auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
    size_t o = 0;
    L.a8 = o; o = up(o + (size_t)L.nkc * L.Tpad * kKC);
    L.as = o; o = up(o + (size_t)nbp * L.Tpad * 8);
    L.w8 = o; o = up(o + (size_t)L.nkc * L.Fpad * kKC);
    L.ws = o; o = up(o + (size_t)nbp * L.Fpad * 4);
    L.wm = o; o = up(o + (size_t)nbp * L.Fpad * 4);
    L.total = o;
    return L;
}

// Offline weight pre-pack: the weight half of the layout above in a
// buffer of its own, built once for static weights: [ w8 | ws | wm ].
struct MmqPack { size_t w8, ws, wm, total; int Fpad, nkc; };
static MmqPack mmq_pack_layout(int F, int K) {
    MmqPack P;
    P.Fpad = (F + kBN - 1) / kBN * kBN;
    P.nkc = (K + kKC - 1) / kKC;
    const int nbp = P.nkc * kBlocksPerStage;
    auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
    size_t o = 0;
    P.w8 = o; o = up(o + (size_t)P.nkc * P.Fpad * kKC);
    P.ws = o; o = up(o + (size_t)nbp * P.Fpad * 4);
    P.wm = o; o = up(o + (size_t)nbp * P.Fpad * 4);
    P.total = o;
    return P;
}

// ---------------------------------------------------------------------------
// prepass 1: q8_1 AoS -> swizzled s8 tiles + (d_a, c_a) slabs
//   a8[kc][t][16-byte chunk c ^ (t & 7)]            (chunk c = (b & 3) * 2 + {0, 1})
//   as[t / 128][b][t % 128] = (d_a, coef * s_a)     coef: -8 (q4_0), -16 (q5_0), 1/4 or 1 (q4_1/q5_1), 0 (q8_0)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void repack_act_body(int64_t gid, const uint8_t* __restrict__ act, uint8_t* __restrict__ a8,
                                                float2* __restrict__ as, int T, int Tpad, int nb, int nbp, float coef, int pairs = 0) {
    if (gid >= (int64_t)Tpad * nbp) return;
    // 4 consecutive lanes = the 4 blocks of one 128-byte operand row: contiguous reads (4 x 36 B) and
    // a contiguous 128-byte write per row, 8 rows per warp
    const int j4 = (int)(gid & 3);
    const int t = (int)((gid >> 2) % Tpad), b = (int)((gid >> 2) / Tpad) * 4 + j4;
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    float2 sc = make_float2(0.f, 0.f);
    if (t < T && b < nb) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(act + ((size_t)t * nb + b) * kQ81Bytes);
        const uint32_t ds = __ldg(src);
        sc.x = half_bits_to_float(ds);
        sc.y = __fmul_rn(coef, half_bits_to_float(ds >> 16));
        q0 = make_uint4(__ldg(src + 1), __ldg(src + 2), __ldg(src + 3), __ldg(src + 4));
        q1 = make_uint4(__ldg(src + 5), __ldg(src + 6), __ldg(src + 7), __ldg(src + 8));
    }
    const int kc = b >> 2, c = (b & 3) * 2;
    uint8_t* row = a8 + ((size_t)kc * Tpad + t) * kKC;
    *reinterpret_cast<uint4*>(row + ((c ^ (t & 7)) << 4)) = q0;
    *reinterpret_cast<uint4*>(row + (((c + 1) ^ (t & 7)) << 4)) = q1;
    if (!pairs) {
        as[((size_t)(t / kBM) * nbp + b) * kBM + (t % kBM)] = sc;
    } else {   // (d_a, d_a') (c_a, c_a') per pair of tokens: the form the weight-major kernel reads as packed pairs
        float* asf = reinterpret_cast<float*>(as + ((size_t)(t / kBM) * nbp + b) * kBM) + (((t % kBM) >> 1) << 2) + (t & 1);
        asf[0] = sc.x;
        asf[2] = sc.y;
    }
}
__global__ void __launch_bounds__(256)
mmq_repack_act_kernel(const uint8_t* __restrict__ act, uint8_t* __restrict__ a8, float2* __restrict__ as, int T,
                      int Tpad, int nb, int nbp, float coef, unsigned* zero = nullptr, int nzero = 0, int pairs = 0) {
    ptx::griddep_launch_dependents();   // the GEMM kernel behind this one may set itself up while we run
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < nzero) zero[gid] = 0u;    // split-K arrival counters of the GEMM behind us
    repack_act_body(gid, act, a8, as, T, Tpad, nb, nbp, coef, pairs);
}

// ---------------------------------------------------------------------------
// prepass 2: weight blocks -> swizzled u8 (s8 for q8_0) tiles + d (+m) slabs
//   w8[kc][f][chunk ^ (f & 7)],  ws[f / BN][b][f % BN] = d_w,  wm[...] = m_w
// ---------------------------------------------------------------------------
template <int WT>
__device__ __forceinline__ void unpack_weight_body(int64_t gid, const uint8_t* __restrict__ wgt, uint8_t* __restrict__ w8,
                                                   float* __restrict__ ws, float* __restrict__ wm, int F, int Fpad, int nb,
                                                   int nbp) {
    using Fm = Fmt<WT>;
    if (gid >= (int64_t)Fpad * nbp) return;
    const int j4 = (int)(gid & 3);  // lanes 4i..4i+3 = the 4 blocks of one operand row (see the activation prepass)
    const int f = (int)((gid >> 2) % Fpad), b = (int)((gid >> 2) / Fpad) * 4 + j4;
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    WScale sc{0.f, 0.f};
    if (f < F && b < nb) {
        const uint8_t* blk = wgt + ((size_t)f * nb + b) * Fm::bytes;
        unpack_block<WT>(blk, w);
        sc = load_wscale<WT>(blk);
    }
    const int kc = b >> 2, c = (b & 3) * 2;
    uint8_t* row = w8 + ((size_t)kc * Fpad + f) * kKC;
    *reinterpret_cast<uint4*>(row + ((c ^ (f & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(row + (((c + 1) ^ (f & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
    const size_t so = ((size_t)(f / kBN) * nbp + b) * kBN + (f % kBN);
    ws[so] = sc.d;
    if constexpr (Fm::m >= 0) wm[so] = sc.m;
}
template <int WT>
__global__ void __launch_bounds__(256)
mmq_unpack_weight_kernel(const uint8_t* __restrict__ wgt, uint8_t* __restrict__ w8, float* __restrict__ ws,
                         float* __restrict__ wm, int F, int Fpad, int nb, int nbp) {
    unpack_weight_body<WT>((int64_t)blockIdx.x * blockDim.x + threadIdx.x, wgt, w8, ws, wm, F, Fpad, nb, nbp);
}
// both prepasses as one launch: blocks [0, act_blocks) repack the activations, the rest unpack the weights
template <int WT>
__global__ void __launch_bounds__(256)
mmq_prepass_kernel(const uint8_t* __restrict__ act, uint8_t* __restrict__ a8, float2* __restrict__ as, int T, int Tpad,
                   float coef, const uint8_t* __restrict__ wgt, uint8_t* __restrict__ w8, float* __restrict__ ws,
                   float* __restrict__ wm, int F, int Fpad, int nb, int nbp, unsigned act_blocks) {
    ptx::griddep_launch_dependents();
    if (blockIdx.x < act_blocks)
        repack_act_body((int64_t)blockIdx.x * blockDim.x + threadIdx.x, act, a8, as, T, Tpad, nb, nbp, coef);
    else
        unpack_weight_body<WT>((int64_t)(blockIdx.x - act_blocks) * blockDim.x + threadIdx.x, wgt, w8, ws, wm, F, Fpad, nb, nbp);
}

struct MmqParams {
    const uint8_t* a8;
    const float2* as;
    const uint8_t* w8;
    const float* ws;
    const float* wm;
    float* C;
    int32_t* sumi;   // non-null: dump the raw s32 block sums instead of folding
    int T, F, nb, nkc, Tpad, Fpad;
    int64_t ldc_t, ldc_f;
    int tiles_m, tiles_n;
    int dbg;  // tuning aid: 1 = skip the fold, 2 = also load only half of the TMEM columns, 3 = no TMEM load, 4 = no C store
    PeerOut peer;  // fused all-gather: the epilogue stores its tile into every rank's gathered C (world <= 1: plain store)
    int tma_out;   // 1: finished tiles are staged in shared memory and leave through bulk copies (needs ldc_t == 1,
                   //    16-byte rows; kMmqSmemOut bytes of dynamic shared memory)
};

template <int WT, bool kDump>
__global__ void __launch_bounds__(kMmqThreads, 1) mmq_kernel(const MmqParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte aligned base (SWIZZLE_128B atoms); offset arithmetic keeps the shared address space
    const uint32_t raw = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kMmqStages * kStageBytes);  // [stages]
    uint64_t* empty = full + kMmqStages;                                            // [stages]
    uint64_t* tfull = empty + kMmqStages;                                           // [kTmemBufs]
    uint64_t* tempty = tfull + kTmemBufs;                                           // [kTmemBufs]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kTmemBufs);
    float* out_tile = reinterpret_cast<float*>(smem + kMmqStages * kStageBytes + 256);  // [kBN][kBM], only with p.tma_out

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkc = p.nkc;
    const int nbp = nkc * kBlocksPerStage;
    const int ntiles = p.tiles_m * p.tiles_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMmqStages; s++) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1 + kEpiWarps);  // MMA commit + every epilogue warp
        }
        for (int i = 0; i < kTmemBufs; i++) {
            ptx::mbar_init(&tfull[i], 1);
            ptx::mbar_init(&tempty[i], kEpiWarps);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) t5::alloc(tmem_slot, kTmemCols);
    t5::fence_before();
    __syncthreads();
    t5::fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::griddep_wait();   // launched behind the operand prepass with programmatic serialization: barriers and TMEM
                           // were set up while it ran; from here on its output (and C) may be touched

    if (warp == 0) {
        // ===================== producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            PROF_DECL;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int mt = tile % p.tiles_m, nt = tile / p.tiles_m;
                for (int kc = 0; kc < nkc; kc++) {
                    PROF_WAIT(pf_wait, ptx::mbar_wait_backoff(&empty[s], ph ^ 1));
                    uint8_t* st = smem + s * kStageBytes;
                    constexpr uint32_t bytes = kBM * kKC + kBN * kKC + kBlocksPerStage * kBM * 8 +
                                               kBlocksPerStage * kBN * 4 * (Fmt<WT>::m >= 0 ? 2 : 1);
                    ptx::mbar_arrive_expect_tx(&full[s], bytes);
                    ptx::bulk_g2s(st + kStageA, p.a8 + ((size_t)kc * p.Tpad + (size_t)mt * kBM) * kKC, kBM * kKC, &full[s]);
                    ptx::bulk_g2s(st + kStageW, p.w8 + ((size_t)kc * p.Fpad + (size_t)nt * kBN) * kKC, kBN * kKC, &full[s]);
                    ptx::bulk_g2s(st + kStageAS, p.as + ((size_t)mt * nbp + (size_t)kc * kBlocksPerStage) * kBM,
                                  kBlocksPerStage * kBM * 8, &full[s]);
                    ptx::bulk_g2s(st + kStageWS, p.ws + ((size_t)nt * nbp + (size_t)kc * kBlocksPerStage) * kBN,
                                  kBlocksPerStage * kBN * 4, &full[s]);
                    if constexpr (Fmt<WT>::m >= 0)
                        ptx::bulk_g2s(st + kStageWM, p.wm + ((size_t)nt * nbp + (size_t)kc * kBlocksPerStage) * kBN,
                                      kBlocksPerStage * kBN * 4, &full[s]);
                    if (++s == kMmqStages) { s = 0; ph ^= 1; }
                }
            }
#ifdef QGEMM_MMQ_PROFILE
            if (blockIdx.x == 0 && (p.dbg & 32)) printf("producer: total %lld, waiting for empty %lld\n", clock64() - pf_t0, pf_wait);
#endif
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // instruction descriptor: D = s32, A = s8 (activations), B = u8 (s8 for q8_0), both K-major
        constexpr uint32_t idesc = (2u << 4) | (1u << 7) | ((Fmt<WT>::bits == 8 ? 1u : 0u) << 10) |
                                   ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        int s = 0, buf = 0;
        uint32_t ph = 0, tph = 0;
        PROF_DECL;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int kc = 0; kc < nkc; kc++) {
                PROF_WAIT(pf_wait, ptx::mbar_wait_backoff(&full[s], ph));
                t5::fence_after();
                const uint32_t sa = ptx::smem_u32(smem + s * kStageBytes + kStageA);
                const uint32_t sw = ptx::smem_u32(smem + s * kStageBytes + kStageW);
                const uint64_t adesc = t5::smem_desc(sa), bdesc = t5::smem_desc(sw);
#pragma unroll
                for (int j = 0; j < kBlocksPerStage; j++) {
                    PROF_WAIT(pf_wait2, ptx::mbar_wait_backoff(&tempty[buf], tph ^ 1));
                    t5::fence_after();
                    if (lane == 0) {
                        // one instruction = one quantization block (K = 32 bytes = +2 in the >>4 address field)
                        t5::mma_i8(tmem_base + buf * kBN, adesc + 2 * j, bdesc + 2 * j, idesc, 0u);
                        if (p.dbg & 16) t5::mma_i8(tmem_base + buf * kBN, adesc + 2 * j, bdesc + 2 * j, idesc, 1u);  // timing experiment
                        t5::commit(&tfull[buf]);
                    }
                    __syncwarp();
                    if (++buf == kTmemBufs) { buf = 0; tph ^= 1; }
                }
                if (lane == 0) t5::commit(&empty[s]);  // operand tiles consumed once these MMAs retire
                __syncwarp();
                if (++s == kMmqStages) { s = 0; ph ^= 1; }
            }
        }
#ifdef QGEMM_MMQ_PROFILE
        if (blockIdx.x == 0 && lane == 0 && (p.dbg & 32)) printf("mma: total %lld, waiting for full %lld, for tempty %lld\n", clock64() - pf_t0, pf_wait, pf_wait2);
#endif
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;
        const int quarter = warp & 3;           // TMEM lane quarter this warp may touch
        const int cgrp = ew >> 2;               // which kEpiCols-wide column group
        const int row = quarter * 32 + lane;    // token row inside the tile
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        static_assert(kEpiCols == 32, "one tcgen05.ld.x32 per block per thread");
        int s = 0, buf = 0;
        uint32_t ph = 0, tph = 0;
        PROF_DECL;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int mt = tile % p.tiles_m, nt = tile / p.tiles_m;
            uint64_t acc[kEpiCols / 2];  // fp32 accumulators as packed pairs (columns 2i, 2i+1)
#pragma unroll
            for (int i = 0; i < kEpiCols / 2; i++) acc[i] = 0ull;
            for (int kc = 0; kc < nkc; kc++) {
                PROF_WAIT(pf_wait, ptx::mbar_wait(&full[s], ph));  // scale slabs of this stage are visible
                const uint8_t* st = smem + s * kStageBytes;
#pragma unroll (Fmt<WT>::m >= 0 ? 1 : 2)
                for (int j = 0; j < kBlocksPerStage; j++) {
                    const int b = kc * kBlocksPerStage + j;
                    PROF_WAIT(pf_wait2, ptx::mbar_wait(&tfull[buf], tph));
                    t5::fence_after();
                    int x[kEpiCols];
                    t5::ld32(tmem_base + lane_addr + buf * kBN + cgrp * kEpiCols, x);
                    t5::wait_ld();
                    t5::fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tempty[buf]);  // values are in registers: free the buffer
                    if (++buf == kTmemBufs) { buf = 0; tph ^= 1; }

                    if constexpr (kDump) {
                        const int t = mt * kBM + row;
                        if (t < p.T && b < p.nb) {
#pragma unroll
                            for (int i = 0; i < kEpiCols; i++) {
                                const int f = nt * kBN + cgrp * kEpiCols + i;
                                if (f < p.F) p.sumi[((size_t)t * p.F + f) * p.nb + b] = x[i];
                            }
                        }
                        continue;
                    }
                    const float2 a = reinterpret_cast<const float2*>(st + kStageAS)[j * kBM + row];
                    const uint64_t da = pk(a.x, a.x), ca = pk(a.y, a.y);
                    const ulonglong2* dw2 = reinterpret_cast<const ulonglong2*>(st + kStageWS) + (j * kBN + cgrp * kEpiCols) / 4;
                    const ulonglong2* mw2 = reinterpret_cast<const ulonglong2*>(st + kStageWM) + (j * kBN + cgrp * kEpiCols) / 4;
#pragma unroll
                    for (int i4 = 0; i4 < kEpiCols / 4; i4++) {
                        const ulonglong2 dw = dw2[i4];  // d_w of columns 4*i4 .. 4*i4+3 (broadcast read)
                        ulonglong2 mw = make_ulonglong2(0ull, 0ull);
                        if constexpr (Fmt<WT>::m >= 0) mw = mw2[i4];
                        acc[2 * i4] = fold_pair<WT>(acc[2 * i4], x[4 * i4], x[4 * i4 + 1], dw.x, mw.x, da, ca);
                        acc[2 * i4 + 1] = fold_pair<WT>(acc[2 * i4 + 1], x[4 * i4 + 2], x[4 * i4 + 3], dw.y, mw.y, da, ca);
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&empty[s]);  // scale slabs consumed
                if (++s == kMmqStages) { s = 0; ph ^= 1; }
            }
            if constexpr (!kDump) {
                const int t = mt * kBM + row;
                if (p.tma_out) {
                    // Fused all-gather, bulk variant: the tile is staged in shared memory as [f][t] and carried to
                    // every rank's gathered C (this rank's included) by the TMA engine, 512-byte rows, while the
                    // epilogue warps go on folding the next tile.  Lane 0 of every epilogue warp issues the rows of
                    // 8 f values for all ranks and owns their bulk group.
                    if (lane == 0) ptx::bulk_wait_read();          // my rows of the previous tile have left smem
                    ptx::bar_sync(3, kEpiWarps * 32);
#pragma unroll
                    for (int i = 0; i < kEpiCols / 2; i++) {
                        float v0, v1;
                        unpk(acc[i], v0, v1);
                        out_tile[(cgrp * kEpiCols + 2 * i) * kBM + row] = v0;
                        out_tile[(cgrp * kEpiCols + 2 * i + 1) * kBM + row] = v1;
                    }
                    ptx::fence_proxy_async();
                    ptx::bar_sync(3, kEpiWarps * 32);
                    if (lane == 0) {
                        const int t0 = mt * kBM;
                        const uint32_t bytes = (uint32_t)min(kBM, p.T - t0) * 4u;
                        constexpr int kRowsPerWarp = kBN / kEpiWarps;
#pragma unroll 1
                        for (int q = 0; q < p.peer.world; q++) {
                            // start at the next rank and wrap: at any moment the ranks target different receivers
                            // instead of all converging on rank 0's links first
                            int r = p.peer.rank + 1 + q;
                            if (r >= p.peer.world) r -= p.peer.world;
                            float* Cr = p.peer.C[r];
#pragma unroll 1
                            for (int k = 0; k < kRowsPerWarp; k++) {
                                const int fl = ew * kRowsPerWarp + k;
                                const int f = nt * kBN + fl;
                                if (f < p.F) ptx::bulk_s2g(Cr + (int64_t)f * p.ldc_f + t0, out_tile + fl * kBM, bytes);
                            }
                        }
                        ptx::bulk_commit();
                    }
                } else if (t < p.T && !(p.dbg & 4)) {
                    // world > 1 without the staging tile (odd strides): the same values go to every rank's copy of the
                    // gathered buffer straight from the registers.  Measured on 2 B200s (4096 x 14336 x 8192): +0.05 ms
                    // for the local store pass, +0.2 ms for the remote one; at 8 ranks the bursts stall the epilogue.
                    // With an NVLS multicast mapping (p.peer.mc) one pass is enough: the switch replicates each store.
                    const int nrank = (p.peer.world > 1 && !p.peer.mc) ? p.peer.world : 1;
#pragma unroll 1
                    for (int q = 0; q < nrank; q++) {
                        int r = p.peer.rank + 1 + q;   // staggered start, see above
                        if (r >= nrank) r -= nrank;
                        float* crow = (nrank > 1 ? p.peer.C[r] : (p.peer.mc ? p.peer.mc : p.C)) + (int64_t)t * p.ldc_t;
#pragma unroll
                        for (int i = 0; i < kEpiCols / 2; i++) {
                            float v0, v1;
                            unpk(acc[i], v0, v1);
                            const int f = nt * kBN + cgrp * kEpiCols + 2 * i;
                            if (f < p.F) crow[(int64_t)f * p.ldc_f] = v0;
                            if (f + 1 < p.F) crow[(int64_t)(f + 1) * p.ldc_f] = v1;
                        }
                    }
                }
            }
        }
#ifdef QGEMM_MMQ_PROFILE
        if (blockIdx.x == 0 && lane == 0 && (p.dbg & 32) && (ew == 0 || ew == 15)) printf("epilogue warp %d: total %lld, waiting for full %lld, for tfull %lld\n", ew, clock64() - pf_t0, pf_wait, pf_wait2);
#endif
        if constexpr (!kDump) {
            if (p.tma_out && lane == 0) {   // every bulk store of this thread is complete before the CTA signals
                ptx::bulk_wait_all();
                ptx::fence_proxy_async_all();
            }
        }
    }
    t5::fence_before();
    __syncthreads();
    if (warp == 2) t5::dealloc(tmem_base, kTmemCols);
    if constexpr (!kDump) {
        if (threadIdx.x == 0) peer_signal_done(p.peer, gridDim.x);  // the barrier above ordered every epilogue store
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
bool mmq_supported(int wtype, const void* act, const void* wgt, int T, int F, int K) {
    (void)wgt;
    if (block_bytes(wtype) == 0 || T < 1 || F < 1 || K < 32) return false;
    return reinterpret_cast<uintptr_t>(act) % 4 == 0;
}

bool mmq_native_supported(int wtype, const void* wgt, int T, int F, int K);
size_t mmq_native_split_bytes(int T, int F, int K, uint32_t flags, int num_sms);
unsigned* mmq_native_split_counters(int T, int F, int K, uint32_t flags, int num_sms, void* split_ws, size_t split_ws_bytes, bool dump,
                                    const PeerOut* peer, int* count);
constexpr int kSplitPlanSms = 148;   // workspace sizing happens without a device: plan for a full B200

// Scratch of the tensor-core path.  Shapes the native-layout kernel takes (K % 256 == 0) need the repacked activations
// only; the prepass kernel also keeps the unpacked weights there.  The pointer-free form (qgemm_workspace_bytes) assumes
// a 16-byte aligned weight base, which every allocator gives; mmq_workspace_need() is what a concrete call requires.
size_t mmq_workspace_bytes(int wtype, int T, int F, int K) {
    if (T < 1 || F < 1 || K < 32) return 0;
    const MmqWs L = mmq_layout(T, F, K);
    // the split-K scratch of small-T calls is sized for the default flags (FOLD_REFSEQ needs none)
    return mmq_native_supported(wtype, nullptr, T, F, K) ? L.w8 + mmq_native_split_bytes(T, F, K, 0, kSplitPlanSms) : L.total;
}
size_t mmq_workspace_need(int wtype, const void* wgt, int T, int F, int K, uint32_t flags) {
    if (T < 1 || F < 1 || K < 32) return 0;
    const MmqWs L = mmq_layout(T, F, K);
    if (flags & QGEMM_WEIGHTS_PREPACKED) return L.w8;
    // split-K scratch is optional: a call that brings only the activation part runs unsplit
    if (mmq_native_supported(wtype, wgt, T, F, K) && !QGEMM_ENV("QGEMM_MMQ_LEGACY")) return L.w8;
    return L.total;
}

cudaError_t launch_mmq_native(int wtype, const uint8_t* a8, const float2* as, const void* wgt, float* C, int32_t* sumi, int T,
                              int F, int K, int Tpad, int64_t ldc_t, int64_t ldc_f, uint32_t flags, int num_sms, cudaStream_t st,
                              const PeerOut* peer, void* split_ws, size_t split_ws_bytes, int tokn = 0);
int mmq_native_tokn(int T, const PeerOut* peer, bool dump, uint32_t flags);

template <int WT>
static cudaError_t launch_mmq_t(const void* act, const void* wgt, float* C, int32_t* sumi, int T, int F, int K,
                                int64_t ldc_t, int64_t ldc_f, uint32_t flags, void* ws, size_t ws_bytes, int num_sms, cudaStream_t st,
                                const PeerOut* peer) {
    const MmqWs L = mmq_layout(T, F, K);
    const int nb = K / 32, nbp = L.nkc * kBlocksPerStage;
    uint8_t* base = (uint8_t*)ws;
    float coef = 0.f;
    if (WT == QGEMM_TYPE_Q4_0) coef = -8.f;
    else if (WT == QGEMM_TYPE_Q5_0) coef = -16.f;
    else if (WT == QGEMM_TYPE_Q4_1 || WT == QGEMM_TYPE_Q5_1) coef = (flags & QGEMM_MS_EXACT) ? 1.f : 0.25f;
    const uint8_t* w8 = base + L.w8;
    const float* wsp = (const float*)(base + L.ws);
    const float* wmp = (const float*)(base + L.wm);
    const unsigned act_blocks = (unsigned)(((int64_t)L.Tpad * nbp + 255) / 256);
    if (!(flags & QGEMM_WEIGHTS_PREPACKED) && mmq_native_supported(WT, wgt, T, F, K) && !QGEMM_ENV("QGEMM_MMQ_LEGACY")) {
        // native blocks are unpacked inside the kernel (mmq_native.cu): only the activations are repacked per call
        int ncount = 0;
        unsigned* counters = mmq_native_split_counters(T, F, K, flags, num_sms, ws_bytes > L.w8 ? base + L.w8 : nullptr,
                                                       ws_bytes > L.w8 ? ws_bytes - L.w8 : 0, sumi != nullptr, peer, &ncount);
        const int tokn = mmq_native_tokn(T, peer, sumi != nullptr, flags);   // small batch: weight-major kernel form
        mmq_repack_act_kernel<<<act_blocks, 256, 0, st>>>((const uint8_t*)act, base + L.a8, (float2*)(base + L.as), T, L.Tpad,
                                                          nb, nbp, coef, counters, ncount, tokn != 0);
        note_launch();
        if (cudaError_t e = cudaGetLastError()) return e;
        return launch_mmq_native(WT, base + L.a8, (const float2*)(base + L.as), wgt, C, sumi, T, F, K, L.Tpad, ldc_t, ldc_f, flags,
                                 num_sms, st, peer, ws_bytes > L.w8 ? base + L.w8 : nullptr, ws_bytes > L.w8 ? ws_bytes - L.w8 : 0, tokn);
    }
    if (flags & QGEMM_WEIGHTS_PREPACKED) {  // `wgt` is a qgemm_prepack_weights() buffer: nothing to unpack
        const MmqPack P = mmq_pack_layout(F, K);
        w8 = (const uint8_t*)wgt + P.w8;
        wsp = (const float*)((const uint8_t*)wgt + P.ws);
        wmp = (const float*)((const uint8_t*)wgt + P.wm);
        mmq_repack_act_kernel<<<act_blocks, 256, 0, st>>>((const uint8_t*)act, base + L.a8, (float2*)(base + L.as), T, L.Tpad,
                                                          nb, nbp, coef);
    } else {
        const unsigned w_blocks = (unsigned)(((int64_t)L.Fpad * nbp + 255) / 256);
        mmq_prepass_kernel<WT><<<act_blocks + w_blocks, 256, 0, st>>>(
            (const uint8_t*)act, base + L.a8, (float2*)(base + L.as), T, L.Tpad, coef, (const uint8_t*)wgt, base + L.w8,
            (float*)(base + L.ws), (float*)(base + L.wm), F, L.Fpad, nb, nbp, act_blocks);
    }
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = sumi ? smem_optin(reinterpret_cast<const void*>(mmq_kernel<WT, true>), kMmqSmem)
             : smem_optin(reinterpret_cast<const void*>(mmq_kernel<WT, false>), kMmqSmemOut);
    if (e != cudaSuccess) return e;
    MmqParams p;
    p.a8 = base + L.a8; p.as = (const float2*)(base + L.as);
    p.w8 = w8; p.ws = wsp; p.wm = wmp;
    p.C = C; p.sumi = sumi;
    p.T = T; p.F = F; p.nb = nb; p.nkc = L.nkc; p.Tpad = L.Tpad; p.Fpad = L.Fpad;
    p.ldc_t = ldc_t; p.ldc_f = ldc_f;
    p.tiles_m = L.Tpad / kBM; p.tiles_n = L.Fpad / kBN;
    p.dbg = QGEMM_ENV("QGEMM_MMQ_DBG") ? atoi(QGEMM_ENV("QGEMM_MMQ_DBG")) : 0;
    p.peer = peer ? *peer : PeerOut{};
    p.tma_out = 0;
    if (p.peer.world > 1 && !p.peer.mc && !sumi && ldc_t == 1 && T % 4 == 0 && ldc_f % 4 == 0 && !QGEMM_ENV("QGEMM_MMQ_NO_TMA_OUT")) {
        p.tma_out = 1;
        for (int r = 0; r < p.peer.world; r++)
            if (reinterpret_cast<uintptr_t>(p.peer.C[r]) % 16 != 0) p.tma_out = 0;
    }
    const int ntiles = p.tiles_m * p.tiles_n;
    // programmatic dependent launch behind the prepass (the kernel waits for it after its own setup)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(min(ntiles, num_sms));
    cfg.blockDim = dim3(kMmqThreads);
    cfg.dynamicSmemBytes = p.tma_out ? kMmqSmemOut : kMmqSmem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (sumi) e = cudaLaunchKernelEx(&cfg, mmq_kernel<WT, true>, p);
    else e = cudaLaunchKernelEx(&cfg, mmq_kernel<WT, false>, p);
    note_launch();
    return e;
}

cudaError_t launch_quantize_q8_1_tiles(const float* x, uint8_t* a8, float2* as, int T, int Tpad, int K, float coef, uint32_t flags,
                                       cudaStream_t st, unsigned* zero, int nzero, const float* gate);

// fp32 activations straight into the tensor-core path: quantize_q8_1 writes the operand tiles itself, then the
// native-layout GEMM -- two launches, no block_q8_1 round trip (successor of kernels/gemm/gemm_fused.cuh:76-302).
// Returns cudaErrorNotSupported where that pairing does not apply (the caller then chains the separate kernels).
cudaError_t launch_mmq_f32act(int wtype, const float* act_f32, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                              int64_t ldc_f, uint32_t flags, uint32_t qflags, void* ws, size_t ws_bytes, int num_sms, cudaStream_t st,
                              const float* gate) {
    if ((flags & QGEMM_WEIGHTS_PREPACKED) || K % 128 != 0 || !mmq_native_supported(wtype, wgt, T, F, K) || QGEMM_ENV("QGEMM_MMQ_LEGACY") ||
        QGEMM_ENV("QGEMM_NO_FUSED_QUANT"))
        return cudaErrorNotSupported;
    const MmqWs L = mmq_layout(T, F, K);
    if (ws_bytes < L.w8 || reinterpret_cast<uintptr_t>(ws) % 256 != 0) return cudaErrorNotSupported;
    float coef = 0.f;
    if (wtype == QGEMM_TYPE_Q4_0) coef = -8.f;
    else if (wtype == QGEMM_TYPE_Q5_0) coef = -16.f;
    else if (wtype == QGEMM_TYPE_Q4_1 || wtype == QGEMM_TYPE_Q5_1) coef = (flags & QGEMM_MS_EXACT) ? 1.f : 0.25f;
    uint8_t* base = (uint8_t*)ws;
    int ncount = 0;
    unsigned* counters = mmq_native_split_counters(T, F, K, flags, num_sms, ws_bytes > L.w8 ? base + L.w8 : nullptr,
                                                   ws_bytes > L.w8 ? ws_bytes - L.w8 : 0, false, nullptr, &ncount);
    if (cudaError_t e = launch_quantize_q8_1_tiles(act_f32, base + L.a8, (float2*)(base + L.as), T, L.Tpad, K, coef, qflags, st, counters, ncount, gate))
        return e;
    return launch_mmq_native(wtype, base + L.a8, (const float2*)(base + L.as), wgt, C, nullptr, T, F, K, L.Tpad, ldc_t, ldc_f, flags, num_sms,
                             st, nullptr, ws_bytes > L.w8 ? base + L.w8 : nullptr, ws_bytes > L.w8 ? ws_bytes - L.w8 : 0);
}

size_t mmq_prepack_bytes(int wtype, int F, int K) {
    if (block_bytes(wtype) == 0 || wtype == QGEMM_TYPE_Q8_1 || F < 1 || K < 32) return 0;
    return mmq_pack_layout(F, K).total;
}

cudaError_t launch_mmq_prepack(int wtype, const void* wgt, void* packed, int F, int K, cudaStream_t st) {
    const MmqPack P = mmq_pack_layout(F, K);
    const int nb = K / 32, nbp = P.nkc * kBlocksPerStage;
    uint8_t* base = (uint8_t*)packed;
    const int64_t n = (int64_t)P.Fpad * nbp;
    const unsigned grid = (unsigned)((n + 255) / 256);
    const uint8_t* w = (const uint8_t*)wgt;
    float* ws = (float*)(base + P.ws);
    float* wm = (float*)(base + P.wm);
    switch (wtype) {
    case QGEMM_TYPE_Q4_0: mmq_unpack_weight_kernel<QGEMM_TYPE_Q4_0><<<grid, 256, 0, st>>>(w, base + P.w8, ws, wm, F, P.Fpad, nb, nbp); break;
    case QGEMM_TYPE_Q4_1: mmq_unpack_weight_kernel<QGEMM_TYPE_Q4_1><<<grid, 256, 0, st>>>(w, base + P.w8, ws, wm, F, P.Fpad, nb, nbp); break;
    case QGEMM_TYPE_Q5_0: mmq_unpack_weight_kernel<QGEMM_TYPE_Q5_0><<<grid, 256, 0, st>>>(w, base + P.w8, ws, wm, F, P.Fpad, nb, nbp); break;
    case QGEMM_TYPE_Q5_1: mmq_unpack_weight_kernel<QGEMM_TYPE_Q5_1><<<grid, 256, 0, st>>>(w, base + P.w8, ws, wm, F, P.Fpad, nb, nbp); break;
    case QGEMM_TYPE_Q8_0: mmq_unpack_weight_kernel<QGEMM_TYPE_Q8_0><<<grid, 256, 0, st>>>(w, base + P.w8, ws, wm, F, P.Fpad, nb, nbp); break;
    default: return cudaErrorInvalidValue;
    }
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_mmq(int wtype, const void* act, const void* wgt, float* C, int32_t* sumi_out, int T, int F, int K,
                       int64_t ldc_t, int64_t ldc_f, uint32_t flags, void* ws, size_t ws_bytes, int num_sms,
                       cudaStream_t st, const PeerOut* peer) {
    if (ws_bytes < mmq_workspace_need(wtype, wgt, T, F, K, flags) || reinterpret_cast<uintptr_t>(ws) % 256 != 0)
        return cudaErrorInvalidValue;
    switch (wtype) {
    case QGEMM_TYPE_Q4_0: return launch_mmq_t<QGEMM_TYPE_Q4_0>(act, wgt, C, sumi_out, T, F, K, ldc_t, ldc_f, flags, ws, ws_bytes, num_sms, st, peer);
    case QGEMM_TYPE_Q4_1: return launch_mmq_t<QGEMM_TYPE_Q4_1>(act, wgt, C, sumi_out, T, F, K, ldc_t, ldc_f, flags, ws, ws_bytes, num_sms, st, peer);
    case QGEMM_TYPE_Q5_0: return launch_mmq_t<QGEMM_TYPE_Q5_0>(act, wgt, C, sumi_out, T, F, K, ldc_t, ldc_f, flags, ws, ws_bytes, num_sms, st, peer);
    case QGEMM_TYPE_Q5_1: return launch_mmq_t<QGEMM_TYPE_Q5_1>(act, wgt, C, sumi_out, T, F, K, ldc_t, ldc_f, flags, ws, ws_bytes, num_sms, st, peer);
    case QGEMM_TYPE_Q8_0: return launch_mmq_t<QGEMM_TYPE_Q8_0>(act, wgt, C, sumi_out, T, F, K, ldc_t, ldc_f, flags, ws, ws_bytes, num_sms, st, peer);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace qgemm
