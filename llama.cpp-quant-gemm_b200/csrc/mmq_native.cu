// mmq_native.cu -- prefill path (T >= 96), weights read in their NATIVE llama.cpp block layout.
//
// Same contraction as mmq.cu (one tcgen05.mma kind::i8 per quantization block, s32 block sums in TMEM, per-block scale
// fold on the CUDA cores with the reference GPU kernel's FMA sequence, kernels/gemm/gemm_quant_formats.cuh:73-334), but the
// 4/5/8-bit weight blocks are unpacked to the int8 operand tile INSIDE the kernel, in shared memory:
//
//   warps 0-15 epilogue: tcgen05.ld the s32 tiles (lane = token, 32 columns per thread) and fold them into fp32
//              register accumulators in block order
//   warps 16-17 unpack: per raw stage ONE 2-D TMA tensor load brings the next 8 raw blocks of the tile's 128 weight rows
//              (box 128 x 144..272 bytes of the native matrix) into a raw ring; every thread then expands two rows of it
//              -- nibbles / qh bits -> u8 (s8 for q8_0) -- straight into the 128-byte-swizzled K-major operand tile the
//              UMMA descriptor expects, and drops d_w (m_w) into the scale slab of the stage.  No per-call weight
//              prepass, no 2x workspace; the weight side of the L2 -> SM traffic shrinks from 1 byte to the native
//              0.56 .. 1.06 bytes per element.
//   warp 18    one elected thread issues the MMAs; TMEM is two halves of two block buffers each (2 x 2 x 128 columns),
//              one commit per half
//   warp 19    producer of the activation side: pre-swizzled s8 tile + (d_a, c_a) slab per operand stage (the activation
//              prepass of mmq.cu; activations are reused by every weight tile, weights only by the token tiles)
// The light, latency-critical roles sit in the highest warp ids: the issue arbiter favours them over the 16 math warps.
//
// Needs K % 256 == 0 (a raw stage is 8 blocks so that every format's row chunk is a 16-byte multiple) and a 16-byte
// aligned weight base; other shapes take mmq.cu's prepass kernel.
#include <algorithm>
#include <cstdlib>

#include <cuda.h>   // CUtensorMap + enums only; the encoder is looked up through the runtime, libcuda is not linked

#include "tc05.cuh"

#ifdef QGEMM_MMQ_PROFILE
#include <cstdio>
#define PROF_DECL long long pf_wait = 0, pf_wait2 = 0, pf_t0 = clock64()
#define PROF_WAIT(acc, stmt) do { const long long c0_ = clock64(); stmt; acc += clock64() - c0_; } while (0)
#define PROF_STAMP(i) do { if (threadIdx.x == 0) pf_ts[i] = clock64(); } while (0)
#else
#define PROF_DECL
#define PROF_WAIT(acc, stmt) stmt
#define PROF_STAMP(i)
#endif

namespace qgemm {
#ifndef QGEMM_NAT_EARLY_WEIGHTS
#define QGEMM_NAT_EARLY_WEIGHTS 1
#endif

namespace nat {

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

constexpr int kBM = 128;            // tokens per tile = TMEM lanes
constexpr int kBN = 128;            // weight rows per tile = TMEM columns per block buffer
constexpr int kKC = 128;            // K elements per operand stage: 4 blocks, one 128-byte swizzle row
constexpr int kBPS = 4;             // blocks per operand stage
constexpr int kMaxStages = 3;       // operand ring
constexpr int kMaxRaw = 3;          // raw ring (stage = 8 blocks of every row of the tile)
constexpr int kRawBlocks = 8;
constexpr int kEpiWarps = 16;       // warps 0-15: 4 per TMEM lane quarter, 32 columns each
constexpr int kEpiCols = kBN / (kEpiWarps / 4);
constexpr int kUnpackWarps = 2;     // warps 16-17: 64 threads, two weight rows each
constexpr int kUT = kUnpackWarps * 32;
constexpr int kRowsPerThread = kBN / kUT;
constexpr int kWarpUnpack = kEpiWarps, kWarpMma = kWarpUnpack + kUnpackWarps, kWarpProd = kWarpMma + 1;
constexpr int kThreads = (kWarpProd + 1) * 32;
constexpr int kTmemCols = 512;

// operand stage (bytes); tiles 1024-byte aligned for SWIZZLE_128B
constexpr int kStageA = 0;                          // [128 tokens][128 B] s8
constexpr int kStageW = kStageA + kBM * kKC;        // [128 rows][128 B] u8 / s8
constexpr int kStageAS = kStageW + kBN * kKC;       // [4 blocks][128 tokens] float2 (d_a, c_a)
constexpr int kStageWS = kStageAS + kBPS * kBM * 8; // [4 blocks][128 rows] float d_w
constexpr int kStageWM = kStageWS + kBPS * kBN * 4; // [4 blocks][128 rows] float m_w
constexpr int kStageBytes = kStageWM + kBPS * kBN * 4;
static_assert(kStageBytes % 1024 == 0, "stage must keep 1024-byte alignment");
constexpr int kBarBytes = 1024;    // barriers in front of the 1024-byte aligned stages
constexpr int kOutTileBytes = kBM * kBN * 4;

template <int WT> constexpr int raw_row_bytes() { return kRawBlocks * Fmt<WT>::bytes; }
template <int WT> constexpr int raw_stage_bytes() { return kBN * raw_row_bytes<WT>(); }

struct Params {
    const uint8_t* a8;     // [nkc][Tpad][128] swizzled s8 (activation prepass)
    const float2* as;      // [Tpad / 128][nbp][128] (d_a, c_a)
    float* C;
    int32_t* sumi;         // non-null: dump the raw s32 block sums instead of folding
    int T, F, nb, nkc, Tpad;
    int64_t ldc_t, ldc_f;
    int tiles_m, tiles_n;
    int nfull;             // tiles 0 .. nfull-1 are one work unit each, dealt round robin; the raw stages of the others
    int sk_ctas;           // are shared by CTAs 0 .. sk_ctas-1 in equal contiguous ranges, cut at tile boundaries
    float* partial;        // sk_ctas > 0: [2][cta][128 tokens][128 rows] partial sums of a CTA's first / second cut unit
    unsigned* tile_count;  //              [cut tile][2]: tickets taken, partial sums published (both return to zero)
    int stages, raw_stages;
    int dbg;
    PeerOut peer;
    int tma_out;
};


// 2-D tensor load (TMA): box of the weight matrix viewed as [F][row bytes / 2] uint16 -> dense [128][box bytes] in smem
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            ptx::smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(ptx::smem_u32(bar))
        : "memory");
}

// ---- compile-time field extraction from a block's byte stream held as 32-bit words (word 0 = bytes 0..3) ----
template <int OFF, int NW>
__device__ __forceinline__ uint32_t word_at(const uint32_t (&y)[NW]) {
    static_assert(OFF % 2 == 0 && OFF / 4 < NW, "field outside the block");
    if constexpr (OFF % 4 == 0) {
        return y[OFF / 4];
    } else {
        static_assert(OFF / 4 + 1 < NW, "misaligned word crosses the end of the run");
        return __funnelshift_r(y[OFF / 4], y[OFF / 4 + 1], 16);
    }
}
template <int OFF, int NW>
__device__ __forceinline__ float half_at(const uint32_t (&y)[NW]) {
    return half_bits_to_float(y[OFF / 4] >> ((OFF % 4) * 8));
}

// One weight block (J-th of the 4 of an operand stage) of row `row` -> two 16-byte chunks of the swizzled operand row
// + the block's scale(s) in the slab.  y = the row's 4 blocks as 32-bit words (word 0 = bytes 0..3).
template <int WT, int J, int NW>
__device__ __forceinline__ void unpack_one(const uint32_t (&y)[NW], uint8_t* tile_row, int sw, float* ws, float* wm, int row) {
    using Fm = Fmt<WT>;
    constexpr int B = J * Fm::bytes;
    uint32_t w[8];
    if constexpr (Fm::bits == 8) {
        w[0] = word_at<B + 2, NW>(y);  w[1] = word_at<B + 6, NW>(y);  w[2] = word_at<B + 10, NW>(y); w[3] = word_at<B + 14, NW>(y);
        w[4] = word_at<B + 18, NW>(y); w[5] = word_at<B + 22, NW>(y); w[6] = word_at<B + 26, NW>(y); w[7] = word_at<B + 30, NW>(y);
    } else {
        const uint32_t q0 = word_at<B + Fm::qs, NW>(y), q1 = word_at<B + Fm::qs + 4, NW>(y);
        const uint32_t q2 = word_at<B + Fm::qs + 8, NW>(y), q3 = word_at<B + Fm::qs + 12, NW>(y);
        w[0] = q0 & 0x0f0f0f0fu; w[1] = q1 & 0x0f0f0f0fu; w[2] = q2 & 0x0f0f0f0fu; w[3] = q3 & 0x0f0f0f0fu;
        w[4] = (q0 >> 4) & 0x0f0f0f0fu; w[5] = (q1 >> 4) & 0x0f0f0f0fu; w[6] = (q2 >> 4) & 0x0f0f0f0fu; w[7] = (q3 >> 4) & 0x0f0f0f0fu;
        if constexpr (Fm::bits == 5) {
            const uint32_t qh = word_at<B + (Fm::qh >= 0 ? Fm::qh : 0), NW>(y);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                w[i] |= spread_qh4(qh, 4 * i);
                w[i + 4] |= spread_qh4(qh, 16 + 4 * i);
            }
        }
    }
    *reinterpret_cast<uint4*>(tile_row + (((2 * J) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(tile_row + (((2 * J + 1) ^ sw) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
    ws[J * kBN + row] = half_at<B, NW>(y);
    if constexpr (Fm::m >= 0) wm[J * kBN + row] = half_at<B + (Fm::m >= 0 ? Fm::m : 0), NW>(y);
}

// The 4 blocks of one operand stage of weight row `row`: raw bytes (shared memory, 8-byte aligned) -> operand tile row.
template <int WT>
__device__ __forceinline__ void unpack_half_row(const uint8_t* raw_half, uint8_t* stage, int row) {
    constexpr int kHalf = kBPS * Fmt<WT>::bytes;   // 72 / 80 / 88 / 96 / 136 bytes
    static_assert(kHalf % 8 == 0, "half rows are read as 8-byte words");
    constexpr int NW = kHalf / 4;
    uint32_t y[NW];
    const uint2* src = reinterpret_cast<const uint2*>(raw_half);
#pragma unroll
    for (int i = 0; i < NW / 2; i++) {
        const uint2 v = src[i];
        y[2 * i] = v.x;
        y[2 * i + 1] = v.y;
    }
    uint8_t* tile_row = stage + kStageW + row * kKC;
    float* ws = reinterpret_cast<float*>(stage + kStageWS);
    float* wm = reinterpret_cast<float*>(stage + kStageWM);
    const int sw = row & 7;
    unpack_one<WT, 0, NW>(y, tile_row, sw, ws, wm, row);
    unpack_one<WT, 1, NW>(y, tile_row, sw, ws, wm, row);
    unpack_one<WT, 2, NW>(y, tile_row, sw, ws, wm, row);
    unpack_one<WT, 3, NW>(y, tile_row, sw, ws, wm, row);
}

// kTokN = 0: tokens on the M side of the MMA (TMEM lanes), 128 weight rows on N (columns).  kTokN = 32 / 64 (small
// batches): the operands change places -- 128 weight rows on M, kTokN tokens on N -- so the fold, which bounds the kernel,
// shrinks with the batch instead of being paid for 128 padded tokens; a thread then owns one weight row and kTokN / 4
// tokens, d_w / m_w are per lane and (d_a, c_a) per column (the prepass writes them pair-wise for that, Params::as).
constexpr int kSwapUnpackWarps = 2;   // four were measured: no gain (the unpack is 85 % busy with two, but not the limiter)
constexpr int kSwapThreads = (kEpiWarps + kSwapUnpackWarps + 2) * 32;
template <int WT, bool kDump, bool kRefSeq, int kTokN = 0>
__global__ void __launch_bounds__(kTokN ? kSwapThreads : kThreads, 1) mmq_native_kernel(const Params p, const __grid_constant__ CUtensorMap wmap) {
    // warp roles: 16 epilogue warps, 2 unpack warps in both forms.  (Weight-major with kTokN / 8 epilogue warps of 32 token
    // columns each and 4 unpack warps was measured and is slower -- 32 x 11008 x 4096: 44 vs 34 us: one epilogue warp per
    // sub-partition cannot hide the TMEM-load and barrier latencies.)
    constexpr int kEW = kEpiWarps;
    constexpr int kUW = kTokN ? kSwapUnpackWarps : kUnpackWarps;
    constexpr int kWU = kEW, kWM = kWU + kUW, kWP = kWM + 1;
    constexpr int kUTl = kUW * 32, kRPT = kBN / kUTl;
    if (kTokN && threadIdx.x >= (kWP + 1) * 32) return;   // launched with exactly (kWP + 1) warps; a guard, not a path
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    // [ barriers | operand stages | raw ring | (peer staging tile) ]: every offset below is a compile-time constant
    // except the two ring bases
    const int nstages = p.stages, nraw = p.raw_stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);       // [kMaxStages] producer tx + 2 unpack warps
    uint64_t* empty = full + kMaxStages;                      // [kMaxStages] MMA commit + 16 epilogue warps
    uint64_t* rawfull = empty + kMaxStages;                   // [kMaxRaw]    tx of the tensor load
    // TMEM slots of two blocks each.  (Four slots in the weight-major form, whose blocks are only kTokN columns wide, were
    // measured and change nothing: 32 x 11008 x 4096 29.3 us either way -- the operand-stage round trip, not the TMEM
    // hand-off, is its floor.)
    constexpr int kSlots = 2;
    constexpr int kColStep = kTokN ? kTokN : kBN;             // TMEM columns between consecutive block buffers
    static_assert(2 * kSlots * kColStep <= kTmemCols, "TMEM block buffers");
    uint64_t* tfull = rawfull + kMaxRaw;                      // [kSlots]     MMA commit per TMEM slot
    uint64_t* tempty = tfull + 4;                             // [kSlots]     16 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 4);
    volatile int* split_info = reinterpret_cast<volatile int*>(tmem_slot + 1);   // [3]: this CTA completes the cut tile; its first / last CTA
    uint8_t* stages = smem + kBarBytes;
    uint8_t* raw_ring = stages + nstages * kStageBytes;
    float* out_tile = reinterpret_cast<float*>(raw_ring + nraw * raw_stage_bytes<WT>());   // [kBN][kBM], only with p.tma_out

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkc = p.nkc;
    const int nbp = nkc * kBPS;
    const int ntiles = p.tiles_m * p.tiles_n;
    // Work units of this CTA: whole tiles blockIdx.x, blockIdx.x + grid, ... below nfull, then its share of the cut tiles:
    // a contiguous range of their raw stages, i.e. one segment of a tile or (when the range crosses a tile boundary) the
    // end of one tile and the start of the next.  With few tiles (T <= 256 at Llama widths) every tile is cut so that the
    // whole chip works on the call; with many, the tiles of the ragged last wave are.  A cut tile's segments are added in
    // K order by the CTA that arrives last, so the result does not depend on the arrival order.
    const int nrs = nkc >> 1;                   // raw stages (two operand stages) per tile
    const int nfull = p.nfull;
    const int sk_total = (ntiles - nfull) * nrs;
    auto sk_begin = [&](int c) { return (int)((long long)c * sk_total / p.sk_ctas); };   // first raw stage of CTA c's share
    const int sk_w0 = (int)blockIdx.x < p.sk_ctas ? sk_begin(blockIdx.x) : 0;
    const int sk_w1 = (int)blockIdx.x < p.sk_ctas ? sk_begin(blockIdx.x + 1) : 0;
    struct Cursor { int full, w; };
    const Cursor cur0{(int)blockIdx.x, sk_w0};
    // next unit: tile, first raw stage, raw stages, cut unit number (-1: a whole tile, stored directly; else 0 / 1)
    auto next_unit = [&](Cursor& cu, int& tile, int& rs0, int& n, int& seg) -> bool {
        if (cu.full < nfull) { tile = cu.full; rs0 = 0; n = nrs; seg = -1; cu.full += gridDim.x; return true; }
        if (cu.w < sk_w1) {
            const int t = cu.w / nrs;
            seg = cu.w == sk_w0 ? 0 : 1;
            tile = nfull + t; rs0 = cu.w - t * nrs; n = min(nrs - rs0, sk_w1 - cu.w);
            cu.w += n;
            return true;
        }
        return false;
    };

    if (threadIdx.x == kWP * 32) {
        for (int s = 0; s < kMaxStages; s++) {
            ptx::mbar_init(&full[s], 1 + kUW);
            ptx::mbar_init(&empty[s], 1 + kEW);
        }
        for (int r = 0; r < kMaxRaw; r++) ptx::mbar_init(&rawfull[r], 1);
        for (int h = 0; h < kSlots; h++) {
            ptx::mbar_init(&tfull[h], 1);
            ptx::mbar_init(&tempty[h], kEW);
        }
        ptx::fence_mbar_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
    }
    if (warp == kWM) t5::alloc(tmem_slot, kTmemCols);
    t5::fence_before();
    __syncthreads();
    t5::fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Launched behind the activation prepass with programmatic serialization.  Only what the prepass produces (operand
    // tiles, scale slabs, cleared split counters) has to wait for it: the unpack warps start streaming and unpacking the
    // weights at once (whatever wrote the weights ran before the prepass, which itself waited for it).
    if (warp < kWU || warp >= kWU + kUW || QGEMM_NAT_EARLY_WEIGHTS == 0) ptx::griddep_wait();

    if (warp == kWP) {
        // ===================== activation producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            PROF_DECL;
            Cursor cu = cur0;
            int tile, rs0, n, seg;
            while (next_unit(cu, tile, rs0, n, seg)) {
                const int mt = tile % p.tiles_m;
                for (int kc = 2 * rs0; kc < 2 * (rs0 + n); kc++) {
                    PROF_WAIT(pf_wait, ptx::mbar_wait_backoff_guarded(&empty[s], ph ^ 1));
                    uint8_t* st = stages + s * kStageBytes;
                    ptx::mbar_arrive_expect_tx(&full[s], kBM * kKC + kBPS * kBM * 8);
                    ptx::bulk_g2s(st + kStageA, p.a8 + ((size_t)kc * p.Tpad + (size_t)mt * kBM) * kKC, kBM * kKC, &full[s]);
                    ptx::bulk_g2s(st + kStageAS, p.as + ((size_t)mt * nbp + (size_t)kc * kBPS) * kBM, kBPS * kBM * 8, &full[s]);
                    if (++s == nstages) { s = 0; ph ^= 1; }
                }
            }
#ifdef QGEMM_MMQ_PROFILE
            if (blockIdx.x == 0 && (p.dbg & 32)) printf("producer: total %lld, waiting for empty %lld\n", clock64() - pf_t0, pf_wait);
#endif
        }
    } else if (warp == kWM) {
        // ===================== MMA issuer =====================
        // D[token, row] (s32) = A (s8 activations, M side) . B (u8 / s8 weights, N side)^T, both K-major
        constexpr uint32_t wfmt = Fmt<WT>::bits == 8 ? 1u : 0u;   // s8 (q8_0) or u8
        constexpr uint32_t idesc = kTokN == 0 ? (2u << 4) | (1u << 7) | (wfmt << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24)
                                              : (2u << 4) | (wfmt << 7) | (1u << 10) | ((uint32_t)(kTokN >> 3) << 17) | ((uint32_t)(kBN >> 4) << 24);
        int s = 0;
        uint32_t ph = 0, tq = 0;   // tq: TMEM slot uses so far (slot = tq % kSlots, parity = its use count)
        PROF_DECL;
        Cursor cu = cur0;
        int tile, rs0, n, seg;
        while (next_unit(cu, tile, rs0, n, seg)) {
            for (int kc = 0; kc < 2 * n; kc++) {
                PROF_WAIT(pf_wait, ptx::mbar_wait_guarded(&full[s], ph));
                t5::fence_after();
                // kTokN: the weight tile is the M-side operand, the first kTokN rows of the activation tile the N side
                const uint64_t adesc = t5::smem_desc(ptx::smem_u32(stages + s * kStageBytes + (kTokN ? kStageW : kStageA)));
                const uint64_t bdesc = t5::smem_desc(ptx::smem_u32(stages + s * kStageBytes + (kTokN ? kStageA : kStageW)));
#pragma unroll
                for (int h = 0; h < 2; h++, tq++) {
                    const uint32_t q = tq % kSlots, qph = (tq / kSlots) & 1u;
                    PROF_WAIT(pf_wait2, ptx::mbar_wait_guarded(&tempty[q], qph ^ 1));
                    t5::fence_after();
                    if (lane == 0) {
                        // one instruction = one quantization block (32 bytes of K = +2 in the >>4 address field)
                        t5::mma_i8(tmem_base + (2 * q) * kColStep, adesc + 2 * (2 * h), bdesc + 2 * (2 * h), idesc, 0u);
                        t5::mma_i8(tmem_base + (2 * q + 1) * kColStep, adesc + 2 * (2 * h + 1), bdesc + 2 * (2 * h + 1), idesc, 0u);
                        t5::commit(&tfull[q]);
                    }
                    __syncwarp();
                }
                if (lane == 0) t5::commit(&empty[s]);   // operand tiles consumed once these MMAs retire
                __syncwarp();
                if (++s == nstages) { s = 0; ph ^= 1; }
            }
        }
#ifdef QGEMM_MMQ_PROFILE
        if (blockIdx.x == 0 && lane == 0 && (p.dbg & 32)) printf("mma: total %lld, waiting for full %lld, for tempty %lld\n", clock64() - pf_t0, pf_wait, pf_wait2);
#endif
    } else if (warp >= kWU) {
        // ===================== weight unpack =====================
        const int u = threadIdx.x - kWU * 32;   // this thread owns weight rows u, u + 64 of the tile
        constexpr int kRow = raw_row_bytes<WT>();
        constexpr int kHalf = kBPS * Fmt<WT>::bytes;
        int total = 0;                                    // raw stages of this CTA, units back to back
        {
            Cursor cu = cur0;
            int tile, rs0, n, seg;
            while (next_unit(cu, tile, rs0, n, seg)) total += n;
        }
        // issue side: the unit being requested, the next raw stage inside it, and the ring slot it goes to
        Cursor icu = cur0;
        int itile = 0, irs0 = 0, inrs = 0, islotp = 0, irs = 0, islot = 0, issued = 0;
        next_unit(icu, itile, irs0, inrs, islotp);
        auto issue = [&]() {   // one 2-D tensor load: the next 8 raw blocks of the tile's 128 rows (rows >= F arrive as zeros)
            if (u == 0) {
                ptx::mbar_arrive_expect_tx(&rawfull[islot], (uint32_t)raw_stage_bytes<WT>());
                tma_load_2d(raw_ring + islot * raw_stage_bytes<WT>(), &wmap, (irs0 + irs) * (kRow / 2), (itile / p.tiles_m) * kBN, &rawfull[islot]);
            }
            if (++irs == inrs) { irs = 0; next_unit(icu, itile, irs0, inrs, islotp); }
            if (++islot == nraw) islot = 0;
            issued++;
        };
        while (issued < nraw && issued < total) issue();
        int s = 0, r = 0;
        uint32_t ph = 0, rph = 0;
        PROF_DECL;
        for (int q = 0; q < total; q++) {
            PROF_WAIT(pf_wait, ptx::mbar_wait_guarded(&rawfull[r], rph));
            const uint8_t* rslot = raw_ring + r * raw_stage_bytes<WT>();
#pragma unroll
            for (int h = 0; h < 2; h++) {
                PROF_WAIT(pf_wait2, ptx::mbar_wait_guarded(&empty[s], ph ^ 1));
#pragma unroll
                for (int i = 0; i < kRPT; i++) {
                    const int row = u + i * kUTl;
                    unpack_half_row<WT>(rslot + row * kRow + h * kHalf, stages + s * kStageBytes, row);
                }
                ptx::fence_proxy_async();                 // generic-proxy stores -> visible to the tensor core's reads
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&full[s]);
                if (++s == nstages) { s = 0; ph ^= 1; }
            }
            if (++r == nraw) { r = 0; rph ^= 1; }
            if (issued < total) {
                ptx::bar_sync(1, kUTl);                    // both warps are done reading the slot
                issue();
            }
        }
#ifdef QGEMM_MMQ_PROFILE
        if (blockIdx.x == 0 && u == 0 && (p.dbg & 32)) printf("unpack: total %lld, waiting for raw %lld, for empty %lld\n", clock64() - pf_t0, pf_wait, pf_wait2);
#endif
    } else {
        // ===================== epilogue =====================
        const int ew = warp;
        const int quarter = warp & 3;           // TMEM lane quarter this warp may touch
        const int cgrp = ew >> 2;               // which column group
        const int row = quarter * 32 + lane;    // TMEM lane = token row inside the tile (kTokN: weight row)
        constexpr int NC = kTokN ? kTokN / 4 : kEpiCols;   // columns per thread
        const uint32_t tm = tmem_base + ((uint32_t)(quarter * 32) << 16) + cgrp * NC;
        static_assert(kEpiCols == 32 && (NC == 32 || NC == 16 || NC == 8), "one tcgen05.ld per block per thread");
        static_assert(!(kDump && kTokN), "the block-sum dump uses the token-major form");
        int s = 0;
        uint32_t ph = 0, tq = 0;
        PROF_DECL;
        Cursor cu = cur0;
        int tile, rs0, n, seg;
        while (next_unit(cu, tile, rs0, n, seg)) {
            const int mt = tile % p.tiles_m, nt = tile / p.tiles_m;
#ifdef QGEMM_MMQ_PROFILE
            long long pf_ts[6] = {0, 0, 0, 0, 0, 0};
            PROF_STAMP(0);
#endif
            uint64_t acc[NC / 2];  // fp32 accumulators as packed pairs (columns 2i, 2i+1)
#pragma unroll
            for (int i = 0; i < NC / 2; i++) acc[i] = 0ull;
#pragma unroll 1
            for (int kc = 2 * rs0; kc < 2 * (rs0 + n); kc++) {
                PROF_WAIT(pf_wait, ptx::mbar_wait_guarded(&full[s], ph));  // scale slabs of this stage are visible
                const uint8_t* st = stages + s * kStageBytes;
#pragma unroll 1
                for (int h = 0; h < 2; h++, tq++) {   // one TMEM slot = two blocks per iteration; not unrolled further so that
                                                      // the accumulators keep their registers
                    const uint32_t q = tq % kSlots, qph = (tq / kSlots) & 1u;
                    PROF_WAIT(pf_wait2, ptx::mbar_wait_guarded(&tfull[q], qph));
                    t5::fence_after();
#pragma unroll
                    for (int jj = 0; jj < 2; jj++) {
                        const int j = 2 * h + jj;
                        int x[NC];
                        t5::ldn<NC>(tm + (2 * q + jj) * kColStep, x);
                        t5::wait_ld();
                        if (jj == 1) {   // both blocks of this slot are in registers: hand it back to the tensor core
                            t5::fence_before();
                            __syncwarp();
                            if (lane == 0) ptx::mbar_arrive(&tempty[q]);
                        }
                        if constexpr (kDump) {
                            const int t = mt * kBM + row, b = kc * kBPS + j;
                            if (t < p.T && b < p.nb) {
#pragma unroll
                                for (int i = 0; i < kEpiCols; i++) {
                                    const int f = nt * kBN + cgrp * kEpiCols + i;
                                    if (f < p.F) p.sumi[((size_t)t * p.F + f) * p.nb + b] = x[i];
                                }
                            }
                        } else if constexpr (kTokN > 0) {
                            // lane = weight row: its d_w (m_w) of this block; columns = tokens: (d_a, d_a'), (c_a, c_a') per pair
                            const float dwf = reinterpret_cast<const float*>(st + kStageWS)[j * kBN + row];
                            const uint64_t dw = pk(dwf, dwf);
                            uint64_t mw = 0ull;
                            if constexpr (Fmt<WT>::m >= 0) {
                                const float mwf = reinterpret_cast<const float*>(st + kStageWM)[j * kBN + row];
                                mw = pk(mwf, mwf);
                            }
                            const ulonglong2* sc = reinterpret_cast<const ulonglong2*>(st + kStageAS) + (j * kBM + cgrp * NC) / 2;
#pragma unroll
                            for (int i = 0; i < NC / 2; i++) {
                                const ulonglong2 a = sc[i];   // broadcast read
                                acc[i] = fold_pair<WT, kRefSeq>(acc[i], x[2 * i], x[2 * i + 1], dw, mw, a.x, a.y);
                            }
                        } else {
                            const float2 a = reinterpret_cast<const float2*>(st + kStageAS)[j * kBM + row];
                            const uint64_t da = pk(a.x, a.x), ca = pk(a.y, a.y);
                            const ulonglong2* dw2 = reinterpret_cast<const ulonglong2*>(st + kStageWS) + (j * kBN + cgrp * kEpiCols) / 4;
                            const ulonglong2* mw2 = reinterpret_cast<const ulonglong2*>(st + kStageWM) + (j * kBN + cgrp * kEpiCols) / 4;
#pragma unroll
                            for (int i4 = 0; i4 < kEpiCols / 4; i4++) {
                                const ulonglong2 dw = dw2[i4];  // d_w of columns 4*i4 .. 4*i4+3 (broadcast read)
                                ulonglong2 mw = make_ulonglong2(0ull, 0ull);
                                if constexpr (Fmt<WT>::m >= 0) mw = mw2[i4];
                                acc[2 * i4] = fold_pair<WT, kRefSeq>(acc[2 * i4], x[4 * i4], x[4 * i4 + 1], dw.x, mw.x, da, ca);
                                acc[2 * i4 + 1] = fold_pair<WT, kRefSeq>(acc[2 * i4 + 1], x[4 * i4 + 2], x[4 * i4 + 3], dw.y, mw.y, da, ca);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&empty[s]);  // scale slabs consumed
                if (++s == nstages) { s = 0; ph ^= 1; }
            }
            if constexpr (!kDump) {
                PROF_STAMP(1);
                if (seg >= 0) {
                    // A cut tile.  Every segment takes a ticket; all but the last to arrive put their partial sums into
                    // scratch and publish them; the last one waits for those, adds the segments in K order and stores C.
                    // A CTA's final unit keeps its own sums on chip while it finds out whether it is last; a unit with
                    // another one behind it (the end of a tile whose successor the CTA also works on) publishes first.
                    const int ct = tile - nfull, x0 = ct * nrs;
                    const bool final_unit = cu.w >= sk_w1;
                    // scratch layout of a partial tile: float4 (columns 4j .. 4j+3 of token `row`) at [j][row], so that the
                    // lanes of a warp (consecutive tokens) touch consecutive 16-byte words
                    const size_t off = (size_t)(cgrp * (NC / 4)) * kBM + row;
                    float4* part = reinterpret_cast<float4*>(p.partial + ((size_t)seg * p.sk_ctas + blockIdx.x) * (kBM * kBN)) + off;
                    auto publish = [&]() {
#pragma unroll
                        for (int i = 0; i < NC / 4; i++) {
                            float4 v;
                            unpk(acc[2 * i], v.x, v.y);
                            unpk(acc[2 * i + 1], v.z, v.w);
                            __stcg(part + i * kBM, v);
                        }
                        ptx::bar_sync(3, kEW * 32);
                        if (threadIdx.x == 0) {
                            __threadfence();   // cumulative: covers the stores the barrier has ordered before this thread
                            atomicAdd(p.tile_count + 2 * ct + 1, 1u);
                        }
                    };
                    if (!final_unit) publish();
                    if (threadIdx.x == 0) {
                        // CTAs c0 .. c1 hold the tile's raw stages [x0, x0 + nrs)
                        int c0 = (int)((long long)x0 * p.sk_ctas / sk_total);
                        while (c0 + 1 < p.sk_ctas && sk_begin(c0 + 1) <= x0) c0++;
                        int c1 = (int)((long long)(x0 + nrs - 1) * p.sk_ctas / sk_total);
                        while (c1 + 1 < p.sk_ctas && sk_begin(c1 + 1) <= x0 + nrs - 1) c1++;
                        const bool last = atomicAdd(p.tile_count + 2 * ct, 1u) == (unsigned)(c1 - c0);
                        if (last) {
                            // everybody else has a ticket, so their sums are on their way
                            const unsigned want = (unsigned)(c1 - c0) + (final_unit ? 0u : 1u);
                            unsigned polls = 0;
                            while (ld_acquire_gpu(p.tile_count + 2 * ct + 1) < want) {
                                __nanosleep(64);
                                if (++polls > (1u << 24)) asm volatile("trap;");
                            }
                            p.tile_count[2 * ct] = 0u;       // ready for the next call
                            p.tile_count[2 * ct + 1] = 0u;
                        }
                        split_info[0] = last; split_info[1] = c0; split_info[2] = c1;
                    }
                    ptx::bar_sync(3, kEW * 32);
                    const bool last = split_info[0] != 0;
                    const int c0 = split_info[1], c1 = split_info[2];
                    ptx::bar_sync(3, kEW * 32);            // the words may be rewritten by the next unit
                    PROF_STAMP(2);
                    if (!last) {
                        if (final_unit) publish();
                        PROF_STAMP(3);
#ifdef QGEMM_MMQ_PROFILE
                        if (threadIdx.x == 0 && (p.dbg & 64) && blockIdx.x % 37 == 0)
                            printf("cta %d tile %d rs0 %d n %d: loop %lld ticket %lld publish %lld (not last, final %d)\n", blockIdx.x, tile, rs0, n,
                                   pf_ts[1] - pf_ts[0], pf_ts[2] - pf_ts[1], pf_ts[3] - pf_ts[2], (int)final_unit);
#endif
                        continue;
                    }
                    PROF_STAMP(3);
                    // own sums of a final unit -> shared memory (the operand stages are idle by now), so that the sum can be
                    // formed in place, in K order, wherever this CTA's segment lies
                    float4* own = reinterpret_cast<float4*>(stages) + threadIdx.x;
                    const int me = final_unit ? (int)blockIdx.x : -1;
                    if (me >= 0 && me != c0) {
#pragma unroll
                        for (int i = 0; i < NC / 4; i++) {
                            float4 v;
                            unpk(acc[2 * i], v.x, v.y);
                            unpk(acc[2 * i + 1], v.z, v.w);
                            own[i * (kEW * 32)] = v;
                        }
                    }
#pragma unroll 1
                    for (int c = (me == c0 ? c0 + 1 : c0); c <= c1; c++) {
                        // CTA c's segment of this tile is its first cut unit unless its share began in the tile before
                        const float4* ps = reinterpret_cast<const float4*>(p.partial + ((size_t)(sk_begin(c) >= x0 ? 0 : 1) * p.sk_ctas + c) * (kBM * kBN)) + off;
                        float4 in[NC / 4];
                        if (c == me) {
#pragma unroll
                            for (int i = 0; i < NC / 4; i++) in[i] = own[i * (kEW * 32)];
                        } else {
#pragma unroll
                            for (int i = 0; i < NC / 4; i++) in[i] = __ldcg(ps + i * kBM);
                        }
#pragma unroll
                        for (int i = 0; i < NC / 4; i++) {
                            const float4 v = in[i];
                            if (c == c0) {
                                acc[2 * i] = pk(v.x, v.y);
                                acc[2 * i + 1] = pk(v.z, v.w);
                            } else {
                                fadd2_acc(acc[2 * i], pk(v.x, v.y));
                                fadd2_acc(acc[2 * i + 1], pk(v.z, v.w));
                            }
                        }
                    }
                }
                PROF_STAMP(4);
                if constexpr (kTokN > 0) {
                    // row f of C, tokens cgrp * NC ..: contiguous along t in the ggml layout (ldc_t = 1)
                    const int f = nt * kBN + row;
                    if (f < p.F && !(p.dbg & 4)) {
                        float* crow = p.C + (int64_t)f * p.ldc_f;
#pragma unroll
                        for (int i = 0; i < NC / 2; i++) {
                            float v0, v1;
                            unpk(acc[i], v0, v1);
                            const int t0 = cgrp * NC + 2 * i;
                            if (t0 < p.T) crow[(int64_t)t0 * p.ldc_t] = v0;
                            if (t0 + 1 < p.T) crow[(int64_t)(t0 + 1) * p.ldc_t] = v1;
                        }
                    }
                } else {
                const int t = mt * kBM + row;
                if (p.tma_out) {
                    // Fused all-gather, bulk variant (see mmq.cu): the tile is staged as [f][t] and carried to every rank's
                    // gathered C by the TMA engine, 512-byte rows, while the epilogue warps go on with the next tile.
                    if (lane == 0) ptx::bulk_wait_read();
                    ptx::bar_sync(3, kEW * 32);
#pragma unroll
                    for (int i = 0; i < kEpiCols / 2; i++) {
                        float v0, v1;
                        unpk(acc[i], v0, v1);
                        out_tile[(cgrp * kEpiCols + 2 * i) * kBM + row] = v0;
                        out_tile[(cgrp * kEpiCols + 2 * i + 1) * kBM + row] = v1;
                    }
                    ptx::fence_proxy_async();
                    ptx::bar_sync(3, kEW * 32);
                    if (lane == 0) {
                        const int t0 = mt * kBM;
                        const uint32_t bytes = (uint32_t)min(kBM, p.T - t0) * 4u;
                        constexpr int kRowsPerWarp = kBN / kEW;
                        const int ndst = p.peer.mc ? 1 : p.peer.world;   // one store to the multicast mapping reaches every rank
#pragma unroll 1
                        for (int q = 0; q < ndst; q++) {
                            int r = p.peer.rank + 1 + q;   // staggered start: the ranks target different receivers
                            if (r >= p.peer.world) r -= p.peer.world;
                            float* Cr = p.peer.mc ? p.peer.mc : p.peer.C[r];
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
                    const int nrank = (p.peer.world > 1 && !p.peer.mc) ? p.peer.world : 1;
#pragma unroll 1
                    for (int q = 0; q < nrank; q++) {
                        int r = p.peer.rank + 1 + q;
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
                }   // token-major store
#ifdef QGEMM_MMQ_PROFILE
                PROF_STAMP(5);
                if (threadIdx.x == 0 && (p.dbg & 64) && blockIdx.x % 37 == 0)
                    printf("cta %d tile %d rs0 %d n %d: loop %lld write+fence %lld sync+count %lld reduce %lld store %lld\n", blockIdx.x, tile, rs0, n,
                           pf_ts[1] - pf_ts[0], pf_ts[2] - pf_ts[1], pf_ts[3] - pf_ts[2], pf_ts[4] - pf_ts[3], pf_ts[5] - pf_ts[4]);
#endif
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
    if (warp == kWM) t5::dealloc(tmem_base, kTmemCols);
    if constexpr (!kDump) {
        if (threadIdx.x == 0) peer_signal_done(p.peer, gridDim.x);  // the barrier above ordered every epilogue store
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// the native weight matrix as a 2-D tensor of uint16 [F][row bytes / 2]; box = 128 rows x the 8 raw blocks of a stage
template <int WT>
static bool make_weight_map(CUtensorMap* map, const void* wgt, int F, int nb) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t rowbytes = (cuuint64_t)nb * Fmt<WT>::bytes;
    const cuuint64_t gdim[2] = {rowbytes / 2, (cuuint64_t)F};
    const cuuint64_t gstride[1] = {rowbytes};
    const cuuint32_t box[2] = {(cuuint32_t)raw_row_bytes<WT>() / 2, (cuuint32_t)kBN};
    const cuuint32_t estride[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(wgt), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int WT>
static cudaError_t launch_t(Params p, const void* wgt, bool refseq, int num_sms, cudaStream_t st, int tokn) {
    CUtensorMap wmap;
    if (!make_weight_map<WT>(&wmap, wgt, p.F, p.nb)) return cudaErrorNotSupported;
    // operand ring 3 deep (2 next to the peer staging tile), raw ring as deep as fits
    p.stages = p.tma_out ? 2 : kMaxStages;
    const size_t fixed = 1024 + (size_t)p.stages * kStageBytes + kBarBytes + (p.tma_out ? kOutTileBytes : 0);
    p.raw_stages = (int)min((size_t)kMaxRaw, (227 * 1024 - fixed) / raw_stage_bytes<WT>());
    if (p.raw_stages < 2) return cudaErrorInvalidValue;
    const size_t smem = fixed + (size_t)p.raw_stages * raw_stage_bytes<WT>();
    // q4_1 / q5_1 have a cheaper fold than the reference's operation sequence (tc05.cuh); refseq keeps the latter
    constexpr bool kHasM = Fmt<WT>::m >= 0;
    const bool fast = kHasM && !refseq;
    using KernelFn = void (*)(const Params, const CUtensorMap);
    KernelFn kfn;
    if (p.sumi) kfn = mmq_native_kernel<WT, true, true>;
    else if (tokn == 32) kfn = fast ? mmq_native_kernel<WT, false, !kHasM, 32> : mmq_native_kernel<WT, false, true, 32>;
    else if (tokn == 64) kfn = fast ? mmq_native_kernel<WT, false, !kHasM, 64> : mmq_native_kernel<WT, false, true, 64>;
    else kfn = fast ? mmq_native_kernel<WT, false, !kHasM> : mmq_native_kernel<WT, false, true>;
    if (cudaError_t e = smem_optin(reinterpret_cast<const void*>(kfn), smem)) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(max(min(p.nfull, num_sms), p.sk_ctas));
    cfg.blockDim = dim3(tokn ? kSwapThreads : kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // behind the activation prepass
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kfn, p, wmap);
    note_launch();
    return e;
}

}  // namespace nat

// K % 256 == 0 and bulk-copyable rows (a multiple of 256 elements makes every format's row a 16-byte multiple)
bool mmq_native_supported(int wtype, const void* wgt, int T, int F, int K) {
    if (block_bytes(wtype) == 0 || wtype == QGEMM_TYPE_Q8_1 || T < 1 || F < 1 || K < 256 || K % 256 != 0) return false;
    return reinterpret_cast<uintptr_t>(wgt) % 16 == 0 && nat::encode_tiled_fn() != nullptr;   // nullptr counts as aligned
}

// Split-K plan of a call.  Unsplit when the caller wants the reference's summation order, or when nothing is gained.
// The whole waves of tiles (nfull, a multiple of the SM count) run as they are, round robin.  The tiles of the ragged last
// wave -- all tiles when there is less than one wave -- are cut along K:
//  * at most half a wave of them: each tile into as many segments as there are SMs per such tile (at least two raw stages
//    = 16 blocks per segment, lengths differing by at most one raw stage), one segment per CTA, so that the wave costs
//    1/segments of a tile time.  Measured (q4_0): 128 x 4096 x 4096 (32 tiles) 62 -> 32 us, 256 x 4096 x 4096 64.5 -> 42 us;
//    896 tiles (config 5 on 8 GPUs, per rank) 0.70 -> 0.63 ms.
//  * more than that: all SMs share the raw stages of those tiles in equal contiguous ranges; a range that crosses a tile
//    boundary makes two units (the end of one tile, the start of the next), each tile is summed from two or three
//    segments, and the wave costs tiles/SMs of a tile time.  Every extra partial tile costs its CTA a few thousand cycles
//    (profiles/r02_split_k.md), so this needs a tile long enough to gain more than that.
struct NatSplit { int nfull, ncut, sk_ctas, slots; };
static NatSplit mmq_native_plan(int T, int F, int K, uint32_t flags, int num_sms) {
    const int tiles = ((T + nat::kBM - 1) / nat::kBM) * ((F + nat::kBN - 1) / nat::kBN), nrs = K / (2 * nat::kKC);
    NatSplit pl{tiles, 0, 0, 0};
    if ((flags & QGEMM_FOLD_REFSEQ) || QGEMM_ENV("QGEMM_MMQ_NO_SPLITK") || num_sms < 1) return pl;
    const int rest = tiles % num_sms;
    if (rest == 0 || (tiles > num_sms && QGEMM_ENV("QGEMM_MMQ_NO_TAILSPLIT"))) return pl;
    const int k = std::min(num_sms / rest, nrs / 2);
    if (k >= 2) return NatSplit{tiles - rest, rest, rest * k, 1};
    // shared ranges: worth it when a CTA saves more raw stages than its second partial tile and the reduction cost (at
    // 2.2 raw stages saved, 512 x 4096 x 4096, it is a draw: 64.5 us either way)
    const int min_gain = QGEMM_ENV("QGEMM_MMQ_SK_MINGAIN") ? atoi(QGEMM_ENV("QGEMM_MMQ_SK_MINGAIN")) : 3;
    if (2 * rest > num_sms && (long long)nrs * (num_sms - rest) >= (long long)min_gain * num_sms)
        return NatSplit{tiles - rest, rest, num_sms, 2};
    return pl;
}
size_t mmq_native_split_bytes(int T, int F, int K, uint32_t flags, int num_sms) {
    const NatSplit pl = mmq_native_plan(T, F, K, flags, num_sms);
    if (pl.sk_ctas == 0) return 0;
    return (size_t)pl.sk_ctas * pl.slots * nat::kBM * nat::kBN * sizeof(float) + ((size_t)pl.ncut * 2 * sizeof(unsigned) + 255) / 256 * 256;
}

// Where the arrival counters of a split-K call live inside its scratch (nullptr / 0: the call runs unsplit).  They must be
// zero when the GEMM kernel starts: the activation prepass in front of it clears them (a memset node between the two
// kernels would break their programmatic dependency and cost ~15 us).
unsigned* mmq_native_split_counters(int T, int F, int K, uint32_t flags, int num_sms, void* split_ws, size_t split_ws_bytes, bool dump,
                                    const PeerOut* peer, int* count) {
    *count = 0;
    if (dump) return nullptr;
    (void)peer;
    const NatSplit pl = mmq_native_plan(T, F, K, flags, num_sms);
    if (pl.sk_ctas == 0 || !split_ws || split_ws_bytes < mmq_native_split_bytes(T, F, K, flags, num_sms) || reinterpret_cast<uintptr_t>(split_ws) % 16 != 0)
        return nullptr;
    *count = 2 * pl.ncut;
    return (unsigned*)((char*)split_ws + (size_t)pl.sk_ctas * pl.slots * nat::kBM * nat::kBN * sizeof(float));
}

// Small batches take the kernel form with the weights on the M side and 32 / 64 tokens on the N side (0: the token-major
// form).  The prepass must then write the (d_a, c_a) slabs pair-wise (see mmq_repack_act_kernel).  Not with peers (their
// epilogue stores are token-major) and not for the block-sum dump.
int mmq_native_tokn(int T, const PeerOut* peer, bool dump, uint32_t flags) {
    (void)flags;
    if (dump || (peer && peer->world > 1) || QGEMM_ENV("QGEMM_MMQ_NO_SWAP")) return 0;
    return T <= 32 ? 32 : T <= 64 ? 64 : 0;
}

// a8 / as: the activation prepass of mmq.cu (Tpad tokens, nkc operand stages); split_ws: mmq_native_split_bytes() bytes whose
// counters (mmq_native_split_counters) the caller has cleared on this stream
cudaError_t launch_mmq_native(int wtype, const uint8_t* a8, const float2* as, const void* wgt, float* C, int32_t* sumi, int T,
                              int F, int K, int Tpad, int64_t ldc_t, int64_t ldc_f, uint32_t flags, int num_sms, cudaStream_t st,
                              const PeerOut* peer, void* split_ws, size_t split_ws_bytes, int tokn) {
    if (tokn != 0 && (tokn != mmq_native_tokn(T, peer, sumi != nullptr, flags))) return cudaErrorInvalidValue;
    nat::Params p;
    p.a8 = a8; p.as = as; p.C = C; p.sumi = sumi;
    p.T = T; p.F = F; p.nb = K / 32; p.nkc = K / nat::kKC; p.Tpad = Tpad;
    p.ldc_t = ldc_t; p.ldc_f = ldc_f;
    p.tiles_m = Tpad / nat::kBM; p.tiles_n = (F + nat::kBN - 1) / nat::kBN;
    p.stages = 0; p.raw_stages = 0;
    p.nfull = p.tiles_m * p.tiles_n; p.sk_ctas = 0; p.partial = nullptr; p.tile_count = nullptr;
    {
        int ncount = 0;
        unsigned* counters = mmq_native_split_counters(T, F, K, flags, num_sms, split_ws, split_ws_bytes, sumi != nullptr, peer, &ncount);
        if (counters) {
            const NatSplit pl = mmq_native_plan(T, F, K, flags, num_sms);
            p.nfull = pl.nfull;
            p.sk_ctas = pl.sk_ctas;
            p.partial = (float*)split_ws;
            p.tile_count = counters;
        }
    }
    const bool refseq = (flags & QGEMM_FOLD_REFSEQ) != 0;
    p.dbg = QGEMM_ENV("QGEMM_MMQ_DBG") ? atoi(QGEMM_ENV("QGEMM_MMQ_DBG")) : 0;
    p.peer = peer ? *peer : PeerOut{};
    p.tma_out = 0;
    if (p.peer.world > 1 && (!p.peer.mc || QGEMM_ENV("QGEMM_MMQ_MC_TMA")) && !sumi && ldc_t == 1 && T % 4 == 0 && ldc_f % 4 == 0 &&
        !QGEMM_ENV("QGEMM_MMQ_NO_TMA_OUT")) {
        p.tma_out = 1;
        for (int r = 0; r < p.peer.world; r++)
            if (reinterpret_cast<uintptr_t>(p.peer.C[r]) % 16 != 0) p.tma_out = 0;
        if (p.peer.mc && reinterpret_cast<uintptr_t>(p.peer.mc) % 16 != 0) p.tma_out = 0;
    }
    switch (wtype) {
    case QGEMM_TYPE_Q4_0: return nat::launch_t<QGEMM_TYPE_Q4_0>(p, wgt, refseq, num_sms, st, tokn);
    case QGEMM_TYPE_Q4_1: return nat::launch_t<QGEMM_TYPE_Q4_1>(p, wgt, refseq, num_sms, st, tokn);
    case QGEMM_TYPE_Q5_0: return nat::launch_t<QGEMM_TYPE_Q5_0>(p, wgt, refseq, num_sms, st, tokn);
    case QGEMM_TYPE_Q5_1: return nat::launch_t<QGEMM_TYPE_Q5_1>(p, wgt, refseq, num_sms, st, tokn);
    case QGEMM_TYPE_Q8_0: return nat::launch_t<QGEMM_TYPE_Q8_0>(p, wgt, refseq, num_sms, st, tokn);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace qgemm
