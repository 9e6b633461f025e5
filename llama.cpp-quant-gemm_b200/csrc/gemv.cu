// gemv.cu -- decode path (T <= 8 tokens per pass): an HBM-bound weight stream.
//
// Replaces every decode-shaped kernel of the reference
// (kernels/gemm/gemm_warp_optimized.cuh:108-1070, gemm_async_copy.cuh:65-232,
// gemm_vectorized.cuh:66-230): those re-derive one idea -- keep the q8_1
// activations close and stream the weight rows once -- and top out at 42 % of
// their GPU's DRAM bandwidth.
//
// Design (B200):
//   * two CTAs per SM (<= 110 KB smem, <= 113 registers each), every CTA owns a CONTIGUOUS
//     span of weight rows (F split to one-row granularity: at most 1 row of imbalance); since rows are
//     contiguous in memory the span is one byte range, cut into tiles of RT rows
//   * a producer warp streams tiles HBM -> smem with ONE 1-D bulk async copy
//     per tile (cp.async.bulk, the TMA engine) through a ring of mbarrier-guarded
//     stages: tens of KB in flight per SM, no registers or L1 involved
//   * 8 consumer warps; a warp (or WPR warps for long rows) owns a row of the
//     tile, a lane owns fixed K positions: pairs of adjacent blocks (a pair starts
//     4-byte aligned in every format, a single 18/22/34-byte block does not)
//   * because a lane's K positions never change, its q8_1 activations live in
//     REGISTERS for the whole kernel (T*PPL <= 3 pairs); larger T keep them in smem,
//     re-laid out so consecutive lanes read consecutive 16-byte quads
//   * integer dot by dp4a on UN-offset weights, then the reference's exact
//     per-block fold (qgemm_common.cuh); lane-sequential over K, butterfly across
//     lanes, fixed order across the warps of a row -> deterministic
//   * optional programmatic dependent launch (QGEMM_WEIGHTS_STATIC): the weight
//     prefetch of launch n+1 overlaps the tail of launch n
#include <cstdlib>

#include "ptx.cuh"
#include "qgemm_common.cuh"

namespace qgemm {

constexpr int kGemvWarps = 8;                        // consumer warps
constexpr int kGemvThreads = (kGemvWarps + 1) * 32;  // + 1 producer warp
constexpr int kGemvMaxStages = 8;
constexpr int kGemvInflightTarget = 56 * 1024;     // bytes of weight tiles a CTA keeps in flight (see gemv_plan)
constexpr int kGemvSlotsOff = 128;                                         // after full[] / empty[]
constexpr int kGemvAbarOff = kGemvSlotsOff + 2 * kGemvWarps * 8 * 4;      // mbarrier of the activation bulk copy
constexpr int kGemvActOff = (kGemvAbarOff + 8 + 127) / 128 * 128;         // barriers + slots + abar
static_assert(2 * kGemvMaxStages * 8 <= kGemvSlotsOff, "ring barriers overlap the slots");
static_assert(kGemvAbarOff % 8 == 0 && kGemvAbarOff + 8 <= kGemvActOff, "abar must not overlap the bulk-copy destination");
constexpr int kGemvSmemBudget = 110 * 1024;          // two CTAs per SM, always
constexpr int kGemvCtasPerSm = 2;
constexpr int kGemvMaxGroup = 8;                      // matrices per grouped launch
constexpr int kGemvTileTarget = 24 * 1024;           // bytes per tile we aim for

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

// 4-bit: high nibbles stay in place as 16*hi (a valid u8 < 256); sumi4() undoes the factor exactly.
__device__ __forceinline__ void expand4(const uint32_t (&q)[4], uint32_t (&w)[8]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        w[i] = q[i] & 0x0f0f0f0fu;
        w[i + 4] = q[i] & 0xf0f0f0f0u;
    }
}
template <int WT>
__device__ __forceinline__ int pair_sumi(const uint32_t (&w)[8], const int (&a)[8]) {
    if constexpr (Fmt<WT>::bits == 4) {
        int lo = 0, hi = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            lo = dp4a_us(w[i], a[i], lo);
            hi = dp4a_us(w[i + 4], a[i + 4], hi);
        }
        return lo + (hi >> 4);  // hi is a multiple of 16: exact
    } else {
        return block_sumi<WT>(w, a);
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

// q8_1 activations of one pair of blocks (72 bytes, 4-byte aligned) for one token
struct ActPair {
    int q[2][8];
    ActScale s[2];
};

struct GemvParams {
    const uint8_t* act;   // q8_1, first token of this pass
    const uint8_t* wgt;
    float* C;             // already offset to the first token of this pass
    int F, nb;
    int64_t ldc_t, ldc_f;
    int RT;               // weight rows per tile (multiple of 8 / WPR)
    int WPR;              // warps sharing one row: 1, 2, 4 or 8
    int stages;
    int stage_bytes;      // 128-byte multiple
    int pdl;              // programmatic dependent launch: 1 wait before the activations, 2 wait before exit
    int nocompute;        // tuning aid: stream the tiles, skip the math (memory-system ceiling)
    int act_bulk;         // activations are 16-byte aligned and a 16-byte multiple: staged with one bulk copy
    PeerOut peer;         // fused all-gather (world <= 1: plain store to C)
    const uint8_t* pf_ptr; // next launch's weights to pull into L2 (or null)
    unsigned long long pf_bytes;
    // grouped launch: nmat matrices that share the activations (fused q/k/v, gate/up); rows are
    // numbered across the group, matrix m owns [fstart[m], fstart[m+1]); nmat <= 1: wgt / C above
    int nmat;
    const uint8_t* wgt_m[kGemvMaxGroup];
    float* C_m[kGemvMaxGroup];
    int fstart[kGemvMaxGroup + 1];
    float skew;           // uneven row shares (see the kernel): 0 = equal
#ifdef QGEMM_GEMV_TRACE
    int trace_slot;       // timeline build only (profiles/microbench/decode_trace.cu)
#endif
};

#ifdef QGEMM_GEMV_TRACE
// Timeline build: per launch slot and CTA, %globaltimer at kernel entry, after the dependency wait, after the activations
// are in registers, and after the last store.  Never compiled into the product library.
constexpr int kTraceSlots = 256, kTraceCtas = 304;
__device__ unsigned long long g_gemv_trace[kTraceSlots][kTraceCtas][6];
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define GEMV_TRACE(i) do { if (threadIdx.x == 0 && blockIdx.x < kTraceCtas) g_gemv_trace[p.trace_slot][blockIdx.x][i] = gtimer(); } while (0)
#else
#define GEMV_TRACE(i)
#endif

// tile that starts at group row r0: rows until the tile size, the CTA's span or the matrix ends
struct TileRef { int m, local, rows; };
__device__ __forceinline__ TileRef tile_at(const GemvParams& p, int r0, int r_end) {
    TileRef t{0, r0, min(p.RT, r_end - r0)};
    if (p.nmat > 1) {
        int m = 0;
        while (m + 1 < p.nmat && r0 >= p.fstart[m + 1]) m++;
        t.m = m;
        t.local = r0 - p.fstart[m];
        t.rows = min(t.rows, p.fstart[m + 1] - r0);
    }
    return t;
}

// PPL > 0: activations in registers, PPL pairs per lane.  PPL == 0: activations in smem.
// kFull: every lane owns exactly PPL valid pairs (np == PPL * WPR * 32): no bounds checks, so the
// compiler can interleave the independent pairs of a row.
template <int WT, int TT, int PPL, bool kMsExact, bool kFull>
__global__ void __launch_bounds__(kGemvThreads, kGemvCtasPerSm) gemv_kernel(const GemvParams p) {
    using Fm = Fmt<WT>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = p.nb, np = nb >> 1;

    // ---- carve shared memory (integer offsets from the __shared__ base keep every access an LDS)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [kGemvMaxStages]
    uint64_t* empty = full + kGemvMaxStages;                     // [kGemvMaxStages]
    float* slots = reinterpret_cast<float*>(smem + kGemvSlotsOff);  // [2][kGemvWarps][8] cross-warp partials
    // PPL > 0 : raw q8_1 copy [TT][nb][36 B];  PPL == 0 : [TT][np] float4 scales + [TT][4][np] quads
    uint8_t* a_raw = smem + kGemvActOff;
    float4* a_scale = reinterpret_cast<float4*>(smem + kGemvActOff);
    uint4* a_qs = reinterpret_cast<uint4*>(smem + kGemvActOff + (size_t)TT * np * 16);
    const uint32_t act_bytes = (PPL > 0) ? (uint32_t)TT * nb * 36u : (uint32_t)TT * nb * 40u;
    uint8_t* stage0 = smem + (((uint32_t)kGemvActOff + act_bytes + 127u) & ~127u);

    // ---- this CTA's contiguous span of weight rows.  Not quite equal shares: CTAs with low indices come in first and find
    // the head of the weights in L2 (the previous launch's hint), and the timeline (profiles/r02_decode_timeline.md) has them
    // finish 0.5-0.8 us before the others, so they get linearly more rows (density 1 + skew * (1/2 - x) over x = index / grid).
    auto span_at = [&](unsigned c) -> int {
        if (c >= gridDim.x) return p.F;
        const float x = (float)c / (float)gridDim.x;
        const float g = x * (1.0f + 0.5f * p.skew) - 0.5f * p.skew * x * x;
        return min(p.F, (int)((float)p.F * g + 0.5f));
    };
    const int r_begin = p.skew == 0.0f ? (int)(((int64_t)p.F * blockIdx.x) / gridDim.x) : span_at(blockIdx.x);
    const int r_end = p.skew == 0.0f ? (int)(((int64_t)p.F * (blockIdx.x + 1)) / gridDim.x) : span_at(blockIdx.x + 1);
    const size_t rowbytes = (size_t)nb * Fm::bytes;

    uint64_t* abar = reinterpret_cast<uint64_t*>(smem + kGemvAbarOff);  // activation copy: own 8 bytes in front of a_raw
    if (tid == 0) {
        for (int s = 0; s < p.stages; s++) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], kGemvWarps);
        }
        ptx::mbar_init(abar, 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();
    if (p.pdl) ptx::griddep_launch_dependents();  // dependents may start their own weight prefetch

    if (warp == kGemvWarps) {
        // ================= producer warp: weights do not depend on the previous launch
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int r0 = r_begin; r0 < r_end;) {
                const TileRef tr = tile_at(p, r0, r_end);
                const uint8_t* src = (p.nmat > 1 ? p.wgt_m[tr.m] : p.wgt) + (size_t)tr.local * rowbytes;
                const uint32_t bytes = (uint32_t)(tr.rows * rowbytes);
                ptx::mbar_wait(&empty[s], ph ^ 1);
                ptx::mbar_arrive_expect_tx(&full[s], bytes);
                ptx::bulk_g2s(stage0 + (size_t)s * p.stage_bytes, src, bytes, &full[s]);
                if (++s == p.stages) { s = 0; ph ^= 1; }
                r0 += tr.rows;
            }
            // own stream issued: now pull this CTA's share of the NEXT launch's weights into L2
            if (p.pf_ptr) {
                const unsigned long long chunk = ((p.pf_bytes / gridDim.x) + 15ull) & ~15ull;
                unsigned long long off = chunk * blockIdx.x;
                const unsigned long long end = min(off + chunk, p.pf_bytes & ~15ull);
                for (; off < end; off += 16384ull)
                    ptx::bulk_prefetch_l2(p.pf_ptr + off, (uint32_t)min(16384ull, end - off));
            }
        }
        return;
    }

    // ================= consumer warps =================
    GEMV_TRACE(0);
    if (p.pdl == 1) ptx::griddep_wait();  // activations / C may belong to the previous launch
    GEMV_TRACE(1);
    if (p.peer.world > 1) {          // ... or to an earlier launch of a peer GPU
        if (tid == 0) peer_wait_prior(p.peer);
        ptx::bar_sync(1, kGemvWarps * 32);
    }
    const int WPR = p.WPR;
    const int rpp = kGemvWarps / WPR;            // rows per pass
    const int rslot = warp / WPR, sub = warp - rslot * WPR;

    ActPair areg[PPL > 0 ? TT : 1][PPL > 0 ? PPL : 1];
    if constexpr (PPL > 0) {
        // one coalesced pass global -> smem (every word once per CTA), then each lane lifts the
        // pairs it owns into registers for the rest of the kernel
        const uint32_t* a32 = reinterpret_cast<const uint32_t*>(p.act);
        uint32_t* dst = reinterpret_cast<uint32_t*>(a_raw);
        const int total = TT * nb * 9;
        if (p.act_bulk) {   // one bulk copy (TMA) instead of load / store / barrier through the registers
            if (tid == 0) {
                ptx::mbar_arrive_expect_tx(abar, (uint32_t)total * 4u);
                ptx::bulk_g2s(a_raw, p.act, (uint32_t)total * 4u, abar);
            }
            ptx::mbar_wait(abar, 0);
            GEMV_TRACE(4);
        } else {
            for (int i = tid; i < total; i += kGemvWarps * 32) dst[i] = __ldg(a32 + i);
            ptx::bar_sync(1, kGemvWarps * 32);
        }
#pragma unroll
        for (int j = 0; j < PPL; j++) {
            const int pg = (j * WPR + sub) * 32 + lane;
#pragma unroll
            for (int t = 0; t < TT; t++) {
                if (kFull || j < PPL - 1 || pg < np) {
                    const uint32_t* w = dst + ((size_t)t * nb + 2 * pg) * 9;
#pragma unroll
                    for (int b = 0; b < 2; b++) {
                        const uint32_t ds = w[9 * b];
                        areg[t][j].s[b] = prep_act_scale<WT, kMsExact>(half_bits_to_float(ds), half_bits_to_float(ds >> 16));
#pragma unroll
                        for (int i = 0; i < 8; i++) areg[t][j].q[b][i] = (int)w[9 * b + 1 + i];
                    }
                }
            }
        }
    } else {
        // stage q8_1 AoS -> quad-interleaved SoA + fp32 scales (consumer warps only)
        const uint32_t* a32 = reinterpret_cast<const uint32_t*>(p.act);
        float* sc = reinterpret_cast<float*>(a_scale);
        uint32_t* qs = reinterpret_cast<uint32_t*>(a_qs);
        const int total = TT * nb * 9;
        for (int i = tid; i < total; i += kGemvWarps * 32) {
            const int blk = i / 9, wd = i - blk * 9;
            const int t = blk / nb, b = blk - t * nb;
            const uint32_t v = __ldg(a32 + i);
            if (wd == 0) {
                const ActScale ps = prep_act_scale<WT, kMsExact>(half_bits_to_float(v), half_bits_to_float(v >> 16));
                float* d = sc + ((size_t)t * np + (b >> 1)) * 4 + (b & 1) * 2;
                d[0] = ps.d;
                d[1] = ps.s;
            } else {
                const int e = wd - 1;
                const int quad = (b & 1) * 2 + (e >> 2);
                qs[(((size_t)t * 4 + quad) * np + (b >> 1)) * 4 + (e & 3)] = v;
            }
        }
        ptx::bar_sync(1, kGemvWarps * 32);
    }

    GEMV_TRACE(2);
    int s = 0, spar = 0;
    uint32_t ph = 0;
    const int npl = (PPL > 0) ? PPL : (np + WPR * 32 - 1) / (WPR * 32);  // pairs per lane
    for (int g0 = r_begin; g0 < r_end;) {
        const TileRef tr = tile_at(p, g0, r_end);
        const int r0 = tr.local, rows = tr.rows;      // rows r0 .. r0+rows-1 of matrix tr.m
        float* Cm = p.nmat > 1 ? p.C_m[tr.m] : p.C;
        g0 += rows;
        ptx::mbar_wait(&full[s], ph);
        const uint8_t* tile = stage0 + (size_t)s * p.stage_bytes;
        for (int pass = 0; pass * rpp < rows && !p.nocompute; pass++) {
            const int r = pass * rpp + rslot;
            float acc[TT];
#pragma unroll
            for (int tt = 0; tt < TT; tt++) acc[tt] = 0.f;
            if (r < rows) {
                const uint8_t* rowp = tile + (size_t)r * rowbytes;
                auto do_pair = [&](int j, int pg) {
                    uint32_t x[Pair<WT>::words];
                    uint32_t w[2][8];
                    WScale ws[2];
                    load_pair<WT>(rowp + (size_t)pg * (2 * Fm::bytes), x);
                    expand_pair<WT>(x, w, ws);
#pragma unroll
                    for (int tt = 0; tt < TT; tt++) {
                        if constexpr (PPL > 0) {
                            const ActPair& a = areg[tt][j];
                            acc[tt] = fold_block_pre<WT>(acc[tt], pair_sumi<WT>(w[0], a.q[0]), ws[0], a.s[0]);
                            acc[tt] = fold_block_pre<WT>(acc[tt], pair_sumi<WT>(w[1], a.q[1]), ws[1], a.s[1]);
                        } else {
                            const float4 sc = a_scale[(size_t)tt * np + pg];
                            int a0[8], a1[8];
                            const uint4 q0 = a_qs[((size_t)tt * 4 + 0) * np + pg];
                            const uint4 q1 = a_qs[((size_t)tt * 4 + 1) * np + pg];
                            const uint4 q2 = a_qs[((size_t)tt * 4 + 2) * np + pg];
                            const uint4 q3 = a_qs[((size_t)tt * 4 + 3) * np + pg];
                            a0[0] = q0.x; a0[1] = q0.y; a0[2] = q0.z; a0[3] = q0.w;
                            a0[4] = q1.x; a0[5] = q1.y; a0[6] = q1.z; a0[7] = q1.w;
                            a1[0] = q2.x; a1[1] = q2.y; a1[2] = q2.z; a1[3] = q2.w;
                            a1[4] = q3.x; a1[5] = q3.y; a1[6] = q3.z; a1[7] = q3.w;
                            acc[tt] = fold_block_pre<WT>(acc[tt], pair_sumi<WT>(w[0], a0), ws[0], ActScale{sc.x, sc.y});
                            acc[tt] = fold_block_pre<WT>(acc[tt], pair_sumi<WT>(w[1], a1), ws[1], ActScale{sc.z, sc.w});
                        }
                    }
                };
                if constexpr (PPL > 0) {
#pragma unroll
                    for (int j = 0; j < PPL; j++) {
                        const int pg = (j * WPR + sub) * 32 + lane;
                        // by construction of PPL only the last index can fall off the row: the others run
                        // unpredicated so the compiler interleaves their independent chains
                        if (kFull || j < PPL - 1 || pg < np) do_pair(j, pg);
                    }
                } else {
                    for (int j = 0; j < npl; j++) {
                        const int pg = (j * WPR + sub) * 32 + lane;
                        if (pg < np) do_pair(0, pg);
                    }
                }
            }
#pragma unroll
            for (int tt = 0; tt < TT; tt++) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[tt] += __shfl_xor_sync(0xffffffffu, acc[tt], o);
            }
            if (WPR == 1) {
                if (r < rows && lane < TT) {
                    float v = acc[0];
#pragma unroll
                    for (int tt = 1; tt < TT; tt++) v = (lane == tt) ? acc[tt] : v;
                    peer_store(p.peer, Cm, (int64_t)lane * p.ldc_t + (int64_t)(r0 + r) * p.ldc_f, v, tr.m);
                }
            } else {
                // only the WPR warps that share this row meet (named barrier 2 + row slot); the first
                // of them adds the partials in warp order and stores
                float* sl = slots + spar * (kGemvWarps * 8);
                if (lane == 0) {
#pragma unroll
                    for (int tt = 0; tt < TT; tt++) sl[warp * 8 + tt] = acc[tt];
                }
                ptx::bar_sync(2 + rslot, WPR * 32);
                if (sub == 0 && lane < TT && r < rows) {
                    float v = 0.f;
                    for (int k = 0; k < WPR; k++) v += sl[(rslot * WPR + k) * 8 + lane];
                    peer_store(p.peer, Cm, (int64_t)lane * p.ldc_t + (int64_t)(r0 + r) * p.ldc_f, v, tr.m);
                }
                spar ^= 1;
            }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    GEMV_TRACE(3);
    if (p.peer.world > 1) {
        if (!(p.peer.dbg & 4)) __threadfence();  // this thread's peer stores are ordered before the CTA barrier
        ptx::bar_sync(1, kGemvWarps * 32);
        if (tid == 0) peer_signal_done(p.peer, gridDim.x);
    }
    if (p.pdl == 2) ptx::griddep_wait();  // QGEMM_INPUTS_READY: ran ahead, but do not complete before the predecessor
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct GemvPlan {
    int tt;      // tokens per pass
    int ppl;     // pairs per lane held in registers (0: activations in smem)
    int wpr;     // warps per row
    int rt;      // rows per tile
    int stages, stage_bytes;
    size_t smem;
};

static bool gemv_plan(int wtype, int T, int F, int K, int grid, bool pdl, GemvPlan* pl);

// Rows must be bulk-copyable: whole rows 16-byte multiples from a 16-byte aligned base, and at least a
// one-token plan must fit the shared-memory budget (very long rows do not: those shapes belong to the next path).
bool gemv_supported(int wtype, const void* act, const void* wgt, int F, int K) {
    const int nb = K / 32;
    const size_t rowbytes = (size_t)nb * block_bytes(wtype);
    if (F < 1 || nb < 2 || (nb & 1)) return false;
    if (rowbytes % 16 != 0 || rowbytes > 64 * 1024) return false;
    if (reinterpret_cast<uintptr_t>(wgt) % 16 != 0) return false;
    if (reinterpret_cast<uintptr_t>(act) % 4 != 0) return false;
    GemvPlan pl;
    return gemv_plan(wtype, 1, F, K, 1, false, &pl);
}

static bool gemv_plan(int wtype, int T, int F, int K, int grid, bool pdl, GemvPlan* pl) {
    const int nb = K / 32, np = nb / 2;
    const size_t rowbytes = (size_t)nb * block_bytes(wtype);
    for (int tt = min(T, 8); tt >= 1; tt--) {
        // register-resident activations when T*PPL <= 6 with the fewest warps per row
        int ppl = 0, wpr = 1;
        for (int w = 1; w <= kGemvWarps; w <<= 1) {
            const int need = (np + 32 * w - 1) / (32 * w);
            if (need * tt <= 3) { ppl = need; wpr = w; break; }  // <= 60 activation registers
        }
        if (QGEMM_ENV("QGEMM_GEMV_FORCE_SMEM")) ppl = 0;  // tuning aid
        if (ppl == 0) {  // smem activations: spread long rows over more warps
            wpr = 1;
            while (wpr < kGemvWarps && np > 64 * wpr) wpr <<= 1;
        }
        const int rpp = kGemvWarps / wpr;
        int rt = rpp;
        while (rt < 8 && (size_t)(2 * rt) * rowbytes <= (size_t)kGemvTileTarget) rt <<= 1;
        if (const char* e = QGEMM_ENV("QGEMM_GEMV_RT")) rt = max(rpp, atoi(e) / rpp * rpp);  // tuning aid
        const size_t fixed = kGemvActOff + (size_t)tt * nb * (ppl == 0 ? 40 : 36) + 128;
        auto stage_of = [&](int r) { return (int)(((size_t)r * rowbytes + 127) / 128 * 128); };
        // long rows: a tile may hold fewer rows than one pass covers (the other row slots idle) before the shape
        // is given up to the next path
        while (rt > 1 && fixed + 2 * (size_t)stage_of(rt) > (size_t)kGemvSmemBudget) rt >>= 1;
        const int stage_bytes = stage_of(rt);
        if (fixed + 2 * (size_t)stage_bytes > (size_t)kGemvSmemBudget) continue;
        int stages = (int)(((size_t)kGemvSmemBudget - fixed) / stage_bytes);
        stages = max(2, min(kGemvMaxStages, stages));
        // no point in more ring than one CTA can ever use; with PDL stay under half an SM's
        // shared memory when possible so the next launch's CTA can move in early
        const int rows_per_cta = (F + grid - 1) / grid;
        stages = max(2, min(stages, (rows_per_cta + rt - 1) / rt));
        // A CTA's stream is short (a few tiles), and every bulk copy in flight shares the CTA's bandwidth: with the
        // whole ring issued at once the FIRST tile arrives late and the consumers start late.  ~56 KB in flight per
        // CTA (2 x 296 CTAs = 16 MB, still more than HBM latency x bandwidth) measured best: q4_0 11008x4096 T=1
        // 7.1 -> 6.2 us, decode stack +2.3 % (A/B on one box).
        stages = max(2, min(stages, kGemvInflightTarget / stage_bytes));
        (void)pdl;
        if (const char* e = QGEMM_ENV("QGEMM_GEMV_STAGES")) stages = max(2, min(kGemvMaxStages, atoi(e)));  // tuning aid
        *pl = {tt, ppl, wpr, rt, stages, stage_bytes, fixed + (size_t)stages * stage_bytes};
        return true;
    }
    return false;
}

template <int WT, int TT, int PPL>
static cudaError_t launch_gemv_inst(const GemvParams& p, size_t smem, int grid, bool ms_exact, cudaStream_t st) {
    auto launch = [&](auto kernel, int variant) -> cudaError_t {
        (void)variant;
        if (cudaError_t e = smem_optin(reinterpret_cast<const void*>(kernel), smem)) return e;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kGemvThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = p.pdl ? 1 : 0;
        return cudaLaunchKernelEx(&cfg, kernel, p);
    };
    cudaError_t e;
    const bool full = PPL > 0 && (p.nb / 2) == PPL * p.WPR * 32;
    if constexpr (Fmt<WT>::m >= 0) {  // only q4_1 / q5_1 have an m*s term
        if (full) e = ms_exact ? launch(gemv_kernel<WT, TT, PPL, true, true>, 2) : launch(gemv_kernel<WT, TT, PPL, false, true>, 3);
        else e = ms_exact ? launch(gemv_kernel<WT, TT, PPL, true, false>, 1) : launch(gemv_kernel<WT, TT, PPL, false, false>, 0);
    } else {
        e = full ? launch(gemv_kernel<WT, TT, PPL, false, true>, 3) : launch(gemv_kernel<WT, TT, PPL, false, false>, 0);
    }
    note_launch();
    return e;
}

template <int WT>
static cudaError_t launch_gemv_wt(const GemvPlan& pl, const GemvParams& p, int grid, bool ms, cudaStream_t st) {
#define QG_CASE(TTv, PPLv) \
    if (pl.tt == TTv && pl.ppl == PPLv) return launch_gemv_inst<WT, TTv, PPLv>(p, pl.smem, grid, ms, st);
    QG_CASE(1, 1) QG_CASE(1, 2) QG_CASE(1, 3) QG_CASE(2, 1) QG_CASE(3, 1)
    QG_CASE(1, 0) QG_CASE(2, 0) QG_CASE(3, 0) QG_CASE(4, 0) QG_CASE(5, 0) QG_CASE(6, 0) QG_CASE(7, 0) QG_CASE(8, 0)
#undef QG_CASE
    return cudaErrorInvalidValue;
}

struct GemvGroup { int nmat; const void* wgt[kGemvMaxGroup]; float* C[kGemvMaxGroup]; int F[kGemvMaxGroup]; };

cudaError_t launch_gemv(int wtype, const void* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                        int64_t ldc_f, uint32_t flags, int num_sms, cudaStream_t st, const PeerOut* peer,
                        const void* pf_ptr, size_t pf_bytes, const GemvGroup* group);

cudaError_t launch_gemv(int wtype, const void* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                        int64_t ldc_f, uint32_t flags, int num_sms, cudaStream_t st, const PeerOut* peer,
                        const void* pf_ptr, size_t pf_bytes) {
    return launch_gemv(wtype, act, wgt, C, T, F, K, ldc_t, ldc_f, flags, num_sms, st, peer, pf_ptr, pf_bytes, nullptr);
}

// grouped: F = total rows of the group; wgt / C ignored when group != nullptr
cudaError_t launch_gemv(int wtype, const void* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                        int64_t ldc_f, uint32_t flags, int num_sms, cudaStream_t st, const PeerOut* peer,
                        const void* pf_ptr, size_t pf_bytes, const GemvGroup* group) {
    const int nb = K / 32;
    const bool ms = flags & QGEMM_MS_EXACT;
    GemvPlan pl;
    int ctas_per_sm = kGemvCtasPerSm;
    if (const char* e = QGEMM_ENV("QGEMM_GEMV_CTAS")) ctas_per_sm = max(1, min(kGemvCtasPerSm, atoi(e)));  // tuning aid
    const int grid = min(F, ctas_per_sm * num_sms);
    const bool pdl = flags & QGEMM_WEIGHTS_STATIC;
    if (!gemv_plan(wtype, T, F, K, grid, pdl, &pl)) return cudaErrorInvalidValue;
    for (int t0 = 0; t0 < T; t0 += pl.tt) {
        GemvPlan cur = pl;
        if (T - t0 < pl.tt && !gemv_plan(wtype, T - t0, F, K, grid, pdl, &cur)) return cudaErrorInvalidValue;
        GemvParams p;
        p.act = (const uint8_t*)act + (size_t)t0 * nb * kQ81Bytes;
        p.wgt = (const uint8_t*)wgt;
        p.C = C + (int64_t)t0 * ldc_t;
        p.F = F; p.nb = nb; p.ldc_t = ldc_t; p.ldc_f = ldc_f;
        p.RT = cur.rt; p.WPR = cur.wpr; p.stages = cur.stages; p.stage_bytes = cur.stage_bytes;
        p.pdl = pdl ? ((flags & QGEMM_INPUTS_READY) ? 2 : 1) : 0;
        p.nocompute = QGEMM_ENV("QGEMM_GEMV_NOCOMPUTE") ? 1 : 0;
        p.act_bulk = (reinterpret_cast<uintptr_t>(p.act) % 16 == 0 && ((size_t)cur.tt * nb * kQ81Bytes) % 16 == 0 &&
                      !QGEMM_ENV("QGEMM_GEMV_NO_ACT_BULK")) ? 1 : 0;
        p.pf_ptr = (t0 + pl.tt >= T && reinterpret_cast<uintptr_t>(pf_ptr) % 16 == 0) ? (const uint8_t*)pf_ptr : nullptr;
        p.pf_bytes = pf_bytes;
        // measured on the decode stack (profiles/r02_decode_timeline.md): 0 -> 881, 0.1 -> 875, 0.2 -> 866, 0.25 / 0.3 -> 860,
        // 0.35 -> 879, 0.4 -> 893 us per 128-launch step
        p.skew = QGEMM_ENV("QGEMM_GEMV_SKEW") ? (float)atof(QGEMM_ENV("QGEMM_GEMV_SKEW")) : (F >= 8 * grid ? 0.28f : 0.0f);
#ifdef QGEMM_GEMV_TRACE
        {
            static int seq = 0;
            p.trace_slot = seq++ % kTraceSlots;
        }
#endif
        p.nmat = 0;
        if (group && group->nmat == 1) {   // a group of one is a plain GEMV (the kernel reads wgt / C when nmat <= 1)
            p.wgt = (const uint8_t*)group->wgt[0];
            p.C = group->C[0] + (int64_t)t0 * ldc_t;
        } else if (group) {
            p.nmat = group->nmat;
            p.fstart[0] = 0;
            for (int m = 0; m < group->nmat; m++) {
                p.wgt_m[m] = (const uint8_t*)group->wgt[m];
                p.C_m[m] = group->C[m] + (int64_t)t0 * ldc_t;
                p.fstart[m + 1] = p.fstart[m] + group->F[m];
            }
        }
        p.peer = PeerOut{};
        if (peer) {
            if (T > pl.tt) return cudaErrorInvalidValue;  // peer mode: one pass per launch (flag accounting)
            p.peer = *peer;
        }
        cudaError_t e;
        switch (wtype) {
        case QGEMM_TYPE_Q4_0: e = launch_gemv_wt<QGEMM_TYPE_Q4_0>(cur, p, grid, ms, st); break;
        case QGEMM_TYPE_Q4_1: e = launch_gemv_wt<QGEMM_TYPE_Q4_1>(cur, p, grid, ms, st); break;
        case QGEMM_TYPE_Q5_0: e = launch_gemv_wt<QGEMM_TYPE_Q5_0>(cur, p, grid, ms, st); break;
        case QGEMM_TYPE_Q5_1: e = launch_gemv_wt<QGEMM_TYPE_Q5_1>(cur, p, grid, ms, st); break;
        case QGEMM_TYPE_Q8_0: e = launch_gemv_wt<QGEMM_TYPE_Q8_0>(cur, p, grid, ms, st); break;
        default: e = cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}


// ===========================================================================================================
// Chained decode GEMVs: ONE persistent launch walks a list of steps (each a GEMV or a group of GEMVs that share
// their activations), with device-side dependencies between the steps.  A chain of separate launches pays, per
// launch, the kernel boundary, an empty weight ring and the activation prologue (profiles/r02_decode_timeline.md:
// ~2.3 us on top of a 4.7 us stream); here the producer warp never stops -- while the consumers of a CTA sit in
// the barrier between two steps and fetch the next step's activations, the ring fills with the next step's
// weights -- the intent being that a boundary costs the consumers' catch-up instead of an idle HBM.  Measured
// (profiles/r02_chain.md): a boundary inside the kernel (fence + device-wide barrier + activation fetch) costs more than a
// programmatic launch boundary and the consumers have no speed in reserve to catch up with, so this form is SLOWER than one
// launch per projection and is not what AUTO or bench.py use; it is API surface (one launch per layer, in-kernel quantization).
//   * same consumer code as gemv_kernel (one token, activations in registers), instantiated per (PPL, kFull)
//     and selected per step: the steps of a chain may differ in K and F
//   * a step that waits (QGEMM_INPUTS_READY not set) starts only after every CTA has finished every earlier
//     step: per-step arrival counters in `sync` (device fence + barrier + one atomic per CTA and step; acquire
//     spin by one thread).  All CTAs are co-resident (the host clamps the grid to the occupancy).
//   * a step's activations are either ready-made q8_1 blocks (one bulk copy per CTA) or quantized inside the kernel from
//     an fp32 vector -- typically the output of an earlier step of the same chain, optionally through SwiGLU
//     (silu(x) * gate) -- with quantize_q8_1's arithmetic (include/quantize.h:165-193, default flags): the CTAs quantize
//     the vector together (one warp per block, scratch in `sync`, a second device-wide barrier) and every CTA copies the
//     result; the quantize launch between two projections disappears
//   * counters return to zero: the last CTA to leave clears them
// ===========================================================================================================
constexpr int kChainMaxSteps = 160;
constexpr int kChainMaxMat = 3;
constexpr int kChainQStride = 1536 * 36;   // longest row the register-resident consumers take: 3 pairs x 8 warps x 32 lanes

struct ChainStep {
    const uint8_t* act;              // q8_1 [nb][36 B], or null: quantize from x (and gate)
    const float* x;                  // fp32 [K] source of the activations when act == null
    const float* gate;               // optional: quantize silu(x) * gate
    const uint8_t* wgt[kChainMaxMat];
    float* C[kChainMaxMat];
    int fstart[kChainMaxMat + 1];    // rows numbered across the step's matrices
    int nb;
    int ldc_f;                       // C[m][f * ldc_f]  (one token)
    unsigned char nmat, RT, WPR, ppl, full, wait;
};

template <int NS> struct ChainParams {
    int nsteps;
    int stages, stage_bytes;
    uint32_t act_bytes;              // shared memory reserved for the raw activations (largest step)
    unsigned* sync;                  // [2 * nsteps + 1] counters, zero on entry and on exit: step arrivals [0, nsteps), exit
                                     // ticket [nsteps], quantizer arrivals [nsteps + 1, 2 * nsteps + 1)
    uint8_t* qscratch;               // two buffers of kChainQStride bytes: the q8_1 blocks of a step that quantizes its input
    int pdl;
    const uint8_t* pf_ptr;
    unsigned long long pf_bytes;
    ChainStep st[NS];
};
constexpr int kChainSmall = 8;   // short chains (a layer) do not carry the long list's 16 KB of parameters

__device__ __forceinline__ uint32_t ld_acquire_gpu(const unsigned* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct ChainTile { int m, local, rows; };
__device__ __forceinline__ ChainTile chain_tile_at(const ChainStep& st, int r0, int r_end) {
    int m = 0;
    while (m + 1 < st.nmat && r0 >= st.fstart[m + 1]) m++;
    return {m, r0 - st.fstart[m], min(min((int)st.RT, r_end - r0), st.fstart[m + 1] - r0)};
}

// quantize_q8_1 of one block (default flags: roundf, clamp -128, s = fp16 of the sequential fp32 sum), 9 words out
__device__ __forceinline__ int chain_round_half_away(float v) {
    const float t = truncf(v);
    const float r = v - t;
    int q = __float2int_rz(t);
    if (fabsf(r) >= 0.5f) q += (v < 0.0f) ? -1 : 1;
    return q;
}
// One warp quantizes one block, lane = element: quantize_q8_1's default arithmetic (include/quantize.h:165-193: roundf, clamp
// -128, s = fp16 of the fp32 sum taken in element order), optionally on silu(x) * gate in silu_mul_f32_kernel's operation
// sequence (kernels/activation/silu.cuh:97-108).  36 bytes out.
// Not inlined on purpose: inside the chain kernel's register budget (96, with spills) one build of this loop body faulted on its
// first load although every address it was given was right (printed from the kernel); as a function of its own it has its own
// register allocation.  It runs once per block and step: the call costs nothing that matters.
__device__ __noinline__ void chain_quantize_block_lanes(const float* x, const float* gate, uint8_t* out, int lane) {
    float v = __ldcg(x + lane);
    if (gate) v = __fmul_rn(__fdiv_rn(v, __fadd_rn(1.0f, expf(-v))), __ldcg(gate + lane));
    float amax = fabsf(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; j++) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, v, j));
    const float d = __fdiv_rn(amax, 127.0f);
    const float id = (d > 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
    const int q = max(-128, min(127, chain_round_half_away(__fmul_rn(v, id))));
    out[4 + lane] = (uint8_t)(q & 0xff);
    if (lane == 0)
        *reinterpret_cast<uint32_t*>(out) =
            (uint32_t)__half_as_ushort(__float2half_rn(d)) | ((uint32_t)__half_as_ushort(__float2half_rn(sum)) << 16);
}

// the consumer warps' share of one step; ring position (s, ph) and slot parity carry over from step to step
template <int WT, int PPL, bool kFull>
__device__ __forceinline__ void chain_consume(const ChainStep& st, const int r_begin, const int r_end, const uint32_t* araw,
                                              const uint8_t* stage0, const int stage_bytes, const int nstages, uint64_t* full,
                                              uint64_t* empty, float* slots, int& s, uint32_t& ph, int& spar) {
    using Fm = Fmt<WT>;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = st.nb, np = nb >> 1;
    const int WPR = st.WPR;
    const int rpp = kGemvWarps / WPR;
    const int rslot = warp / WPR, sub = warp - rslot * WPR;
    const size_t rowbytes = (size_t)nb * Fm::bytes;

    ActPair areg[PPL];
#pragma unroll
    for (int j = 0; j < PPL; j++) {
        const int pg = (j * WPR + sub) * 32 + lane;
        if (kFull || j < PPL - 1 || pg < np) {
            const uint32_t* w = araw + (size_t)(2 * pg) * 9;
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const uint32_t ds = w[9 * b];
                areg[j].s[b] = prep_act_scale<WT, false>(half_bits_to_float(ds), half_bits_to_float(ds >> 16));
#pragma unroll
                for (int i = 0; i < 8; i++) areg[j].q[b][i] = (int)w[9 * b + 1 + i];
            }
        }
    }
    for (int g0 = r_begin; g0 < r_end;) {
        const ChainTile tr = chain_tile_at(st, g0, r_end);
        const int r0 = tr.local, rows = tr.rows;
        float* Cm = st.C[tr.m];
        g0 += rows;
        ptx::mbar_wait(&full[s], ph);
        const uint8_t* tile = stage0 + (size_t)s * stage_bytes;
        for (int pass = 0; pass * rpp < rows; pass++) {
            const int r = pass * rpp + rslot;
            float acc = 0.f;
            if (r < rows) {
                const uint8_t* rowp = tile + (size_t)r * rowbytes;
#pragma unroll
                for (int j = 0; j < PPL; j++) {
                    const int pg = (j * WPR + sub) * 32 + lane;
                    if (kFull || j < PPL - 1 || pg < np) {
                        uint32_t x[Pair<WT>::words];
                        uint32_t w[2][8];
                        WScale ws[2];
                        load_pair<WT>(rowp + (size_t)pg * (2 * Fm::bytes), x);
                        expand_pair<WT>(x, w, ws);
                        acc = fold_block_pre<WT>(acc, pair_sumi<WT>(w[0], areg[j].q[0]), ws[0], areg[j].s[0]);
                        acc = fold_block_pre<WT>(acc, pair_sumi<WT>(w[1], areg[j].q[1]), ws[1], areg[j].s[1]);
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (WPR == 1) {
                if (r < rows && lane == 0) Cm[(int64_t)(r0 + r) * st.ldc_f] = acc;
            } else {
                float* sl = slots + spar * (kGemvWarps * 8);
                if (lane == 0) sl[warp * 8] = acc;
                ptx::bar_sync(2 + rslot, WPR * 32);
                if (sub == 0 && lane == 0 && r < rows) {
                    float v = 0.f;
                    for (int k = 0; k < WPR; k++) v += sl[(rslot * WPR + k) * 8];
                    Cm[(int64_t)(r0 + r) * st.ldc_f] = v;
                }
                spar ^= 1;
            }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[s]);
        if (++s == nstages) { s = 0; ph ^= 1; }
    }
}

template <int WT, int NS>
__global__ void __launch_bounds__(kGemvThreads, kGemvCtasPerSm) gemv_chain_kernel(const __grid_constant__ ChainParams<NS> p) {
    using Fm = Fmt<WT>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kGemvMaxStages;
    float* slots = reinterpret_cast<float*>(smem + kGemvSlotsOff);
    uint64_t* abar = reinterpret_cast<uint64_t*>(smem + kGemvAbarOff);
    uint8_t* a_raw = smem + kGemvActOff;
    uint8_t* stage0 = smem + (((uint32_t)kGemvActOff + p.act_bytes + 127u) & ~127u);
    const unsigned grid = gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < p.stages; s++) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], kGemvWarps);
        }
        ptx::mbar_init(abar, 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();
    if (p.pdl) ptx::griddep_launch_dependents();

    if (warp == kGemvWarps) {
        // ================= producer: one weight stream across all steps; never waits for a step boundary
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int k = 0; k < p.nsteps; k++) {
                const ChainStep& st = p.st[k];
                const int F = st.fstart[st.nmat];
                const int r_begin = (int)(((int64_t)F * blockIdx.x) / grid);
                const int r_end = (int)(((int64_t)F * (blockIdx.x + 1)) / grid);
                const size_t rowbytes = (size_t)st.nb * Fm::bytes;
                for (int r0 = r_begin; r0 < r_end;) {
                    const ChainTile tr = chain_tile_at(st, r0, r_end);
                    const uint8_t* src = st.wgt[tr.m] + (size_t)tr.local * rowbytes;
                    const uint32_t bytes = (uint32_t)(tr.rows * rowbytes);
                    ptx::mbar_wait(&empty[s], ph ^ 1);
                    ptx::mbar_arrive_expect_tx(&full[s], bytes);
                    ptx::bulk_g2s(stage0 + (size_t)s * p.stage_bytes, src, bytes, &full[s]);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                    r0 += tr.rows;
                }
            }
            if (p.pf_ptr) {
                const unsigned long long chunk = ((p.pf_bytes / grid) + 15ull) & ~15ull;
                unsigned long long off = chunk * blockIdx.x;
                const unsigned long long end = min(off + chunk, p.pf_bytes & ~15ull);
                for (; off < end; off += 16384ull)
                    ptx::bulk_prefetch_l2(p.pf_ptr + off, (uint32_t)min(16384ull, end - off));
            }
        }
        return;
    }

    // ================= consumers
    if (p.pdl) ptx::griddep_wait();   // activations, C and the counters may still belong to the previous launch
    int s = 0, spar = 0;
    uint32_t ph = 0, aph = 0;
    for (int k = 0; k < p.nsteps; k++) {
        const ChainStep& st = p.st[k];
        const int F = st.fstart[st.nmat];
        const int r_begin = (int)(((int64_t)F * blockIdx.x) / grid);
        const int r_end = (int)(((int64_t)F * (blockIdx.x + 1)) / grid);
        if (st.act) {
            if (tid == 0) {
                if (st.wait && k > 0)
                    while (ld_acquire_gpu(p.sync + (k - 1)) < grid) __nanosleep(20);
                const uint32_t bytes = (uint32_t)st.nb * 36u;
                ptx::mbar_arrive_expect_tx(abar, bytes);
                ptx::bulk_g2s(a_raw, st.act, bytes, abar);
            }
            ptx::mbar_wait(abar, aph);
            aph ^= 1;
        } else {
            // fp32 source: the CTAs quantize it TOGETHER, one block per warp (every CTA for itself would be 296-fold redundant:
            // 8 us per step measured, profiles/r02_chain.md), into a scratch buffer in global memory that every CTA then
            // copies.  Such a step always starts after every earlier step has completed (the source is usually one of their
            // outputs, and the two scratch buffers alternate on that order).
            if (k > 0) {
                if (tid == 0)
                    while (ld_acquire_gpu(p.sync + (k - 1)) < grid) __nanosleep(20);
                ptx::bar_sync(1, kGemvWarps * 32);
            }
            uint8_t* qbuf = p.qscratch + (size_t)(k & 1) * kChainQStride;
            for (int b = (int)blockIdx.x + warp * (int)grid; b < st.nb; b += kGemvWarps * (int)grid)
                chain_quantize_block_lanes(st.x + (size_t)b * 32, st.gate ? st.gate + (size_t)b * 32 : nullptr, qbuf + (size_t)b * 36, lane);
            __threadfence();
            ptx::bar_sync(1, kGemvWarps * 32);
            if (tid == 0) {
                atomicAdd(p.sync + p.nsteps + 1 + k, 1u);
                while (ld_acquire_gpu(p.sync + p.nsteps + 1 + k) < grid) __nanosleep(20);
            }
            ptx::bar_sync(1, kGemvWarps * 32);
            const uint32_t* src = reinterpret_cast<const uint32_t*>(qbuf);
            uint32_t* dst = reinterpret_cast<uint32_t*>(a_raw);
            for (int i = tid; i < st.nb * 9; i += kGemvWarps * 32) dst[i] = __ldcg(src + i);
            ptx::fence_proxy_async();   // a later step may overwrite a_raw with a bulk copy
            ptx::bar_sync(1, kGemvWarps * 32);
        }
        const uint32_t* araw = reinterpret_cast<const uint32_t*>(a_raw);
#define QG_CHAIN_CASE(P, FULL) \
    chain_consume<WT, P, FULL>(st, r_begin, r_end, araw, stage0, p.stage_bytes, p.stages, full, empty, slots, s, ph, spar)
        if (st.full) {
            if (st.ppl == 1) QG_CHAIN_CASE(1, true);
            else if (st.ppl == 2) QG_CHAIN_CASE(2, true);
            else QG_CHAIN_CASE(3, true);
        } else {
            if (st.ppl == 1) QG_CHAIN_CASE(1, false);
            else if (st.ppl == 2) QG_CHAIN_CASE(2, false);
            else QG_CHAIN_CASE(3, false);
        }
#undef QG_CHAIN_CASE
        // this CTA's share of step k is stored: publish it (and free a_raw for the next step)
        if (lane == 0) __threadfence();   // lane 0 of every warp is the one that stores
        ptx::bar_sync(1, kGemvWarps * 32);
        if (tid == 0) atomicAdd(p.sync + k, 1u);
    }
    if (tid == 0) {
        // the last CTA to leave has seen every other CTA past its last wait: the counters can go back to zero
        if (atomicAdd(p.sync + p.nsteps, 1u) == grid - 1) {
            for (int k = 0; k <= 2 * p.nsteps; k++) p.sync[k] = 0u;
        }
    }
}

struct ChainStepHost {
    const void* act; const float* x; const float* gate;
    int nmat; const void* wgt[kChainMaxMat]; float* C[kChainMaxMat]; int F[kChainMaxMat];
    int K; int ldc_f; int wait;
};

int gemv_chain_max_steps() { return kChainMaxSteps; }
// The counters come first and the quantizer scratch sits at a FIXED offset behind the longest list's counters: one buffer then
// serves chains of any length (with a length-dependent offset a short chain's scratch would land on -- and un-zero -- the
// counters of a longer one that shares the buffer).
static size_t gemv_chain_counter_bytes(int) { return ((size_t)(2 * kChainMaxSteps + 1) * sizeof(unsigned) + 255) / 256 * 256; }
size_t gemv_chain_sync_bytes(int nsteps) { return gemv_chain_counter_bytes(nsteps) + 2 * (size_t)kChainQStride; }

// Can this list run as one persistent launch?  (one token, register-resident activations for every step, rows and
// activations bulk-copyable).  Fills *out with the kernel parameters.
template <int NS>
static bool gemv_chain_plan(int wtype, const ChainStepHost* steps, int nsteps, int grid, ChainParams<NS>* out, size_t* smem) {
    if (nsteps < 1 || nsteps > NS) return false;
    int stage_bytes = 0;
    size_t act_bytes = 0;
    for (int k = 0; k < nsteps; k++) {
        const ChainStepHost& h = steps[k];
        if (h.nmat < 1 || h.nmat > kChainMaxMat) return false;
        const int nb = h.K / 32;
        int Ftot = 0;
        for (int m = 0; m < h.nmat; m++) {
            if (!gemv_supported(wtype, h.act ? h.act : (const void*)h.x, h.wgt[m], h.F[m], h.K)) return false;
            Ftot += h.F[m];
        }
        if (h.act) {
            if (reinterpret_cast<uintptr_t>(h.act) % 16 != 0 || ((size_t)nb * kQ81Bytes) % 16 != 0) return false;
        } else {
            if (reinterpret_cast<uintptr_t>(h.x) % 16 != 0 || (h.gate && reinterpret_cast<uintptr_t>(h.gate) % 16 != 0)) return false;
        }
        GemvPlan pl;
        if (!gemv_plan(wtype, 1, Ftot, h.K, grid, true, &pl) || pl.tt != 1 || pl.ppl < 1 || pl.ppl > 3) return false;
        ChainStep& st = out->st[k];
        st = ChainStep{};
        st.act = (const uint8_t*)h.act; st.x = h.x; st.gate = h.gate;
        st.fstart[0] = 0;
        for (int m = 0; m < h.nmat; m++) {
            st.wgt[m] = (const uint8_t*)h.wgt[m];
            st.C[m] = h.C[m];
            st.fstart[m + 1] = st.fstart[m] + h.F[m];
        }
        for (int m = h.nmat; m < kChainMaxMat; m++) st.fstart[m + 1] = st.fstart[h.nmat];
        st.nb = nb; st.ldc_f = h.ldc_f;
        st.nmat = (unsigned char)h.nmat; st.RT = (unsigned char)pl.rt; st.WPR = (unsigned char)pl.wpr; st.ppl = (unsigned char)pl.ppl;
        st.full = (nb / 2) == pl.ppl * pl.wpr * 32 ? 1 : 0;
        st.wait = h.wait ? 1 : 0;
        stage_bytes = max(stage_bytes, (int)(((size_t)pl.rt * nb * block_bytes(wtype) + 127) / 128 * 128));
        act_bytes = max(act_bytes, (size_t)nb * kQ81Bytes);
    }
    const size_t fixed = (((size_t)kGemvActOff + act_bytes + 127) & ~(size_t)127);
    if (fixed + 2 * (size_t)stage_bytes > (size_t)kGemvSmemBudget) return false;
    int stages = (int)(((size_t)kGemvSmemBudget - fixed) / stage_bytes);
    stages = max(2, min(kGemvMaxStages, stages));
    if (const char* e = QGEMM_ENV("QGEMM_CHAIN_STAGES")) stages = max(2, min(stages, atoi(e)));  // tuning aid
    out->nsteps = nsteps;
    out->stages = stages;
    out->stage_bytes = stage_bytes;
    out->act_bytes = (uint32_t)act_bytes;
    *smem = fixed + (size_t)stages * stage_bytes;
    return true;
}

template <int WT, int NS>
static cudaError_t launch_chain_wt(const ChainParams<NS>& p, size_t smem, int num_sms, cudaStream_t st) {
    auto kernel = gemv_chain_kernel<WT, NS>;
    if (cudaError_t e = smem_optin(reinterpret_cast<const void*>(kernel), smem)) return e;
    // the grid must be co-resident (steps wait for each other inside the kernel)
    int occ = 0;
    if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kGemvThreads, smem)) return e;
    if (occ < kGemvCtasPerSm) return cudaErrorLaunchOutOfResources;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kGemvCtasPerSm * num_sms);
    cfg.blockDim = dim3(kGemvThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = p.pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
    note_launch();
    return e;
}

template <int NS>
static cudaError_t launch_chain_ns(int wtype, const ChainStepHost* steps, int nsteps, uint32_t flags, unsigned* sync, int num_sms,
                                   cudaStream_t st, const void* pf_ptr, size_t pf_bytes) {
    static thread_local ChainParams<NS> p;   // up to 16 KB: not on the stack of a caller we do not know
    size_t smem = 0;
    const int grid = kGemvCtasPerSm * num_sms;
    if (!gemv_chain_plan<NS>(wtype, steps, nsteps, grid, &p, &smem)) return cudaErrorNotSupported;
    p.sync = sync;
    p.qscratch = reinterpret_cast<uint8_t*>(sync) + gemv_chain_counter_bytes(nsteps);
    p.pdl = (flags & QGEMM_WEIGHTS_STATIC) ? 1 : 0;
    p.pf_ptr = reinterpret_cast<uintptr_t>(pf_ptr) % 16 == 0 ? (const uint8_t*)pf_ptr : nullptr;
    p.pf_bytes = pf_bytes;
    switch (wtype) {
    case QGEMM_TYPE_Q4_0: return launch_chain_wt<QGEMM_TYPE_Q4_0, NS>(p, smem, num_sms, st);
    case QGEMM_TYPE_Q4_1: return launch_chain_wt<QGEMM_TYPE_Q4_1, NS>(p, smem, num_sms, st);
    case QGEMM_TYPE_Q5_0: return launch_chain_wt<QGEMM_TYPE_Q5_0, NS>(p, smem, num_sms, st);
    case QGEMM_TYPE_Q5_1: return launch_chain_wt<QGEMM_TYPE_Q5_1, NS>(p, smem, num_sms, st);
    case QGEMM_TYPE_Q8_0: return launch_chain_wt<QGEMM_TYPE_Q8_0, NS>(p, smem, num_sms, st);
    default: return cudaErrorInvalidValue;
    }
}

// cudaErrorNotSupported: the list does not fit the persistent kernel (the caller launches the steps one by one)
cudaError_t launch_gemv_chain(int wtype, const ChainStepHost* steps, int nsteps, uint32_t flags, unsigned* sync, int num_sms,
                              cudaStream_t st, const void* pf_ptr, size_t pf_bytes) {
    if ((flags & QGEMM_MS_EXACT) && (wtype == QGEMM_TYPE_Q4_1 || wtype == QGEMM_TYPE_Q5_1)) return cudaErrorNotSupported;
    const int grid = kGemvCtasPerSm * num_sms;
    for (int k = 0; k < nsteps; k++) {   // every CTA owns rows of every step
        int Ftot = 0;
        for (int m = 0; m < steps[k].nmat && m < kChainMaxMat; m++) Ftot += steps[k].F[m];
        if (Ftot < grid) return cudaErrorNotSupported;
    }
    if (nsteps <= kChainSmall) return launch_chain_ns<kChainSmall>(wtype, steps, nsteps, flags, sync, num_sms, st, pf_ptr, pf_bytes);
    return launch_chain_ns<kChainMaxSteps>(wtype, steps, nsteps, flags, sync, num_sms, st, pf_ptr, pf_bytes);
}

}  // namespace qgemm

#ifdef QGEMM_GEMV_TRACE
extern "C" __attribute__((visibility("default"))) int qgemm_debug_read_gemv_trace(unsigned long long* dst, size_t bytes) {
    return (int)cudaMemcpyFromSymbol(dst, qgemm::g_gemv_trace, bytes < sizeof(qgemm::g_gemv_trace) ? bytes : sizeof(qgemm::g_gemv_trace));
}
#endif
