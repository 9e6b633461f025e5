/*
 * qgemm_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see qgemm_oracle.h).
 *
 * Plain C, no CUDA headers: fp16 conversion is restated here so the file
 * builds with gcc and runs on a machine without a GPU.  Build with
 * -ffp-contract=off: the reference's host code is compiled without FMA
 * contraction (no -mfma in its Makefile:8-15 / CMakeLists.txt:40-46), so every
 * '*' and '+' below is one IEEE fp32 rounding, in C operator order.
 *
 * Parity: PINNED against the reference's own code and vectors, see
 * tests/test_oracle_golden.py and oracle/ref_shim.cu.
 */
#include "qgemm_oracle.h"

#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* fp16 <-> fp32.  The reference uses cuda_fp16.h's host __float2half /      */
/* __half2float (round-to-nearest-even, subnormals kept, inf/nan preserved). */
/* ------------------------------------------------------------------------- */
static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

uint16_t qo_fp32_to_fp16(float f)
{
    const uint32_t x = f2u(f);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t ax = x & 0x7fffffffu;

    if (ax >= 0x7f800000u) { /* inf / nan */
        if (ax > 0x7f800000u) return (uint16_t)0x7fffu; /* cuda_fp16: canonical NaN */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax >= 0x477ff000u) { /* >= 65520 rounds to inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x33000001u) { /* <= 2^-25 rounds to zero (tie at 2^-25 -> even = 0) */
        return (uint16_t)sign;
    }
    {
        int32_t e = (int32_t)(ax >> 23) - 127; /* unbiased */
        uint32_t m = (ax & 0x7fffffu) | 0x800000u; /* 24-bit significand */
        uint32_t shift, half_ulp, rest, q;
        if (e < -14) {
            /* subnormal half: value = m * 2^(e-23); target unit 2^-24 */
            shift = (uint32_t)(-14 - e) + 13u;
        } else {
            shift = 13u;
        }
        q = m >> shift;
        rest = m & ((1u << shift) - 1u);
        half_ulp = 1u << (shift - 1u);
        if (rest > half_ulp || (rest == half_ulp && (q & 1u))) q++;
        if (e < -14) {
            /* q is the subnormal mantissa, may carry into the normal range (q == 0x400) */
            return (uint16_t)(sign | q);
        }
        /* q holds 1.mmmmmmmmmm as 11 bits (0x400..0x800) */
        return (uint16_t)(sign | ((((uint32_t)(e + 15)) << 10) + (q - 0x400u)));
    }
}

float qo_fp16_to_fp32(uint16_t h)
{
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    const uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    if (e == 0x1fu) return u2f(sign | 0x7f800000u | (m << 13));
    if (e == 0) {
        if (m == 0) return u2f(sign);
        {
            int sh = 0;
            while (!(m & 0x400u)) { m <<= 1; sh++; }
            m &= 0x3ffu;
            return u2f(sign | ((uint32_t)(127 - 15 - sh + 1) << 23) | (m << 13));
        }
    }
    return u2f(sign | ((e + 112u) << 23) | (m << 13));
}

/* ------------------------------------------------------------------------- */
/* Block layouts, compat/ggml_types.h:62-191 (packed, little-endian halves).  */
/*   q4_0: d[2] qs[16]              = 18                                      */
/*   q4_1: d[2] m[2] qs[16]         = 20                                      */
/*   q5_0: d[2] qh[4] qs[16]        = 22                                      */
/*   q5_1: d[2] m[2] qh[4] qs[16]   = 24                                      */
/*   q8_0: d[2] qs[32]              = 34                                      */
/*   q8_1: d[2] s[2] qs[32]         = 36   (ds = half2: d low, s high)        */
/* ------------------------------------------------------------------------- */
size_t qo_block_bytes(int type)
{
    switch (type) {
    case QO_Q4_0: return 18;
    case QO_Q4_1: return 20;
    case QO_Q5_0: return 22;
    case QO_Q5_1: return 24;
    case QO_Q8_0: return 34;
    case QO_Q8_1: return 36;
    default: return 0;
    }
}

static uint16_t ld16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static void st16(uint8_t *p, uint16_t v) { p[0] = (uint8_t)(v & 0xff); p[1] = (uint8_t)(v >> 8); }
static uint32_t ld32(const uint8_t *p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static void st32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

/* ------------------------------------------------------------------------- */
/* quantize_q8_1                                                              */
/* ------------------------------------------------------------------------- */
/* round-half-even of a float already known to be |v| <= 2^22 or so:
 * __float2int_rn semantics (include/quantize.h:335). */
static int rn_even(float v) { return (int)nearbyintf(v); /* default FE_TONEAREST */ }

/* silu(x) * gate, element by element, in the operation order of silu_mul_f32_kernel (kernels/activation/silu.cuh:97-108;
 * CPU form of the silu part: silu_cpu_f32, silu.cuh:23-27): val / (1.0f + expf(-val)), then times gate.  expf is libm's
 * here and CUDA's on the device: both are faithful to well under one unit in the last place of the result, they are not
 * guaranteed to agree in the last bit, so tests compare the quantized bytes with a one-step allowance. */
void qo_silu_mul(const float *x, const float *gate, float *y, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        const float val = x[i];
        const float silu = val / (1.0f + expf(-val));
        y[i] = silu * gate[i];
    }
}

/* rms_norm(x) * weight, row by row: rms_norm_cpu_f32 (kernels/normalization/rms_norm.cuh:32-58) -- the sum of squares
 * in double, rms = sqrtf((float)(sum / n_cols) + eps), inv_rms = 1.0f / rms, y = x * inv_rms * weight in that order. */
void qo_rms_norm(const float *x, const float *weight, float *y, int64_t n_rows, int64_t n_cols, float eps)
{
    for (int64_t row = 0; row < n_rows; row++) {
        const float *xr = x + row * n_cols;
        float *yr = y + row * n_cols;
        double sum_sq = 0.0;
        for (int64_t i = 0; i < n_cols; i++) sum_sq += (double)xr[i] * xr[i];
        const float rms = sqrtf((float)(sum_sq / (double)n_cols) + eps);
        const float inv_rms = 1.0f / rms;
        for (int64_t i = 0; i < n_cols; i++) yr[i] = xr[i] * inv_rms * weight[i];
    }
}

void qo_quantize_q8_1(const float *x, void *y, int64_t n, unsigned flags)
{
    /* include/quantize.h:165-193 (CPU), :302-337 (GPU kernel),
     * tests/framework/test_framework.cuh:195-225 (S_FROM_QSUM|CLAMP127),
     * python/quant_gemm/csrc/gemm_ops.cu:75-110 (CLAMP127). */
    const int64_t nb = n / 32;
    uint8_t *out = (uint8_t *)y;
    const int lo = (flags & QO_Q81_CLAMP127) ? -127 : -128;
    for (int64_t i = 0; i < nb; i++) {
        const float *src = x + i * 32;
        uint8_t *dst = out + i * 36;
        float amax = 0.0f;
        float sum = 0.0f;
        int sum_q = 0;
        for (int j = 0; j < 32; j++) {
            const float a = fabsf(src[j]);
            if (a > amax) amax = a; /* == std::max(amax, |x|) for non-NaN input */
            sum += src[j];
        }
        if (flags & QO_Q81_TREE_SUM) {
            /* kernels/gemm/gemm_fused.cuh:96-127: s[i] += s[i+16] (i < 16), += s[i+8], += s[i+4], += s[i+2], s[0] + s[1] */
            float t[32];
            for (int j = 0; j < 32; j++) t[j] = src[j];
            for (int w = 16; w >= 1; w >>= 1)
                for (int j = 0; j < w; j++) t[j] = t[j] + t[j + w];
            sum = t[0];
        }
        float d = amax / 127.0f;
        float id = (d > 0) ? 1.0f / d : 0.0f;
        if (flags & QO_Q81_ID_FROM_HALF_D) {
            /* gemm_fused.cuh:131-133: every thread re-reads d from the stored half and inverts THAT */
            const float dh = qo_fp16_to_fp32(qo_fp32_to_fp16(d));
            id = (dh != 0.0f) ? 1.0f / dh : 0.0f;
        }
        if ((flags & QO_Q81_ZERO_D1) && amax == 0.0f) d = 1.0f; /* quantize_q8_1.json: "d = 1.0 if amax == 0"; q stays 0 */
        for (int j = 0; j < 32; j++) {
            const float v = src[j] * id;
            int q = (flags & QO_Q81_ROUND_EVEN) ? rn_even(v) : (int)roundf(v);
            /* gemm_fused.cuh:138-139 narrows to int8_t BEFORE clamping (`int8_t q = (int8_t)roundf(val * id)`; nvcc: F2I to
             * s32, low byte sign-extended).  It matters only where the fp16 d is subnormal and much smaller than the fp32 d
             * (amax below ~1e-3): there |x * id| can pass 127.5 and the reference wraps. */
            if (flags & QO_Q81_ID_FROM_HALF_D) q = (int)(int8_t)(uint8_t)((unsigned)q & 0xffu);
            if (q < lo) q = lo;
            if (q > 127) q = 127;
            dst[4 + j] = (uint8_t)(int8_t)q;
            sum_q += q;
        }
        st16(dst, qo_fp32_to_fp16(d));
        if (flags & QO_Q81_S_FROM_QSUM) {
            st16(dst + 2, qo_fp32_to_fp16((float)sum_q * d)); /* framework:224 `sum_q * scale` */
        } else {
            st16(dst + 2, qo_fp32_to_fp16(sum));
        }
    }
}

void qo_quantize_q8_1_f16(const uint16_t *x_f16, void *y, int64_t n, unsigned flags)
{
    /* gemm_fused.cuh:90-93,137: every element enters as __half2float(fp16_data[tid]) */
    float buf[32];
    for (int64_t i = 0; i + 32 <= n; i += 32) {
        for (int j = 0; j < 32; j++) buf[j] = qo_fp16_to_fp32(x_f16[i + j]);
        qo_quantize_q8_1(buf, (uint8_t *)y + (i / 32) * 36, 32, flags);
    }
}

/* ------------------------------------------------------------------------- */
/* weight quantizers (test-data producers)                                    */
/* ------------------------------------------------------------------------- */
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void qo_quantize_q4_0_ref(const float *x, void *y, int64_t n)
{
    /* include/quantize.h:35-70 */
    const int64_t nb = n / 32;
    uint8_t *out = (uint8_t *)y;
    for (int64_t i = 0; i < nb; i++) {
        const float *src = x + i * 32;
        uint8_t *dst = out + i * 18;
        float amax = 0.0f;
        for (int j = 0; j < 32; j++) {
            const float a = fabsf(src[j]);
            if (a > amax) amax = a;
        }
        const float d = amax / 7.0f;
        const float id = (d > 0) ? 1.0f / d : 0.0f;
        st16(dst, qo_fp32_to_fp16(d));
        for (int j = 0; j < 16; j++) {
            int q0 = (int)roundf(src[j] * id) + 8;
            int q1 = (int)roundf(src[j + 16] * id) + 8;
            q0 = clampi(q0, 0, 15);
            q1 = clampi(q1, 0, 15);
            dst[2 + j] = (uint8_t)((q1 << 4) | (q0 & 0x0f));
        }
    }
}

void qo_quantize_q8_0_ref(const float *x, void *y, int64_t n)
{
    /* include/quantize.h:111-135 */
    const int64_t nb = n / 32;
    uint8_t *out = (uint8_t *)y;
    for (int64_t i = 0; i < nb; i++) {
        const float *src = x + i * 32;
        uint8_t *dst = out + i * 34;
        float amax = 0.0f;
        for (int j = 0; j < 32; j++) {
            const float a = fabsf(src[j]);
            if (a > amax) amax = a;
        }
        const float d = amax / 127.0f;
        const float id = (d > 0) ? 1.0f / d : 0.0f;
        st16(dst, qo_fp32_to_fp16(d));
        for (int j = 0; j < 32; j++) {
            const int q = (int)roundf(src[j] * id);
            dst[2 + j] = (uint8_t)(int8_t)clampi(q, -128, 127);
        }
    }
}

void qo_to_q4_0(const float *x, void *y, int64_t n)
{
    /* tests/framework/test_framework.cuh:162-192: the rounded value goes
     * through int8_t before the [-8,7] clamp. */
    const int64_t nb = n / 32;
    uint8_t *out = (uint8_t *)y;
    for (int64_t b = 0; b < nb; b++) {
        const float *src = x + b * 32;
        uint8_t *dst = out + b * 18;
        float max_abs = 0.0f;
        for (int i = 0; i < 32; i++) {
            const float a = fabsf(src[i]);
            if (a > max_abs) max_abs = a;
        }
        const float scale = max_abs / 7.0f;
        const float inv = (scale > 0) ? (1.0f / scale) : 0.0f;
        st16(dst, qo_fp32_to_fp16(scale));
        for (int i = 0; i < 16; i++) {
            int8_t v0 = (int8_t)roundf(src[i] * inv);
            int8_t v1 = (int8_t)roundf(src[i + 16] * inv);
            v0 = (int8_t)clampi(v0, -8, 7);
            v1 = (int8_t)clampi(v1, -8, 7);
            dst[2 + i] = (uint8_t)(((v0 + 8) & 0x0f) | (((v1 + 8) & 0x0f) << 4));
        }
    }
}

void qo_to_q8_0(const float *x, void *y, int64_t n)
{
    /* tests/framework/test_framework.cuh:228-253 (clamp +-127) */
    const int64_t nb = n / 32;
    uint8_t *out = (uint8_t *)y;
    for (int64_t b = 0; b < nb; b++) {
        const float *src = x + b * 32;
        uint8_t *dst = out + b * 34;
        float max_abs = 0.0f;
        for (int i = 0; i < 32; i++) {
            const float a = fabsf(src[i]);
            if (a > max_abs) max_abs = a;
        }
        const float scale = max_abs / 127.0f;
        const float inv = (scale > 0) ? (1.0f / scale) : 0.0f;
        st16(dst, qo_fp32_to_fp16(scale));
        for (int i = 0; i < 32; i++) {
            int8_t v = (int8_t)roundf(src[i] * inv);
            v = (int8_t)clampi(v, -127, 127);
            dst[2 + i] = (uint8_t)v;
        }
    }
}

static void minmax32(const float *src, float *mn, float *mx)
{
    float lo = src[0], hi = src[0];
    for (int i = 1; i < 32; i++) {
        if (src[i] < lo) lo = src[i];
        if (src[i] > hi) hi = src[i];
    }
    *mn = lo;
    *mx = hi;
}

void qo_to_q4_1(const float *x, void *y, int64_t n)
{
    /* tests/framework/test_framework.cuh:256-288 */
    const int64_t nb = n / 32;
    uint8_t *out = (uint8_t *)y;
    for (int64_t b = 0; b < nb; b++) {
        const float *src = x + b * 32;
        uint8_t *dst = out + b * 20;
        float mn, mx;
        minmax32(src, &mn, &mx);
        const float scale = (mx - mn) / 15.0f;
        const float inv = (scale > 0) ? (1.0f / scale) : 0.0f;
        st16(dst, qo_fp32_to_fp16(scale));
        st16(dst + 2, qo_fp32_to_fp16(mn));
        for (int i = 0; i < 16; i++) {
            int q0 = (int)roundf((src[i] - mn) * inv);
            int q1 = (int)roundf((src[i + 16] - mn) * inv);
            q0 = clampi(q0, 0, 15);
            q1 = clampi(q1, 0, 15);
            dst[4 + i] = (uint8_t)((q1 << 4) | q0);
        }
    }
}

void qo_to_q5_0(const float *x, void *y, int64_t n)
{
    /* tests/framework/test_framework.cuh:291-329 */
    const int64_t nb = n / 32;
    uint8_t *out = (uint8_t *)y;
    for (int64_t b = 0; b < nb; b++) {
        const float *src = x + b * 32;
        uint8_t *dst = out + b * 22;
        float max_abs = 0.0f;
        for (int i = 0; i < 32; i++) {
            const float a = fabsf(src[i]);
            if (a > max_abs) max_abs = a;
        }
        const float scale = max_abs / 15.0f;
        const float inv = (scale > 0) ? (1.0f / scale) : 0.0f;
        uint32_t qh = 0;
        st16(dst, qo_fp32_to_fp16(scale));
        for (int i = 0; i < 16; i++) {
            int q0 = (int)roundf(src[i] * inv) + 16;
            int q1 = (int)roundf(src[i + 16] * inv) + 16;
            q0 = clampi(q0, 0, 31);
            q1 = clampi(q1, 0, 31);
            dst[6 + i] = (uint8_t)(((q1 & 0x0f) << 4) | (q0 & 0x0f));
            qh |= (uint32_t)((q0 >> 4) & 1) << i;
            qh |= (uint32_t)((q1 >> 4) & 1) << (i + 16);
        }
        st32(dst + 2, qh);
    }
}

void qo_to_q5_1(const float *x, void *y, int64_t n)
{
    /* tests/framework/test_framework.cuh:332-367 */
    const int64_t nb = n / 32;
    uint8_t *out = (uint8_t *)y;
    for (int64_t b = 0; b < nb; b++) {
        const float *src = x + b * 32;
        uint8_t *dst = out + b * 24;
        float mn, mx;
        minmax32(src, &mn, &mx);
        const float scale = (mx - mn) / 31.0f;
        const float inv = (scale > 0) ? (1.0f / scale) : 0.0f;
        uint32_t qh = 0;
        st16(dst, qo_fp32_to_fp16(scale));
        st16(dst + 2, qo_fp32_to_fp16(mn));
        for (int i = 0; i < 16; i++) {
            int q0 = (int)roundf((src[i] - mn) * inv);
            int q1 = (int)roundf((src[i + 16] - mn) * inv);
            q0 = clampi(q0, 0, 31);
            q1 = clampi(q1, 0, 31);
            dst[8 + i] = (uint8_t)(((q1 & 0x0f) << 4) | (q0 & 0x0f));
            qh |= (uint32_t)((q0 >> 4) & 1) << i;
            qh |= (uint32_t)((q1 >> 4) & 1) << (i + 16);
        }
        st32(dst + 4, qh);
    }
}

/* ------------------------------------------------------------------------- */
/* dequantize                                                                 */
/* ------------------------------------------------------------------------- */
/* Unpacked integer weights of one block, in element order 0..31, WITHOUT the
 * format offset (0..15, 0..31, or int8): exactly the values that enter sumi
 * (include/gemm_reference.h:199-212, tests/unit/test_gemm_all_quants.cu:45-50,
 * :84-88,:125-130,:166-170,:204-206). */
static void unpack_w(int wtype, const uint8_t *w, int q[32])
{
    switch (wtype) {
    case QO_Q4_0:
    case QO_Q4_1: {
        const uint8_t *qs = w + (wtype == QO_Q4_0 ? 2 : 4);
        for (int i = 0; i < 16; i++) {
            q[i] = qs[i] & 0x0f;
            q[i + 16] = (qs[i] >> 4) & 0x0f;
        }
        break;
    }
    case QO_Q5_0:
    case QO_Q5_1: {
        const uint8_t *qhp = w + (wtype == QO_Q5_0 ? 2 : 4);
        const uint8_t *qs = qhp + 4;
        const uint32_t qh = ld32(qhp);
        for (int i = 0; i < 16; i++) {
            q[i] = (qs[i] & 0x0f) | (int)(((qh >> i) & 1u) << 4);
            q[i + 16] = ((qs[i] >> 4) & 0x0f) | (int)(((qh >> (i + 16)) & 1u) << 4);
        }
        break;
    }
    case QO_Q8_0:
        for (int i = 0; i < 32; i++) q[i] = (int8_t)w[2 + i];
        break;
    default:
        for (int i = 0; i < 32; i++) q[i] = 0;
    }
}

void qo_dequantize(int type, const void *x, float *y, int64_t n)
{
    const int64_t nb = n / 32;
    const size_t bs = qo_block_bytes(type);
    const uint8_t *in = (const uint8_t *)x;
    for (int64_t b = 0; b < nb; b++) {
        const uint8_t *blk = in + b * bs;
        float *dst = y + b * 32;
        const float d = qo_fp16_to_fp32(ld16(blk));
        int q[32];
        if (type == QO_Q8_1) {
            /* quantize.h:198-211 */
            for (int i = 0; i < 32; i++) dst[i] = (float)(int8_t)blk[4 + i] * d;
            continue;
        }
        unpack_w(type, blk, q);
        switch (type) {
        case QO_Q4_0: /* quantize.h:84-102: (q - 8) * d */
            for (int i = 0; i < 32; i++) dst[i] = (float)(q[i] - 8) * d;
            break;
        case QO_Q5_0:
            for (int i = 0; i < 32; i++) dst[i] = (float)(q[i] - 16) * d;
            break;
        case QO_Q4_1:
        case QO_Q5_1: {
            const float m = qo_fp16_to_fp32(ld16(blk + 2));
            for (int i = 0; i < 32; i++) dst[i] = (float)q[i] * d + m;
            break;
        }
        case QO_Q8_0: /* quantize.h:140-153 */
            for (int i = 0; i < 32; i++) dst[i] = (float)q[i] * d;
            break;
        default:
            break;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* block dot products                                                         */
/* ------------------------------------------------------------------------- */
int32_t qo_block_sumi(int wtype, const void *wblock, const void *ablock)
{
    int q[32];
    const int8_t *a = (const int8_t *)ablock + 4;
    int32_t sumi = 0;
    unpack_w(wtype, (const uint8_t *)wblock, q);
    for (int i = 0; i < 32; i++) sumi += q[i] * (int32_t)a[i];
    return sumi;
}

/*
 * One block's float contribution, added into *sum the way the reference does.
 * CPU order (default): every operator rounds separately, in C precedence:
 *   q4_0  sum += d_w * (d_a * sumi - 8.0f * s_a)      gemm_reference.h:216
 *   q5_0  sum += d_w * (d_a * sumi - 16.0f * s_a)     test_gemm_all_quants.cu:134
 *   q4_1  sum += d_w * d_a * sumi + m_w * s_a / 4.0f  test_gemm_all_quants.cu:91
 *   q5_1  same                                        test_gemm_all_quants.cu:176
 *   q8_0  sum += sumi * d_a * d_w                     gemm_reference.h:261
 *         sum += d_w * d_a * sumi  (ASSOC_UNIT)       test_gemm_all_quants.cu:209
 * QO_GEMM_FMA: what nvcc -O3 (default -fmad=true) emits for the reference's
 * GPU kernels kernels/gemm/gemm_quant_formats.cuh:73-334, read from the SASS
 * of that file built for sm_100a:
 *   q4_0  t = fma(d_a, sumi, -(8*s_a));  sum = fma(d_w, t, sum)
 *   q5_0  t = fma(d_a, sumi, -(16*s_a)); sum = fma(d_w, t, sum)
 *   q4_1/q5_1  r = fma(d_w*d_a, sumi, (m_w*s_a)/4); sum = sum + r
 *   q8_0  sum = fma(d_w*d_a, sumi, sum)
 */
static void accum_block(int wtype, const uint8_t *w, const uint8_t *a, unsigned flags, float *sum)
{
    const float d_w = qo_fp16_to_fp32(ld16(w));
    const float d_a = qo_fp16_to_fp32(ld16(a));
    const float s_a = qo_fp16_to_fp32(ld16(a + 2));
    const int32_t sumi = qo_block_sumi(wtype, w, a);
    const float fs = (float)sumi;
    const int fma_mode = (flags & QO_GEMM_FMA) != 0;

    switch (wtype) {
    case QO_Q4_0:
    case QO_Q5_0: {
        const float off = (wtype == QO_Q4_0) ? 8.0f : 16.0f;
        if (fma_mode) {
            const float t = fmaf(d_a, fs, -(off * s_a));
            *sum = fmaf(d_w, t, *sum);
        } else {
            *sum += d_w * (d_a * fs - off * s_a);
        }
        break;
    }
    case QO_Q4_1:
    case QO_Q5_1: {
        const float m_w = qo_fp16_to_fp32(ld16(w + 2));
        const float ms = (flags & QO_GEMM_MS_EXACT) ? (m_w * s_a) : (m_w * s_a / 4.0f);
        if (fma_mode) {
            const float r = fmaf(d_w * d_a, fs, ms);
            *sum = *sum + r;
        } else {
            *sum += d_w * d_a * fs + ms;
        }
        break;
    }
    case QO_Q8_0:
        if (fma_mode) {
            *sum = fmaf(d_w * d_a, fs, *sum);
        } else if (flags & QO_GEMM_Q80_ASSOC_UNIT) {
            *sum += d_w * d_a * fs;
        } else {
            *sum += fs * d_a * d_w;
        }
        break;
    default:
        break;
    }
}

float qo_block_dot(int wtype, const void *wblock, const void *ablock, unsigned flags)
{
    float sum = 0.0f;
    accum_block(wtype, (const uint8_t *)wblock, (const uint8_t *)ablock, flags, &sum);
    return sum;
}

/* ------------------------------------------------------------------------- */
/* GEMM drivers: gemm_reference.h:175-267, test_gemm_all_quants.cu:23-215     */
/* ------------------------------------------------------------------------- */
void qo_gemm(int wtype, const void *act_q8_1, const void *weight, float *C,
             int T, int F, int K, int64_t ldc_t, int64_t ldc_f, unsigned flags,
             int t0, int t1)
{
    const int nb = K / 32;
    const size_t bs = qo_block_bytes(wtype);
    const uint8_t *A = (const uint8_t *)act_q8_1;
    const uint8_t *W = (const uint8_t *)weight;
    (void)T;
    for (int t = t0; t < t1; t++) {
        for (int f = 0; f < F; f++) {
            float sum = 0.0f;
            for (int b = 0; b < nb; b++) {
                accum_block(wtype, W + ((size_t)f * nb + b) * bs,
                            A + ((size_t)t * nb + b) * 36, flags, &sum);
            }
            C[(int64_t)t * ldc_t + (int64_t)f * ldc_f] = sum;
        }
    }
}

void qo_gemm_sumi(int wtype, const void *act_q8_1, const void *weight, int32_t *sumi,
                  int T, int F, int K, int t0, int t1)
{
    const int nb = K / 32;
    const size_t bs = qo_block_bytes(wtype);
    const uint8_t *A = (const uint8_t *)act_q8_1;
    const uint8_t *W = (const uint8_t *)weight;
    (void)T;
    for (int t = t0; t < t1; t++)
        for (int f = 0; f < F; f++)
            for (int b = 0; b < nb; b++)
                sumi[((size_t)t * F + f) * nb + b] =
                    qo_block_sumi(wtype, W + ((size_t)f * nb + b) * bs,
                                  A + ((size_t)t * nb + b) * 36);
}

/* ------------------------------------------------------------------------- */
/* W4A16 / W8A16 (fp32 activations): gemm_reference.h:73-147,                  */
/* gemm_cuda_naive.cuh:66-143.  Next-row groundwork.             */
/* ------------------------------------------------------------------------- */
void qo_gemm_f32act_dequant(int wtype, const float *act, const void *weight, float *C,
                            int T, int F, int K, int64_t ldc_t, int64_t ldc_f, unsigned flags)
{
    const int nb = K / 32;
    const size_t bs = qo_block_bytes(wtype);
    const uint8_t *W = (const uint8_t *)weight;
    const int fma = (flags & QO_GEMM_FMA) != 0;
    for (int t = 0; t < T; t++) {
        const float *a = act + (size_t)t * K;
        for (int f = 0; f < F; f++) {
            float sum = 0.0f;
            for (int b = 0; b < nb; b++) {
                const uint8_t *blk = W + ((size_t)f * nb + b) * bs;
                const float d = qo_fp16_to_fp32(ld16(blk));
                const float *ab = a + (size_t)b * 32;
                if (wtype == QO_Q4_0) {
                    for (int k = 0; k < 16; k++) {
                        const int q0 = (blk[2 + k] & 0x0F) - 8, q1 = (blk[2 + k] >> 4) - 8;
                        const float w0 = (float)q0 * d, w1 = (float)q1 * d;
                        if (fma) { sum = fmaf(ab[k], w0, sum); sum = fmaf(ab[k + 16], w1, sum); }
                        else { sum += ab[k] * w0; sum += ab[k + 16] * w1; }
                    }
                } else { /* Q8_0 */
                    for (int k = 0; k < 32; k++) {
                        const float w = (float)(int8_t)blk[2 + k] * d;
                        if (fma) sum = fmaf(ab[k], w, sum);
                        else sum += ab[k] * w;
                    }
                }
            }
            C[(int64_t)t * ldc_t + (int64_t)f * ldc_f] = sum;
        }
    }
}
