/*
 * include/quant_types.h -- the three block types of the include/ API plus their helpers.
 *
 * Drop-in for the reference header (include/quant_types.h:20-181): when GGML_COMMON_DECL is
 * defined (the header sits inside a llama.cpp build) llama.cpp's own block types are used and
 * nothing is declared here; otherwise block_q4_0 / block_q8_0 / block_q8_1 are declared with
 * the reference's names and bytes.  May be combined with compat/ggml_types.h.
 */
#ifndef QUANT_TYPES_H
#define QUANT_TYPES_H

#ifndef GGML_COMMON_DECL

#include <cuda_fp16.h>

#include <cstdint>

#ifndef QK4_0
#define QK4_0 32
#endif
#ifndef QK8_0
#define QK8_0 32
#endif
#ifndef QK8_1
#define QK8_1 32
#endif

#ifndef QGEMM_BLOCK_TYPES
#define QGEMM_BLOCK_TYPES
typedef struct { half d; uint8_t qs[QK4_0 / 2]; } block_q4_0;
typedef struct { half d; int8_t qs[QK8_0]; } block_q8_0;
typedef struct { half2 ds; int8_t qs[QK8_1]; } block_q8_1;
#endif
static_assert(sizeof(block_q4_0) == 18 && sizeof(block_q8_0) == 34 && sizeof(block_q8_1) == 36,
              "llama.cpp block sizes");

/* nibble helpers (reference :128-139) */
__host__ __device__ inline int get_q4_0_low(uint8_t packed) { return packed & 0x0F; }
__host__ __device__ inline int get_q4_0_high(uint8_t packed) { return packed >> 4; }
__host__ __device__ inline uint8_t pack_q4_0(int q0, int q1) { return (uint8_t)((q1 << 4) | (q0 & 0x0F)); }

/* q8_1 scale / sum accessors (reference :142-149) */
__device__ inline float get_q8_1_d(const block_q8_1& b) { return __half2float(__low2half(b.ds)); }
__device__ inline float get_q8_1_s(const block_q8_1& b) { return __half2float(__high2half(b.ds)); }

/* compile-time format facts (reference :156-181) */
template <typename T> struct quant_traits;
template <> struct quant_traits<block_q4_0> {
    static constexpr int block_size = QK4_0;
    static constexpr int bytes_per_block = sizeof(block_q4_0);
    static constexpr float bits_per_element = 4.5f;
    static constexpr bool has_sum = false;
};
template <> struct quant_traits<block_q8_0> {
    static constexpr int block_size = QK8_0;
    static constexpr int bytes_per_block = sizeof(block_q8_0);
    static constexpr float bits_per_element = 8.5f;
    static constexpr bool has_sum = false;
};
template <> struct quant_traits<block_q8_1> {
    static constexpr int block_size = QK8_1;
    static constexpr int bytes_per_block = sizeof(block_q8_1);
    static constexpr float bits_per_element = 9.0f;
    static constexpr bool has_sum = true;
};

#endif /* !GGML_COMMON_DECL */
#endif /* QUANT_TYPES_H */
