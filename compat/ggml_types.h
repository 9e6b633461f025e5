/*
 * compat/ggml_types.h -- llama.cpp block formats for the sm_100a build.
 *
 * Drop-in for the reference header of the same path (compat/ggml_types.h:32-299): same type
 * names, the same bytes (18/20/22/24/34/36, checked below), the same QK*, QuantType numbering
 * (= ggml_type), size helpers and CUDA check macros, so code written against the reference
 * compiles unchanged.  Unlike the reference, this header and include/quant_types.h may be
 * included together (shared guard QGEMM_BLOCK_TYPES).
 */
#ifndef COMPAT_GGML_TYPES_H
#define COMPAT_GGML_TYPES_H

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#ifndef QK4_0
#define QK4_0 32
#endif
#define QK4_1 32
#define QK5_0 32
#define QK5_1 32
#ifndef QK8_0
#define QK8_0 32
#endif
#ifndef QK8_1
#define QK8_1 32
#endif
#define QK_K 256 /* K-quant super-block (types listed in QuantType, no kernels here) */

#ifndef QGEMM_BLOCK_TYPES
#define QGEMM_BLOCK_TYPES
/* x = (q - 8) * d                    q in [0,15], two per byte: low nibble = element i, high = i + 16 */
typedef struct { half d; uint8_t qs[QK4_0 / 2]; } block_q4_0;
/* x = q * d; activations' companion for weights */
typedef struct { half d; int8_t qs[QK8_0]; } block_q8_0;
/* x = q * d, ds = (d, s = sum of the 32 original values): the activation format */
typedef struct { half2 ds; int8_t qs[QK8_1]; } block_q8_1;
#endif
/* x = q * d + m */
typedef struct { half d; half m; uint8_t qs[QK4_1 / 2]; } block_q4_1;
/* x = (q - 16) * d, q = nibble | (bit i of qh) << 4 */
typedef struct { half d; uint8_t qh[4]; uint8_t qs[QK5_0 / 2]; } block_q5_0;
/* x = q * d + m, 5-bit q as above */
typedef struct { half d; half m; uint8_t qh[4]; uint8_t qs[QK5_1 / 2]; } block_q5_1;

static_assert(sizeof(block_q4_0) == 18, "block_q4_0 must be 18 bytes");
static_assert(sizeof(block_q4_1) == 20, "block_q4_1 must be 20 bytes");
static_assert(sizeof(block_q5_0) == 22, "block_q5_0 must be 22 bytes");
static_assert(sizeof(block_q5_1) == 24, "block_q5_1 must be 24 bytes");
static_assert(sizeof(block_q8_0) == 34, "block_q8_0 must be 34 bytes");
static_assert(sizeof(block_q8_1) == 36, "block_q8_1 must be 36 bytes");

enum QuantType {
    QUANT_TYPE_F32 = 0, QUANT_TYPE_F16 = 1,
    QUANT_TYPE_Q4_0 = 2, QUANT_TYPE_Q4_1 = 3, QUANT_TYPE_Q5_0 = 6, QUANT_TYPE_Q5_1 = 7,
    QUANT_TYPE_Q8_0 = 8, QUANT_TYPE_Q8_1 = 9,
    QUANT_TYPE_Q2_K = 10, QUANT_TYPE_Q3_K = 11, QUANT_TYPE_Q4_K = 12, QUANT_TYPE_Q5_K = 13,
    QUANT_TYPE_Q6_K = 14, QUANT_TYPE_Q8_K = 15,
};

__host__ __device__ inline int get_block_size(QuantType type) {
    if (type >= QUANT_TYPE_Q2_K && type <= QUANT_TYPE_Q8_K) return QK_K;
    if (type >= QUANT_TYPE_Q4_0 && type <= QUANT_TYPE_Q8_1 && type != 4 && type != 5) return 32;
    return 1;
}

__host__ __device__ inline int get_block_bytes(QuantType type) {
    switch (type) {
    case QUANT_TYPE_Q4_0: return (int)sizeof(block_q4_0);
    case QUANT_TYPE_Q4_1: return (int)sizeof(block_q4_1);
    case QUANT_TYPE_Q5_0: return (int)sizeof(block_q5_0);
    case QUANT_TYPE_Q5_1: return (int)sizeof(block_q5_1);
    case QUANT_TYPE_Q8_0: return (int)sizeof(block_q8_0);
    case QUANT_TYPE_Q8_1: return (int)sizeof(block_q8_1);
    default: return 0;
    }
}

inline const char* get_type_name(QuantType type) {
    static const char* const names[] = {"F32", "F16", "Q4_0", "Q4_1", nullptr, nullptr, "Q5_0", "Q5_1", "Q8_0", "Q8_1"};
    return (type >= 0 && type <= QUANT_TYPE_Q8_1 && names[type]) ? names[type] : "Unknown";
}

#ifndef CUDA_CHECK
#define CUDA_CHECK(call)                                                                              \
    do {                                                                                              \
        cudaError_t qgemm_err_ = (call);                                                              \
        if (qgemm_err_ != cudaSuccess) {                                                              \
            fprintf(stderr, "CUDA error at %s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(qgemm_err_)); \
            exit(EXIT_FAILURE);                                                                       \
        }                                                                                             \
    } while (0)
#endif

#define KERNEL_CHECK()                                                                   \
    do {                                                                                 \
        cudaError_t qgemm_err_ = cudaGetLastError();                                     \
        if (qgemm_err_ != cudaSuccess) {                                                 \
            fprintf(stderr, "Kernel error: %s\n", cudaGetErrorString(qgemm_err_));      \
            exit(EXIT_FAILURE);                                                          \
        }                                                                                \
        cudaDeviceSynchronize();                                                         \
    } while (0)

#endif /* COMPAT_GGML_TYPES_H */
