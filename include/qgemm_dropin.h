/*
 * include/qgemm_dropin.h -- glue shared by the drop-in C++ headers: forwards the reference's
 * inline launchers (void, device pointers, cudaStream_t = 0, errors left sticky) to the C ABI.
 *
 * The reference's launchers take no workspace.  The tensor-core path needs one, so these shims
 * pass QGEMM_STREAM_ALLOC: for M >= 96 the library borrows the scratch from the stream's memory
 * pool for the duration of the call (cudaMallocAsync / cudaFreeAsync, stream-ordered).  A caller
 * who prefers to own it registers scratch once per device with qgemm_set_default_workspace().
 */
#ifndef QGEMM_DROPIN_H
#define QGEMM_DROPIN_H

#include <cuda_runtime.h>

#include "qgemm.h"

#include <stdio.h>

/*
 * The reference's launchers return void and surface failures only through the sticky CUDA error.  A qgemm status that
 * is not a CUDA error (bad argument, wrong device, missing workspace) would vanish, leaving C unwritten with nothing to
 * observe: the shims keep the calling thread's last non-zero status (qgemm_dropin_last_status(), cleared on read) and
 * report it once on stderr.  Define QGEMM_DROPIN_ABORT to abort() instead.
 */
static inline int* qgemm_dropin_status_slot(void) {
    static thread_local int status = 0;
    return &status;
}
static inline int qgemm_dropin_last_status(void) {
    int* s = qgemm_dropin_status_slot();
    const int v = *s;
    *s = 0;
    return v;
}
static inline void qgemm_dropin_status(int rc, const char* what) {
    if (rc == 0) return;
    *qgemm_dropin_status_slot() = rc;
    fprintf(stderr, "qgemm: %s failed: %s (%s)\n", what, qgemm_strerror(rc), qgemm_last_error_detail());
#ifdef QGEMM_DROPIN_ABORT
    abort();
#endif
}

/* include/ convention: A = q8_1 activations [M rows], B = weights [N rows], C[M, N] row-major */
static inline void qgemm_dropin_include(int wtype, const void* A, const void* B, float* C, int M, int N, int K,
                                        cudaStream_t stream) {
    qgemm_dropin_status(qgemm_gemm(wtype, A, B, C, M, N, K, (int64_t)N, 1, QGEMM_STREAM_ALLOC, nullptr, 0, (void*)stream), "gemm (include/ convention)");
}
/* kernels/gemm convention: weight [M rows], activation [N tokens], output[m * N + n] */
static inline void qgemm_dropin_ggml(int wtype, const void* weight, const void* activation, float* output, int M, int N,
                                     int K, cudaStream_t stream) {
    qgemm_dropin_status(qgemm_gemm(wtype, activation, weight, output, N, M, K, 1, (int64_t)N, QGEMM_STREAM_ALLOC, nullptr, 0, (void*)stream),
                        "gemm (ggml convention)");
}

#endif /* QGEMM_DROPIN_H */
