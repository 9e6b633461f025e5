// ptx.cuh -- thin inline-PTX wrappers for sm_100a: mbarrier, bulk async copy
// (TMA engine, 1-D), named barriers, tcgen05/TMEM.
#pragma once
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

namespace qgemm {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// A wait that would never end (a protocol bug, a lost copy) must not hang the device: after a bounded number of polls
// the thread traps, which fails the launch with a CUDA error instead.  -DQGEMM_MBAR_DEBUG also prints which barrier
// it was stuck on (the printf call costs every kernel a stack frame, so it is off in the shipped build).
#ifdef QGEMM_MBAR_DEBUG
static __device__ __noinline__ void mbar_timeout(uint64_t* bar, uint32_t parity) {
    printf("qgemm: mbarrier wait timed out: block %d thread %d barrier smem+0x%x parity %u\n", (int)blockIdx.x, (int)threadIdx.x,
           smem_u32(bar), parity);
    __trap();
}
#else
__device__ __forceinline__ void mbar_timeout(uint64_t*, uint32_t) { asm volatile("trap;"); }
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// for single-thread role warps that share a scheduler with busy math warps: sleep between polls
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(32);
}
// Guarded forms (the prefill kernel with its five barrier rings uses these): polls before giving up -- try_wait itself
// suspends the thread for a hardware-defined slice, so this is seconds.  The decode kernels keep the bare loops: the
// poll counter cost them 3 % on the decode stack (A/B on one box).
constexpr uint32_t kMbarTimeoutPolls = 1u << 26;
__device__ __forceinline__ void mbar_wait_guarded(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    for (uint32_t n = 0; !mbar_try_wait(bar, parity); n++)
        if (n > kMbarTimeoutPolls) mbar_timeout(bar, parity);
}
__device__ __forceinline__ void mbar_wait_backoff_guarded(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    for (uint32_t n = 0; !mbar_try_wait(bar, parity); n++) {
        __nanosleep(32);
        if (n > (kMbarTimeoutPolls >> 2)) mbar_timeout(bar, parity);
    }
}

// ---- 1-D bulk async copy global -> shared (TMA engine, no tensor map) ---------
// dst/src 16-byte aligned, bytes % 16 == 0; completion = complete_tx on `bar`.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// Same with an L2 evict-first policy: weights are streamed once per launch.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// L2 prefetch of a 16-byte aligned range (no smem, no completion tracking)
// shared -> global bulk copy (TMA store engine), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the sources of all committed groups have been read (shared memory may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all committed groups are complete (the writes are done)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

// ---- programmatic dependent launch (PDL) -----------------------------------------
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- named barriers -----------------------------------------------------------
__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace ptx
}  // namespace qgemm
