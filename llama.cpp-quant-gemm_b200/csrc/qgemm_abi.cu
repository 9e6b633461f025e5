// qgemm_abi.cu -- the extern "C" surface of libqgemm_sm100.so (include/qgemm.h):
// argument validation, device check, path selection, launch bookkeeping.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "qgemm_common.cuh"

namespace qgemm {

// kernels (defined in the other translation units)
cudaError_t launch_quantize_q8_1(const float*, void*, int64_t nblocks, uint32_t flags, cudaStream_t);
cudaError_t launch_quantize_q8_1_f16(const void*, void*, int64_t nblocks, uint32_t flags, cudaStream_t);
cudaError_t launch_quantize_weight(int wtype, const float*, void*, int64_t nblocks, uint32_t flags, cudaStream_t);
cudaError_t launch_dequantize(int type, const void*, float*, int64_t nblocks, cudaStream_t);
cudaError_t launch_gemm_sequential(int wtype, const void* act, const void* wgt, float* C, int T, int F, int K,
                                   int64_t ldc_t, int64_t ldc_f, uint32_t flags, cudaStream_t);
cudaError_t launch_sumi_generic(int wtype, const void* act, const void* wgt, int32_t* out, int T, int F, int K,
                                cudaStream_t);
bool gemv_supported(int wtype, const void* act, const void* wgt, int F, int K);
cudaError_t launch_gemv(int wtype, const void* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                        int64_t ldc_f, uint32_t flags, int num_sms, cudaStream_t, const PeerOut* peer = nullptr,
                        const void* pf_ptr = nullptr, size_t pf_bytes = 0);
struct GemvGroup { int nmat; const void* wgt[8]; float* C[8]; int F[8]; };
cudaError_t launch_gemv(int wtype, const void* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                        int64_t ldc_f, uint32_t flags, int num_sms, cudaStream_t, const PeerOut* peer, const void* pf_ptr,
                        size_t pf_bytes, const GemvGroup* group);
struct ChainStepHost {
    const void* act; const float* x; const float* gate;
    int nmat; const void* wgt[3]; float* C[3]; int F[3];
    int K; int ldc_f; int wait;
};
int gemv_chain_max_steps();
size_t gemv_chain_sync_bytes(int nsteps);
cudaError_t launch_gemv_chain(int wtype, const ChainStepHost* steps, int nsteps, uint32_t flags, unsigned* sync, int num_sms,
                              cudaStream_t st, const void* pf_ptr, size_t pf_bytes);
bool gemv_mma_supported(int wtype, const void* act, const void* wgt, int T, int F, int K);
cudaError_t launch_gemv_mma(int wtype, const void* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                            int64_t ldc_f, uint32_t flags, int num_sms, cudaStream_t, const PeerOut* peer = nullptr);
size_t mmq_prepack_bytes(int wtype, int F, int K);
cudaError_t launch_mmq_prepack(int wtype, const void* wgt, void* packed, int F, int K, cudaStream_t);
bool f32act_supported(int wtype, const void* act, const void* wgt, int K);
cudaError_t launch_gemm_f32act_dequant(int wtype, const float* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                                       int64_t ldc_f, int num_sms, cudaStream_t st);
cudaError_t launch_gemm_f32act_sequential(int wtype, const float* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                                          int64_t ldc_f, cudaStream_t st);
bool mmq_supported(int wtype, const void* act, const void* wgt, int T, int F, int K);
size_t mmq_workspace_bytes(int wtype, int T, int F, int K);
size_t mmq_workspace_need(int wtype, const void* wgt, int T, int F, int K, uint32_t flags);
cudaError_t launch_quantize_q8_1_silu_mul(const float* x, const float* gate, void* y, int64_t nblocks, uint32_t flags, cudaStream_t st);
cudaError_t launch_quantize_q8_1_rms_norm(const float* x, const float* weight, float* inv_rms, void* y, int64_t rows, int K, float eps,
                                          uint32_t flags, cudaStream_t st);
cudaError_t launch_mmq_f32act(int wtype, const float* act_f32, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                              int64_t ldc_f, uint32_t flags, uint32_t qflags, void* ws, size_t ws_bytes, int num_sms, cudaStream_t st, const float* gate = nullptr);
cudaError_t launch_mmq(int wtype, const void* act, const void* wgt, float* C, int32_t* sumi_out, int T, int F, int K,
                       int64_t ldc_t, int64_t ldc_f, uint32_t flags, void* ws, size_t ws_bytes, int num_sms,
                       cudaStream_t, const PeerOut* peer = nullptr);

constexpr int kMmaMinTokens = 2;   // dp4a GEMV for a single token, mma.sync skinny path from here (measured crossover)
constexpr int kMmqMinTokens = 96;  // AUTO always takes the tcgen05 path from here on (a call without scratch is refused)

// Below kMmqMinTokens AUTO takes the tcgen05 path only when scratch is at hand, from the measured crossover against the
// mma.sync passes (profiles/r02_small_batch_crossover.md).  Up to 64 tokens the kernel runs weight-major (tokens on the N
// side of the MMA, mmq_native.cu), so its time shrinks with the batch: 11008 x 4096 T = 48 40.4 -> 38.4 us, T = 64
// 49.6 -> 39.3 us; 4096 x 4096 T = 64 28.0 -> 24.8 us.  Rows of other lengths than 4096 / 8192 have no wide mma.sync
// variant and cross over at T = 16 (4096 x 11008: T = 16 31.9 -> 29.1 us, T = 32 61 -> 30 us, T = 95 183 -> 46 us).
bool mmq_native_supported(int wtype, const void* wgt, int T, int F, int K);
static int mmq_min_tokens(int wtype, const void* wgt, int T, int F, int K) {
    if (!mmq_native_supported(wtype, wgt, T, F, K)) return kMmqMinTokens;
    (void)F;
    return (K == 4096 || K == 8192) ? 48 : 16;
}

static std::atomic<int64_t> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local uint32_t t_last_path = 0;
static thread_local const void* t_pf_ptr = nullptr;  // qgemm_hint_next_weights (one-shot)
static thread_local size_t t_pf_bytes = 0;
static thread_local char t_detail[256] = "";

static int cuda_fail(cudaError_t e, const char* where) {
    snprintf(t_detail, sizeof(t_detail), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    return QGEMM_E_CUDA;
}

struct DeviceInfo {
    int cc_major = -1, cc_minor = -1, sms = 0;
    bool ok = false;
};
static DeviceInfo g_dev[64];
struct DefaultWs { void* ptr = nullptr; size_t bytes = 0; };
static DefaultWs g_default_ws[64];
static std::mutex g_dev_mu;

// 0 on success; fills *info for the current device.
static int device_check(DeviceInfo* info) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return QGEMM_E_CUDA;
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        DeviceInfo& d = g_dev[dev];
        if (d.cc_major < 0) {
            if (cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
                cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess ||
                cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
                d.cc_major = -1;
                return QGEMM_E_CUDA;
            }
            d.ok = (d.cc_major == 10);  // built for sm_100a only: no PTX fallback, no other arch
        }
        *info = d;
    }
    return info->ok ? QGEMM_OK : QGEMM_E_ARCH;
}

static bool is_weight_type(int t) {
    return t == QGEMM_TYPE_Q4_0 || t == QGEMM_TYPE_Q4_1 || t == QGEMM_TYPE_Q5_0 || t == QGEMM_TYPE_Q5_1 ||
           t == QGEMM_TYPE_Q8_0;
}
static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

cudaError_t smem_optin(const void* kernel, size_t smem) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> done;   // (device, kernel) -> bytes already granted
    int d = 0;
    cudaError_t e = cudaGetDevice(&d);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    size_t& cur = done[{d, kernel}];
    if (smem > cur) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cur = smem;
    }
    return cudaSuccess;
}

// QGEMM_STREAM_ALLOC scratch comes from a pool of our own, one per device, that keeps its pages between
// calls (release threshold = max).  The device's default pool trims at every synchronisation, which made
// each call pay the physical allocation again (measured: +1.1 ms for a 270 MB scratch).
static cudaMemPool_t g_scratch_pool[64];
static cudaError_t scratch_alloc(void** ptr, size_t bytes, cudaStream_t st) {
    int d = 0;
    cudaError_t e = cudaGetDevice(&d);
    if (e != cudaSuccess) return e;
    if (d < 0 || d >= 64) return cudaErrorInvalidDevice;
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        if (!g_scratch_pool[d]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = d;
            cudaMemPool_t pool = nullptr;
            e = cudaMemPoolCreate(&pool, &props);
            if (e != cudaSuccess) return e;
            uint64_t keep = ~0ull;
            (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            g_scratch_pool[d] = pool;
        }
    }
    return cudaMallocFromPoolAsync(ptr, bytes, g_scratch_pool[d], st);
}

static int check_gemm_args(int wtype, const void* act, const void* wgt, const void* out, int T, int F, int K) {
    if (!is_weight_type(wtype) || T < 0 || F < 0 || K < 0 || (K % kQK) != 0) return QGEMM_E_BADARG;
    if (T == 0 || F == 0) return QGEMM_OK;
    if (!act || !wgt || !out) return QGEMM_E_BADARG;
    if (!aligned(act, 4) || !aligned(wgt, 2) || !aligned(out, 4)) return QGEMM_E_ALIGN;
    return QGEMM_OK;
}

__global__ void peer_step_advance_kernel(uint32_t* step) { *step = *step + 1u; }
// one thread: block the stream until `li` launches of the current step have landed here from every rank
// (li = launches_per_step: the whole step)
__global__ void peer_wait_kernel(const uint32_t* flag, const uint32_t* step, uint32_t lps, uint32_t li, uint32_t world) {
    const uint32_t target = ((*step) * lps + li) * world;
    while ((int32_t)(ld_acquire_sys(flag) - target) < 0) __nanosleep(64);
}

__global__ void fill_zero_strided(float* C, int T, int F, int64_t ldc_t, int64_t ldc_f) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)T * F) return;
    C[(i / F) * ldc_t + (i % F) * ldc_f] = 0.0f;
}

static int run_gemm(int wtype, const void* act, const void* wgt, float* C, int T, int F, int K, int64_t ldc_t,
                    int64_t ldc_f, uint32_t flags, void* ws, size_t ws_bytes, cudaStream_t st,
                    const DeviceInfo& dev) {
    if (K == 0) {  // empty contraction: the reference's loop leaves sum = 0.0f
        const int64_t n = (int64_t)T * F;
        fill_zero_strided<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(C, T, F, ldc_t, ldc_f);
        note_launch();
        t_last_path = QGEMM_PATH_GENERIC;
        return cudaGetLastError() == cudaSuccess ? QGEMM_OK : QGEMM_E_CUDA;
    }
    if (!ws) {  // caller registered scratch for the signature-compatible shims
        int d = 0;
        if (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) {
            std::lock_guard<std::mutex> lk(g_dev_mu);
            ws = g_default_ws[d].ptr;
            ws_bytes = g_default_ws[d].bytes;
        }
    }
    uint32_t path = flags & QGEMM_PATH_MASK;
    if (flags & QGEMM_SEQUENTIAL) path = QGEMM_PATH_GENERIC;
    // no scratch from the caller: borrow it from the stream's pool for the duration of this call
    void* pool_ws = nullptr;
    const int tc_min = (flags & QGEMM_WEIGHTS_PREPACKED) ? kMmqMinTokens : mmq_min_tokens(wtype, wgt, T, F, K);
    if (!ws && (flags & QGEMM_STREAM_ALLOC) && (path == QGEMM_PATH_TCGEN05 || (path == QGEMM_PATH_AUTO && T >= tc_min)) &&
        mmq_supported(wtype, act, wgt, T, F, K)) {
        // what the call needs, or what it can use (the split-K scratch of small-T calls), whichever is larger
        const size_t need = align_up(std::max(mmq_workspace_need(wtype, wgt, T, F, K, flags), mmq_workspace_bytes(wtype, T, F, K)), 256);
        if (scratch_alloc(&pool_ws, need, st) == cudaSuccess) {
            ws = pool_ws;
            ws_bytes = need;
        } else {
            (void)cudaGetLastError();  // no pool on this device/driver: the other paths need no scratch
            pool_ws = nullptr;
        }
    }
    struct PoolFree {
        void* p; cudaStream_t s;
        ~PoolFree() { if (p) cudaFreeAsync(p, s); }
    } pool_free{pool_ws, st};
    if (flags & QGEMM_WEIGHTS_PREPACKED) {  // only the tensor-core path reads the packed layout
        if (path != QGEMM_PATH_AUTO && path != QGEMM_PATH_TCGEN05) return QGEMM_E_BADARG;
        path = QGEMM_PATH_TCGEN05;
    }
    if (path == QGEMM_PATH_AUTO) {
        const bool have_ws = ws && ws_bytes >= mmq_workspace_need(wtype, wgt, T, F, K, flags);
        const bool tc_size = T >= (have_ws ? tc_min : kMmqMinTokens) && mmq_supported(wtype, act, wgt, T, F, K);
        // a prefill-sized call without scratch would fall to the weight-streaming passes at a fraction of the speed:
        // say so instead of doing it silently (QGEMM_PATH_MMA asks for that path explicitly, QGEMM_STREAM_ALLOC lends scratch)
        if (tc_size && !have_ws) return QGEMM_E_WORKSPACE;
        if (tc_size)
            path = QGEMM_PATH_TCGEN05;
        else if (T >= (K > 8192 ? kMmaMinTokens + 1 : kMmaMinTokens) && gemv_mma_supported(wtype, act, wgt, T, F, K))
            path = QGEMM_PATH_MMA;
        else if (gemv_supported(wtype, act, wgt, F, K))
            path = QGEMM_PATH_GEMV;
        else
            path = QGEMM_PATH_GENERIC;
    }
    cudaError_t e;
    switch (path) {
    case QGEMM_PATH_GEMV:
        if (!gemv_supported(wtype, act, wgt, F, K)) return QGEMM_E_ALIGN;
        e = launch_gemv(wtype, act, wgt, C, T, F, K, ldc_t, ldc_f, flags, dev.sms, st, nullptr, t_pf_ptr, t_pf_bytes);
        t_pf_ptr = nullptr;
        t_pf_bytes = 0;
        break;
    case QGEMM_PATH_MMA:
        if (!gemv_mma_supported(wtype, act, wgt, T, F, K)) return QGEMM_E_ALIGN;
        e = launch_gemv_mma(wtype, act, wgt, C, T, F, K, ldc_t, ldc_f, flags, dev.sms, st);
        break;
    case QGEMM_PATH_TCGEN05:
        if (!mmq_supported(wtype, act, wgt, T, F, K)) return QGEMM_E_ALIGN;
        if (!ws || ws_bytes < mmq_workspace_need(wtype, wgt, T, F, K, flags)) return QGEMM_E_WORKSPACE;
        e = launch_mmq(wtype, act, wgt, C, nullptr, T, F, K, ldc_t, ldc_f, flags, ws, ws_bytes, dev.sms, st);
        break;
    case QGEMM_PATH_GENERIC:
        e = launch_gemm_sequential(wtype, act, wgt, C, T, F, K, ldc_t, ldc_f, flags, st);
        break;
    default:
        return QGEMM_E_BADARG;
    }
    t_last_path = path;
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "gemm launch");
}

}  // namespace qgemm

using namespace qgemm;

extern "C" {

int qgemm_version(void) { return QGEMM_VERSION; }

const char* qgemm_strerror(int code) {
    switch (code) {
    case QGEMM_OK: return "ok";
    case QGEMM_E_BADARG: return "bad argument (null pointer, K % 32 != 0, unknown type or negative size)";
    case QGEMM_E_ALIGN: return "pointer alignment below the format's minimum, or the forced path cannot take this layout";
    case QGEMM_E_ARCH: return "current CUDA device is not sm_100 (B200); this library has no other code path";
    case QGEMM_E_CUDA: return "a CUDA runtime call failed (see cudaGetLastError)";
    case QGEMM_E_WORKSPACE: return "workspace missing or smaller than qgemm_workspace_bytes()";
    default: return "unknown qgemm error code";
    }
}

size_t qgemm_block_bytes(int type) { return (size_t)block_bytes(type); }
int64_t qgemm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void qgemm_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }
uint32_t qgemm_last_path(void) { return t_last_path; }
const char* qgemm_last_error_detail(void) { return t_detail; }

int qgemm_quantize_q8_1_silu_mul(const float* x, const float* gate, void* y, int64_t rows, int64_t K, uint32_t flags, void* stream) {
    if (rows < 0 || K < 0 || (K % kQK) != 0) return QGEMM_E_BADARG;
    if (rows == 0 || K == 0) return QGEMM_OK;
    if (!x || !gate || !y) return QGEMM_E_BADARG;
    if (!aligned(x, 4) || !aligned(gate, 4) || !aligned(y, 4)) return QGEMM_E_ALIGN;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    const cudaError_t e = launch_quantize_q8_1_silu_mul(x, gate, y, rows * (K / kQK), flags, (cudaStream_t)stream);
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "quantize_q8_1_silu_mul launch");
}

int qgemm_quantize_q8_1_rms_norm(const float* x, const float* weight, void* y, int64_t rows, int64_t K, float eps, uint32_t flags,
                                 float* row_scratch, void* stream) {
    if (rows < 0 || K < 0 || (K % kQK) != 0 || K > INT32_MAX || rows > INT32_MAX) return QGEMM_E_BADARG;
    if (rows == 0 || K == 0) return QGEMM_OK;
    if (!x || !weight || !y) return QGEMM_E_BADARG;
    if (!row_scratch) return QGEMM_E_WORKSPACE;
    if (!aligned(x, 4) || !aligned(weight, 16) || !aligned(y, 4) || !aligned(row_scratch, 4)) return QGEMM_E_ALIGN;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    const cudaError_t e = launch_quantize_q8_1_rms_norm(x, weight, row_scratch, y, rows, (int)K, eps, flags, (cudaStream_t)stream);
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "quantize_q8_1_rms_norm launch");
}

int qgemm_quantize_q8_1(const float* x, void* y, int64_t rows, int64_t K, uint32_t flags, void* stream) {
    if (rows < 0 || K < 0 || (K % kQK) != 0) return QGEMM_E_BADARG;
    if (rows == 0 || K == 0) return QGEMM_OK;
    if (!x || !y) return QGEMM_E_BADARG;
    if (!aligned(x, 4) || !aligned(y, 4)) return QGEMM_E_ALIGN;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    const cudaError_t e = launch_quantize_q8_1(x, y, rows * (K / kQK), flags, (cudaStream_t)stream);
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "quantize_q8_1 launch");
}

int qgemm_quantize_q8_1_f16(const void* x_f16, void* y, int64_t rows, int64_t K, uint32_t flags, void* stream) {
    if (rows < 0 || K < 0 || (K % kQK) != 0) return QGEMM_E_BADARG;
    if (rows == 0 || K == 0) return QGEMM_OK;
    if (!x_f16 || !y) return QGEMM_E_BADARG;
    if (!aligned(x_f16, 2) || !aligned(y, 4)) return QGEMM_E_ALIGN;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    const cudaError_t e = launch_quantize_q8_1_f16(x_f16, y, rows * (K / kQK), flags, (cudaStream_t)stream);
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "quantize_q8_1_f16 launch");
}

int qgemm_quantize_weight(int wtype, const float* x, void* y, int64_t rows, int64_t K, uint32_t flags, void* stream) {
    if (!is_weight_type(wtype) || rows < 0 || K < 0 || (K % kQK) != 0) return QGEMM_E_BADARG;
    if (rows == 0 || K == 0) return QGEMM_OK;
    if (!x || !y) return QGEMM_E_BADARG;
    if (!aligned(x, 4) || !aligned(y, 2)) return QGEMM_E_ALIGN;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    const cudaError_t e = launch_quantize_weight(wtype, x, y, rows * (K / kQK), flags, (cudaStream_t)stream);
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "quantize_weight launch");
}

int qgemm_dequantize(int type, const void* x, float* y, int64_t rows, int64_t K, void* stream) {
    if (!(is_weight_type(type) || type == QGEMM_TYPE_Q8_1) || rows < 0 || K < 0 || (K % kQK) != 0)
        return QGEMM_E_BADARG;
    if (rows == 0 || K == 0) return QGEMM_OK;
    if (!x || !y) return QGEMM_E_BADARG;
    if (!aligned(x, 2) || !aligned(y, 16)) return QGEMM_E_ALIGN;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    const cudaError_t e = launch_dequantize(type, x, y, rows * (K / kQK), (cudaStream_t)stream);
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "dequantize launch");
}

size_t qgemm_workspace_bytes(int wtype, int T, int F, int K, uint32_t flags) {
    if (!is_weight_type(wtype) || T <= 0 || F <= 0 || K <= 0 || (K % kQK) != 0) return 0;
    // [ q8_1 copy of A for qgemm_gemm_f32act | tensor-core path scratch (only where that path can run) ]
    const size_t a_q = align_up((size_t)T * (K / kQK) * kQ81Bytes, 256);
    const uint32_t path = flags & QGEMM_PATH_MASK;
    const bool mmq = (path == QGEMM_PATH_TCGEN05 || (flags & QGEMM_WEIGHTS_PREPACKED) ||
                      (path == QGEMM_PATH_AUTO && T >= mmq_min_tokens(wtype, nullptr, T, F, K))) &&
                     !(flags & QGEMM_SEQUENTIAL);
    return a_q + (mmq ? align_up(mmq_workspace_bytes(wtype, T, F, K), 256) : 0);
}

int qgemm_set_default_workspace(void* workspace, size_t workspace_bytes) {
    if ((workspace == nullptr) != (workspace_bytes == 0) || !aligned(workspace, 256)) return QGEMM_E_BADARG;
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return QGEMM_E_CUDA;
    std::lock_guard<std::mutex> lk(g_dev_mu);
    g_default_ws[d].ptr = workspace;
    g_default_ws[d].bytes = workspace_bytes;
    return QGEMM_OK;
}

int qgemm_gemm_group(int wtype, const void* act_q8_1, int nmat, const void* const* weights, float* const* Cs, const int* Fs,
                     int T, int K, int64_t ldc_t, int64_t ldc_f, uint32_t flags, void* stream) {
    if (nmat < 1 || nmat > 8 || !weights || !Cs || !Fs) return QGEMM_E_BADARG;
    GemvGroup g{};
    g.nmat = nmat;
    int Ftot = 0;
    for (int m = 0; m < nmat; m++) {
        if (int rc = check_gemm_args(wtype, act_q8_1, weights[m], Cs[m], T, Fs[m], K)) return rc;
        if (Fs[m] < 1) return QGEMM_E_BADARG;
        g.wgt[m] = weights[m]; g.C[m] = Cs[m]; g.F[m] = Fs[m];
        Ftot += Fs[m];
    }
    if (T < 1 || K < 32) return QGEMM_E_BADARG;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    bool fast = true;
    for (int m = 0; m < nmat; m++) fast = fast && gemv_supported(wtype, act_q8_1, weights[m], Fs[m], K);
    if (!fast || T > 8) {  // same results, one launch per matrix
        for (int m = 0; m < nmat; m++)
            if (int rc = run_gemm(wtype, act_q8_1, weights[m], Cs[m], T, Fs[m], K, ldc_t, ldc_f, flags | QGEMM_STREAM_ALLOC, nullptr, 0, st, dev))
                return rc;
        return QGEMM_OK;
    }
    cudaError_t e = launch_gemv(wtype, act_q8_1, nullptr, nullptr, T, Ftot, K, ldc_t, ldc_f, flags, dev.sms, st, nullptr,
                                t_pf_ptr, t_pf_bytes, &g);
    t_pf_ptr = nullptr;
    t_pf_bytes = 0;
    t_last_path = QGEMM_PATH_GEMV;
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "gemm_group launch");
}

size_t qgemm_gemv_chain_sync_bytes(int nsteps) { return nsteps < 1 ? 0 : gemv_chain_sync_bytes(nsteps); }
int qgemm_gemv_chain_max_steps(void) { return gemv_chain_max_steps(); }

int qgemm_gemv_chain(int wtype, const qgemm_chain_step* steps, int nsteps, uint32_t flags, void* sync, size_t sync_bytes,
                     void* stream) {
    if (!steps || nsteps < 1 || !is_weight_type(wtype)) return QGEMM_E_BADARG;
    if (!sync || sync_bytes < qgemm_gemv_chain_sync_bytes(nsteps)) return QGEMM_E_WORKSPACE;
    if (!aligned(sync, 4)) return QGEMM_E_ALIGN;
    for (int k = 0; k < nsteps; k++) {
        const qgemm_chain_step& s = steps[k];
        if (s.nmat < 1 || s.nmat > QGEMM_CHAIN_MAX_MATS || s.K < 32 || (s.K % kQK) != 0) return QGEMM_E_BADARG;
        if (!s.act_q8_1 && !s.act_f32) return QGEMM_E_BADARG;
        if (s.act_q8_1 && (s.act_f32 || s.gate_f32)) return QGEMM_E_BADARG;
        if (s.flags & ~QGEMM_INPUTS_READY) return QGEMM_E_BADARG;
        if (s.ldc_f < 1 || s.ldc_f > 0x7fffffff) return QGEMM_E_BADARG;
        for (int m = 0; m < s.nmat; m++) {
            if (!s.weights[m] || !s.C[m] || s.F[m] < 1) return QGEMM_E_BADARG;
            if (!aligned(s.weights[m], 2) || !aligned(s.C[m], 4)) return QGEMM_E_ALIGN;
        }
        if (s.act_q8_1 ? !aligned(s.act_q8_1, 4) : (!aligned(s.act_f32, 4) || !aligned(s.gate_f32, 4))) return QGEMM_E_ALIGN;
    }
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const void* pf_ptr = t_pf_ptr;
    const size_t pf_bytes = t_pf_bytes;
    t_pf_ptr = nullptr;
    t_pf_bytes = 0;
    if (nsteps <= gemv_chain_max_steps() && !(flags & (QGEMM_SEQUENTIAL | QGEMM_PATH_MASK) & ~QGEMM_PATH_GEMV)) {
        std::vector<ChainStepHost> hs((size_t)nsteps);
        for (int k = 0; k < nsteps; k++) {
            const qgemm_chain_step& s = steps[k];
            ChainStepHost& h = hs[(size_t)k];
            h = ChainStepHost{};
            h.act = s.act_q8_1; h.x = s.act_f32; h.gate = s.gate_f32;
            h.nmat = s.nmat;
            for (int m = 0; m < s.nmat; m++) { h.wgt[m] = s.weights[m]; h.C[m] = s.C[m]; h.F[m] = s.F[m]; }
            h.K = s.K; h.ldc_f = (int)s.ldc_f; h.wait = (s.flags & QGEMM_INPUTS_READY) ? 0 : 1;
        }
        cudaError_t e = launch_gemv_chain(wtype, hs.data(), nsteps, flags, (unsigned*)sync, dev.sms, st, pf_ptr, pf_bytes);
        if (e == cudaSuccess) {
            t_last_path = QGEMM_PATH_GEMV | QGEMM_PATH_CHAINED;
            return QGEMM_OK;
        }
        if (e != cudaErrorNotSupported) return cuda_fail(e, "gemv_chain launch");
    }
    // one launch per step (and a quantize launch in front of a step that brings fp32 activations): same results
    for (int k = 0; k < nsteps; k++) {
        const qgemm_chain_step& s = steps[k];
        const void* act = s.act_q8_1;
        void* tmp = nullptr;
        if (!act) {
            const size_t bytes = (size_t)(s.K / kQK) * kQ81Bytes;
            if (cudaError_t e = cudaMallocAsync(&tmp, bytes, st)) return cuda_fail(e, "gemv_chain scratch");
            cudaError_t e = s.gate_f32 ? launch_quantize_q8_1_silu_mul(s.act_f32, s.gate_f32, tmp, s.K / kQK, 0, st)
                                       : launch_quantize_q8_1(s.act_f32, tmp, s.K / kQK, 0, st);
            if (e != cudaSuccess) { cudaFreeAsync(tmp, st); return cuda_fail(e, "gemv_chain quantize"); }
            note_launch();
            act = tmp;
        }
        const uint32_t f = (flags & ~QGEMM_INPUTS_READY) | ((s.flags & QGEMM_INPUTS_READY) && s.act_q8_1 ? QGEMM_INPUTS_READY : 0u);
        int rc = qgemm_gemm_group(wtype, act, s.nmat, s.weights, s.C, s.F, 1, s.K, 1, s.ldc_f, f, stream);
        if (tmp) cudaFreeAsync(tmp, st);
        if (rc) return rc;
    }
    return QGEMM_OK;
}

size_t qgemm_prepack_bytes(int wtype, int F, int K) {
    if (!is_weight_type(wtype) || F <= 0 || K <= 0 || (K % kQK) != 0) return 0;
    return mmq_prepack_bytes(wtype, F, K);
}

int qgemm_prepack_weights(int wtype, const void* weight, int F, int K, void* packed, void* stream) {
    if (!is_weight_type(wtype) || F <= 0 || K <= 0 || (K % kQK) != 0 || !weight || !packed) return QGEMM_E_BADARG;
    if (!aligned(weight, 2) || !aligned(packed, 256)) return QGEMM_E_ALIGN;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    cudaError_t e = launch_mmq_prepack(wtype, weight, packed, F, K, (cudaStream_t)stream);
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "prepack launch");
}

int qgemm_hint_next_weights(const void* next_weights, size_t bytes) {
    t_pf_ptr = bytes ? next_weights : nullptr;
    t_pf_bytes = next_weights ? bytes : 0;
    return QGEMM_OK;
}

int qgemm_gemm(int wtype, const void* act_q8_1, const void* weight, float* C, int T, int F, int K, int64_t ldc_t,
               int64_t ldc_f, uint32_t flags, void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = check_gemm_args(wtype, act_q8_1, weight, C, T, F, K)) return rc;
    if (T == 0 || F == 0) return QGEMM_OK;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    return run_gemm(wtype, act_q8_1, weight, C, T, F, K, ldc_t, ldc_f, flags, workspace, workspace_bytes,
                    (cudaStream_t)stream, dev);
}

static int gemm_f32act_impl(int wtype, const float* act_f32, const float* gate, const void* weight, float* C, int T, int F, int K,
                           int64_t ldc_t, int64_t ldc_f, uint32_t flags, void* workspace, size_t workspace_bytes, void* stream);

int qgemm_gemm_f32act(int wtype, const float* act_f32, const void* weight, float* C, int T, int F, int K,
                      int64_t ldc_t, int64_t ldc_f, uint32_t flags, void* workspace, size_t workspace_bytes,
                      void* stream) {
    return gemm_f32act_impl(wtype, act_f32, nullptr, weight, C, T, F, K, ldc_t, ldc_f, flags, workspace, workspace_bytes, stream);
}

int qgemm_gemm_f32act_silu_mul(int wtype, const float* x, const float* gate, const void* weight, float* C, int T, int F, int K,
                               int64_t ldc_t, int64_t ldc_f, uint32_t flags, void* workspace, size_t workspace_bytes,
                               void* stream) {
    if (T > 0 && K > 0 && (!gate || !aligned(gate, 4))) return gate ? QGEMM_E_ALIGN : QGEMM_E_BADARG;
    return gemm_f32act_impl(wtype, x, gate, weight, C, T, F, K, ldc_t, ldc_f, flags, workspace, workspace_bytes, stream);
}

static int gemm_f32act_impl(int wtype, const float* act_f32, const float* gate, const void* weight, float* C, int T, int F, int K,
                           int64_t ldc_t, int64_t ldc_f, uint32_t flags, void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = check_gemm_args(wtype, act_f32, weight, C, T, F, K)) return rc;
    if (T == 0 || F == 0) return QGEMM_OK;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    const size_t a_q = align_up((size_t)T * (K / kQK) * kQ81Bytes, 256);
    if (K > 0 && (!workspace || workspace_bytes < a_q || !aligned(workspace, 16))) return QGEMM_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    {   // tensor-core sizes: the quantizer writes the GEMM's operand tiles directly (two launches in all)
        const uint32_t gflags = flags & 0xffffu, path = gflags & QGEMM_PATH_MASK;
        const bool tc = !(gflags & QGEMM_SEQUENTIAL) && (path == QGEMM_PATH_TCGEN05 || (path == QGEMM_PATH_AUTO && T >= mmq_min_tokens(wtype, weight, T, F, K)));
        if (tc && K > 0 && workspace_bytes > a_q) {
            const cudaError_t e = launch_mmq_f32act(wtype, act_f32, weight, C, T, F, K, ldc_t, ldc_f, gflags, (flags >> 16) & 0xffu,
                                                    (char*)workspace + a_q, workspace_bytes - a_q, dev.sms, st, gate);
            if (e == cudaSuccess) {
                t_last_path = QGEMM_PATH_TCGEN05;
                return QGEMM_OK;
            }
            if (e != cudaErrorNotSupported) return cuda_fail(e, "gemm_f32act launch");
        }
    }
    if (K > 0) {
        const cudaError_t e = gate ? launch_quantize_q8_1_silu_mul(act_f32, gate, workspace, (int64_t)T * (K / kQK), (flags >> 16) & 0xffu, st)
                                   : launch_quantize_q8_1(act_f32, workspace, (int64_t)T * (K / kQK), (flags >> 16) & 0xffu, st);
        if (e != cudaSuccess) return cuda_fail(e, "gemm_f32act quantize launch");
    }
    return run_gemm(wtype, workspace, weight, C, T, F, K, ldc_t, ldc_f, flags & 0xffffu, (char*)workspace + a_q,
                    workspace_bytes - a_q, st, dev);
}

// Explicit-argument forms of the L2 hint: nothing outlives the call (the one-shot qgemm_hint_next_weights() keeps the hint in
// thread-local state between two calls).
int qgemm_gemm_hinted(int wtype, const void* act_q8_1, const void* weight, float* C, int T, int F, int K, int64_t ldc_t,
                      int64_t ldc_f, uint32_t flags, void* workspace, size_t workspace_bytes, void* stream,
                      const void* next_weights, size_t next_bytes) {
    qgemm_hint_next_weights(next_weights, next_bytes);
    const int rc = qgemm_gemm(wtype, act_q8_1, weight, C, T, F, K, ldc_t, ldc_f, flags, workspace, workspace_bytes, stream);
    qgemm_hint_next_weights(nullptr, 0);   // a path that does not take hints (or an error) must not leave it for a later call
    return rc;
}

int qgemm_gemm_group_hinted(int wtype, const void* act_q8_1, int nmat, const void* const* weights, float* const* Cs, const int* Fs,
                            int T, int K, int64_t ldc_t, int64_t ldc_f, uint32_t flags, void* stream, const void* next_weights,
                            size_t next_bytes) {
    qgemm_hint_next_weights(next_weights, next_bytes);
    const int rc = qgemm_gemm_group(wtype, act_q8_1, nmat, weights, Cs, Fs, T, K, ldc_t, ldc_f, flags, stream);
    qgemm_hint_next_weights(nullptr, 0);
    return rc;
}

int qgemm_gemm_f16act(int wtype, const void* act_f16, const void* weight, float* C, int T, int F, int K, int64_t ldc_t,
                      int64_t ldc_f, uint32_t flags, void* workspace, size_t workspace_bytes, void* stream) {
    if (!is_weight_type(wtype) || T < 0 || F < 0 || K < 0 || (K % kQK) != 0) return QGEMM_E_BADARG;
    if (T == 0 || F == 0) return QGEMM_OK;
    if (!act_f16 || !weight || !C) return QGEMM_E_BADARG;
    if (!aligned(act_f16, 2) || !aligned(weight, 2) || !aligned(C, 4)) return QGEMM_E_ALIGN;   // halves: 2-byte alignment
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t a_q = align_up((size_t)T * (K / kQK) * kQ81Bytes, 256);
    void* pool = nullptr;
    if (K > 0 && !workspace && (flags & QGEMM_STREAM_ALLOC)) {   // launcher-style callers bring no scratch
        if (cudaError_t e = cudaMallocAsync(&pool, a_q, st)) return cuda_fail(e, "gemm_f16act scratch");
        workspace = pool;
        workspace_bytes = a_q;
    }
    if (K > 0 && (!workspace || workspace_bytes < a_q || !aligned(workspace, 16))) return QGEMM_E_WORKSPACE;
    int rc = QGEMM_OK;
    if (K > 0) {
        const cudaError_t e = launch_quantize_q8_1_f16(act_f16, workspace, (int64_t)T * (K / kQK), (flags >> 16) & 0xffu, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "gemm_f16act quantize launch");
    }
    if (rc == QGEMM_OK)
        rc = run_gemm(wtype, workspace, weight, C, T, F, K, ldc_t, ldc_f, flags & 0xffffu, workspace_bytes > a_q ? (char*)workspace + a_q : nullptr,
                      workspace_bytes > a_q ? workspace_bytes - a_q : 0, st, dev);
    if (pool) cudaFreeAsync(pool, st);
    return rc;
}

int qgemm_gemm_a16(int wtype, const float* act_f32, const void* weight, float* C, int T, int F, int K, int64_t ldc_t,
                   int64_t ldc_f, uint32_t flags, void* stream) {
    if ((wtype != QGEMM_TYPE_Q4_0 && wtype != QGEMM_TYPE_Q8_0) || T < 0 || F < 0 || K < 0 || (K % kQK) != 0) return QGEMM_E_BADARG;
    if (T == 0 || F == 0) return QGEMM_OK;
    if (!act_f32 || !weight || !C) return QGEMM_E_BADARG;
    if (!aligned(act_f32, 4) || !aligned(weight, 2) || !aligned(C, 4)) return QGEMM_E_ALIGN;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 0) {
        const int64_t n = (int64_t)T * F;
        fill_zero_strided<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(C, T, F, ldc_t, ldc_f);
        note_launch();
        return cudaGetLastError() == cudaSuccess ? QGEMM_OK : cuda_fail(cudaGetLastError(), "gemm_a16 fill");
    }
    cudaError_t e;
    if (!(flags & QGEMM_SEQUENTIAL) && f32act_supported(wtype, act_f32, weight, K)) {
        e = launch_gemm_f32act_dequant(wtype, act_f32, weight, C, T, F, K, ldc_t, ldc_f, dev.sms, st);
        t_last_path = T <= 8 ? QGEMM_PATH_GEMV : QGEMM_PATH_MMA;
    } else {
        e = launch_gemm_f32act_sequential(wtype, act_f32, weight, C, T, F, K, ldc_t, ldc_f, st);
        t_last_path = QGEMM_PATH_GENERIC;
    }
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "gemm_a16 launch");
}

int qgemm_sumi(int wtype, const void* act_q8_1, const void* weight, int32_t* sumi, int T, int F, int K, uint32_t flags,
               void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = check_gemm_args(wtype, act_q8_1, weight, sumi, T, F, K)) return rc;
    if (T == 0 || F == 0 || K == 0) return QGEMM_OK;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t path = flags & QGEMM_PATH_MASK;
    cudaError_t e;
    if (path == QGEMM_PATH_TCGEN05) {
        if (!mmq_supported(wtype, act_q8_1, weight, T, F, K)) return QGEMM_E_ALIGN;
        if (!workspace || workspace_bytes < mmq_workspace_need(wtype, weight, T, F, K, flags)) return QGEMM_E_WORKSPACE;
        e = launch_mmq(wtype, act_q8_1, weight, nullptr, sumi, T, F, K, 0, 0, flags, workspace, workspace_bytes,
                       dev.sms, st);
    } else {
        e = launch_sumi_generic(wtype, act_q8_1, weight, sumi, T, F, K, st);
    }
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "sumi launch");
}

static int make_peer_out(const qgemm_peers* peers, PeerOut* po) {
    if (!peers || peers->world < 1 || peers->world > kMaxPeers || peers->rank < 0 || peers->rank >= peers->world ||
        !peers->done || !peers->step || peers->launches_per_step == 0 || peers->launch_index >= peers->launches_per_step ||
        peers->wait_index > peers->launch_index)
        return QGEMM_E_BADARG;
    for (int r = 0; r < peers->world; r++)
        if (!peers->C[r] || !peers->flag[r]) return QGEMM_E_BADARG;
    *po = PeerOut{};
    po->world = peers->world; po->rank = peers->rank;
    for (int r = 0; r < peers->world; r++) { po->C[r] = peers->C[r]; po->flag[r] = peers->flag[r]; }
    po->mc = peers->world > 1 ? peers->C_multicast : nullptr;
    po->done = peers->done; po->step = peers->step; po->lps = peers->launches_per_step; po->li = peers->wait_index;
    po->dbg = QGEMM_ENV("QGEMM_PEER_DBG") ? atoi(QGEMM_ENV("QGEMM_PEER_DBG")) : 0;
    return QGEMM_OK;
}

int qgemm_gemm_group_peers(int wtype, const void* act_q8_1, int nmat, const void* const* weights, const int* Fs,
                           const int64_t* c_offsets, const qgemm_peers* peers, int T, int K, int64_t ldc_t, int64_t ldc_f,
                           uint32_t flags, void* stream) {
    if (nmat < 1 || nmat > 8 || !weights || !Fs || !c_offsets) return QGEMM_E_BADARG;
    PeerOut po;
    if (int rc = make_peer_out(peers, &po)) return rc;
    GemvGroup g{};
    g.nmat = nmat;
    int Ftot = 0;
    for (int m = 0; m < nmat; m++) {
        if (Fs[m] < 1) return QGEMM_E_BADARG;
        if (int rc = check_gemm_args(wtype, act_q8_1, weights[m], peers->C[peers->rank], T, Fs[m], K)) return rc;
        if (!gemv_supported(wtype, act_q8_1, weights[m], Fs[m], K)) return QGEMM_E_ALIGN;
        g.wgt[m] = weights[m]; g.C[m] = peers->C[peers->rank] + c_offsets[m]; g.F[m] = Fs[m];
        po.moff[m] = c_offsets[m];
        Ftot += Fs[m];
    }
    if (T < 1 || T > 8 || K < 32) return QGEMM_E_BADARG;
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    // peer launches share one CTA-completion counter and the step word: a launch that runs ahead of its predecessor
    // (INPUTS_READY) would mix its arrivals with the predecessor's, so the early-start promise is not honoured here
    flags &= ~QGEMM_INPUTS_READY;
    cudaError_t e = launch_gemv(wtype, act_q8_1, nullptr, nullptr, T, Ftot, K, ldc_t, ldc_f, flags, dev.sms,
                                (cudaStream_t)stream, &po, t_pf_ptr, t_pf_bytes, &g);
    t_pf_ptr = nullptr;
    t_pf_bytes = 0;
    t_last_path = QGEMM_PATH_GEMV;
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "gemm_group_peers launch");
}

int qgemm_gemm_peers(int wtype, const void* act_q8_1, const void* weight, const qgemm_peers* peers, int T, int F, int K,
                     int64_t ldc_t, int64_t ldc_f, uint32_t flags, void* stream) {
    PeerOut po;
    if (int rc = make_peer_out(peers, &po)) return rc;
    float* C = peers->C[peers->rank];
    if (int rc = check_gemm_args(wtype, act_q8_1, weight, C, T, F, K)) return rc;
    if (T == 0 || F == 0 || K == 0) return QGEMM_E_BADARG;  // every rank must launch: no empty shards in peer mode
    DeviceInfo dev;
    if (int rc = device_check(&dev)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    flags &= ~QGEMM_INPUTS_READY;  // see qgemm_gemm_group_peers
    if (T >= kMmaMinTokens && T <= 8 && gemv_mma_supported(wtype, act_q8_1, weight, T, F, K)) {
        e = launch_gemv_mma(wtype, act_q8_1, weight, C, T, F, K, ldc_t, ldc_f, flags, dev.sms, st, &po);
        t_last_path = QGEMM_PATH_MMA;
    } else if (T <= 8 && gemv_supported(wtype, act_q8_1, weight, F, K)) {
        e = launch_gemv(wtype, act_q8_1, weight, C, T, F, K, ldc_t, ldc_f, flags, dev.sms, st, &po, t_pf_ptr, t_pf_bytes);
        t_pf_ptr = nullptr;
        t_pf_bytes = 0;
        t_last_path = QGEMM_PATH_GEMV;
    } else if (T > 8 && mmq_supported(wtype, act_q8_1, weight, T, F, K)) {
        // prefill: peer stores from the tcgen05 epilogue.  Scratch = the registered default workspace, else
        // (QGEMM_STREAM_ALLOC) the stream's pool; the operand prepass reads the activations, so the wait for
        // earlier launches of the step runs as its own one-thread kernel in front of it.
        const size_t need = mmq_workspace_need(wtype, weight, T, F, K, flags);
        void* ws = nullptr;
        size_t ws_bytes = 0;
        int d = 0;
        if (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) {
            std::lock_guard<std::mutex> lk(g_dev_mu);
            ws = g_default_ws[d].ptr;
            ws_bytes = g_default_ws[d].bytes;
        }
        void* pool_ws = nullptr;
        if ((!ws || ws_bytes < need) && (flags & QGEMM_STREAM_ALLOC)) {
            // what the call needs, or what it can use (split-K scratch), whichever is larger
            const size_t want = align_up(std::max(need, mmq_workspace_bytes(wtype, T, F, K)), 256);
            if (scratch_alloc(&pool_ws, want, st) != cudaSuccess) {
                (void)cudaGetLastError();
                return QGEMM_E_WORKSPACE;
            }
            ws = pool_ws;
            ws_bytes = want;
        }
        if (!ws || ws_bytes < need) return QGEMM_E_WORKSPACE;
        if (po.world > 1 && po.li > 0) {
            peer_wait_kernel<<<1, 1, 0, st>>>(po.flag[po.rank], po.step, po.lps, po.li, (uint32_t)po.world);
            note_launch();
        }
        e = launch_mmq(wtype, act_q8_1, weight, C, nullptr, T, F, K, ldc_t, ldc_f, flags, ws, ws_bytes, dev.sms, st, &po);
        if (pool_ws) cudaFreeAsync(pool_ws, st);
        t_last_path = QGEMM_PATH_TCGEN05;
    } else {
        return QGEMM_E_ALIGN;  // peer stores exist in the decode kernels (T <= 8, bulk-copyable rows) and the tcgen05 path
    }
    return e == cudaSuccess ? QGEMM_OK : cuda_fail(e, "gemm_peers launch");
}

int qgemm_peer_step_advance(uint32_t* step, void* stream) {
    if (!step) return QGEMM_E_BADARG;
    peer_step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step);
    note_launch();
    return cudaGetLastError() == cudaSuccess ? QGEMM_OK : QGEMM_E_CUDA;
}

int qgemm_peer_wait(const qgemm_peers* peers, void* stream) {
    if (!peers || peers->world < 1 || peers->world > kMaxPeers || !peers->step) return QGEMM_E_BADARG;
    if (peers->world == 1) return QGEMM_OK;
    peer_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(peers->flag[peers->rank], peers->step, peers->launches_per_step,
                                                        peers->launches_per_step, (uint32_t)peers->world);
    note_launch();
    return cudaGetLastError() == cudaSuccess ? QGEMM_OK : QGEMM_E_CUDA;
}

int qgemm_shard_range(int F, int world, int rank, int align, int* f0, int* f1) {
    if (F < 0 || world < 1 || rank < 0 || rank >= world || align < 1 || !f0 || !f1) return QGEMM_E_BADARG;
    const int64_t units = ((int64_t)F + align - 1) / align;  // align-sized row groups
    const int64_t base = units / world, rem = units % world;
    const int64_t u0 = rank * base + (rank < rem ? rank : rem);
    const int64_t u1 = u0 + base + (rank < rem ? 1 : 0);
    int64_t a = u0 * align, b = u1 * align;
    if (a > F) a = F;
    if (b > F) b = F;
    *f0 = (int)a;
    *f1 = (int)b;
    return QGEMM_OK;
}

}  // extern "C"
