// mmq.cu -- prefill path placeholder: filled in by the tcgen05 kernel.
#include "qgemm_common.cuh"
namespace qgemm {
bool mmq_supported(int, const void*, const void*, int, int, int) { return false; }
size_t mmq_workspace_bytes(int, int, int, int) { return 0; }
cudaError_t launch_mmq(int, const void*, const void*, float*, int32_t*, int, int, int, int64_t, int64_t, uint32_t,
                       void*, size_t, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace qgemm
