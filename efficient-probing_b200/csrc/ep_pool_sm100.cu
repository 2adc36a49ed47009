// tcgen05 / TMA pooling kernels -- placeholder until the fused kernels land: reports "unsupported"
// so that dispatch uses the general kernels.
#include "ep_sm100.cuh"

namespace ep {
bool sm100_supported(int, int, int, int, int) { return false; }
size_t sm100_workspace_bytes(int, int, int, int) { return 0; }
int sm100_pool_fwd(const void*, const float*, float, int, int, int, int, float*, float*, float*, float*, void*,
                   cudaStream_t) { return EP_ERR_UNSUPPORTED; }
int sm100_pool_bwd(const void*, const float*, float, int, int, int, int, const float*, const float*, const float*,
                   const float*, float*, void*, cudaStream_t) { return EP_ERR_UNSUPPORTED; }
}  // namespace ep
