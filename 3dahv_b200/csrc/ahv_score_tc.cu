// tcgen05 tensor-core scoring path (placeholder until the kernel lands).
#include "ahv_common.cuh"
namespace ahv {
size_t score_tc_workspace_bytes(int, int64_t) { return 0; }
int launch_score_tc(const void*, int, const float*, const float*, int, const float*, const float*,
                    const float*, const float*, float*, int, int64_t, void*, size_t, cudaStream_t) {
  return AHV_ENOTSUP;
}
}  // namespace ahv
