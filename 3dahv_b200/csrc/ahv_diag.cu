// Diagnostics: shared-memory read-bandwidth micro-benchmark.  MEASURED_PEAKS.json
// has no shared-memory figure, and the gather stage of the hot path is bounded
// by shared-memory bandwidth, so bench.py measures the
// denominator of that roofline with this kernel on the same device, same run.
#include "ahv_common.cuh"

namespace ahv {

__global__ void __launch_bounds__(1024, 2) smem_read_kernel(float* __restrict__ out, int iters) {
  extern __shared__ float4 sbuf[];  // 2048 float4 = 32 KB
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sbuf[i] = make_float4(i, 1.f, 2.f, 3.f);
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int idx = threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      // consecutive lanes -> consecutive 16 B: conflict-free LDS.128
      const float4 v = sbuf[(idx + u * 256) & 2047];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    idx = (idx + 96) & 2047;
  }
  if (acc.x + acc.y + acc.z + acc.w == -1.0f) out[blockIdx.x] = acc.x;
}

}  // namespace ahv

extern "C" AHV_API int ahv_diag_smem_read(float* out, int ctas, int iters, unsigned long long* bytes,
                                          void* stream) {
  if (ctas < 1 || iters < 1 || !out) return AHV_EINVAL;
  ahv::smem_read_kernel<<<ctas, 1024, 32768, (cudaStream_t)stream>>>(out, iters);
  if (cudaGetLastError() != cudaSuccess) return AHV_ECUDA;
  if (bytes) *bytes = (unsigned long long)ctas * 1024ull * (unsigned long long)iters * 8ull * 16ull;
  return AHV_OK;
}
