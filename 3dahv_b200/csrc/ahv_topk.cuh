// Warp-level building blocks of the selection kernels: a sorted list of 32 64-bit keys (make_key: score
// descending, ties -> lowest index) distributed one per lane, updated with shuffle-based bitonic sort / merge.
#pragma once
#include "ahv_common.cuh"

namespace ahv {

__device__ __forceinline__ u64 umax64(u64 a, u64 b) { return a > b ? a : b; }
__device__ __forceinline__ u64 umin64(u64 a, u64 b) { return a < b ? a : b; }

// full bitonic sort of one key per lane, descending (lane 0 = largest)
__device__ __forceinline__ u64 warp_sort_desc(u64 v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const u64 o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool desc = (lane & k) == 0;
      const bool lower = (lane & j) == 0;
      v = (lower == desc) ? umax64(v, o) : umin64(v, o);
    }
  }
  return v;
}
// sort a bitonic sequence descending
__device__ __forceinline__ u64 warp_bitonic_merge_desc(u64 v, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const u64 o = __shfl_xor_sync(0xffffffffu, v, j);
    v = ((lane & j) == 0) ? umax64(v, o) : umin64(v, o);
  }
  return v;
}
// merge a descending-sorted candidate list into the descending-sorted best list
__device__ __forceinline__ u64 warp_merge_sorted(u64 best, u64 cand_sorted, int lane) {
  const u64 rev = __shfl_sync(0xffffffffu, cand_sorted, 31 - lane);
  return warp_bitonic_merge_desc(umax64(best, rev), lane);
}
// offer one arbitrary candidate per lane
__device__ __forceinline__ u64 warp_offer(u64 best, u64 cand, int lane) {
  const u64 thr = __shfl_sync(0xffffffffu, best, 31);
  if (__any_sync(0xffffffffu, cand > thr)) best = warp_merge_sorted(best, warp_sort_desc(cand, lane), lane);
  return best;
}

}  // namespace ahv
