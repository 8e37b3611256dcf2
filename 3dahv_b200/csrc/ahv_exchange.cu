// Top-k exchange of a hypothesis set sharded over the GPUs of one NVSwitch node (the reference
// runs one GPU, batch 1): every rank holds the top-k of ITS shard per pair; this kernel writes them - score,
// global index and the rotation itself - into every peer's exchange buffer over NVLink, announces them with a
// sequence-numbered flag, waits for the peers' flags and merges (score descending, ties -> lowest global index),
// so that every rank ends with the bit-identical top-k over the WHOLE set and sampled_R[pred_index]
// (modules/model.py:195-196) without an NCCL call, a merge launch or a gather launch.  For k == 1 the scoring
// kernel does the same inside its own epilogue (ahv_tc_common.cuh::peer_exchange_and_merge); both share one
// buffer and one sequence counter (ahv_peer.cuh), so fused and top-k steps may alternate freely.
//
// One CTA: the payload is B*k*64 B per peer (256 KB at B=128, k=32) and the step is latency-bound.
#include "ahv_peer.cuh"
#include "ahv_topk.cuh"

namespace ahv {

constexpr int kXchgThreads = 512, kXchgWarps = kXchgThreads / 32;

__global__ void __launch_bounds__(kXchgThreads)
topk_exchange_kernel(const float* __restrict__ val, const int64_t* __restrict__ idx, const float* __restrict__ R,
                     int r_per_pair, int64_t idx_offset, int64_t N, int B, int k, float* __restrict__ out_val,
                     int64_t* __restrict__ out_idx, float* __restrict__ R_best, peer::Args pa) {
  __shared__ u64 skeys[kXchgWarps][peer::kMaxPeers * kMaxK];
  __shared__ int s_ok;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* mine = pa.bufs[pa.rank];
  uint32_t* hdr = reinterpret_cast<uint32_t*>(mine);
  const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(hdr) + 1u;  // every thread, before thread 0 updates it
  const int par = (int)(seq & 1u);
  if (tid == 0) s_ok = 1;
  // 1. this shard's lists -> every rank's buffer (NVLink peer stores; own buffer: local)
  for (int e = tid; e < B * k; e += kXchgThreads) {
    const int b = e / k, j = e - b * k;
    const int64_t gi = idx[e];
    const int64_t n = gi - idx_offset;
    float r9[9];
    const bool valid = gi >= 0 && n >= 0 && n < N;
    const float* src = R + ((r_per_pair ? (size_t)b * N : 0) + (size_t)(valid ? n : 0)) * 9;
#pragma unroll
    for (int q = 0; q < 9; ++q) r9[q] = valid ? __ldg(src + q) : 0.0f;
    const peer::Entry ent = peer::pack_entry(valid ? gi : -1, val[e], r9);
    for (int p = 0; p < pa.world; ++p)
      peer::store_entry(pa.bufs[p] + peer::entry_off(par, pa.rank, pa.cap_pairs, pa.cap_k, b, j), ent);
  }
  __threadfence_system();  // entries before the flag, system scope
  __syncthreads();
  // 2. announce, wait for the peers
  if (warp == 0) {
    peer::publish_flags(pa, seq, par, lane);
    if (!__all_sync(0xffffffffu, peer::wait_flags(pa, seq, par, lane)) && lane == 0) s_ok = 0;
  }
  __syncthreads();
  __threadfence_system();
  const bool ok = s_ok != 0;
  // 3. merge: one warp per pair
  const int total = pa.world * k;
  const float nanv = __int_as_float(0x7fc00000);
  for (int b = warp; b < B; b += kXchgWarps) {
    for (int e = lane; e < total; e += 32) {
      const int p = e / k, j = e - p * k;
      const uint4 q0 = __ldcv(reinterpret_cast<const uint4*>(mine + peer::entry_off(par, p, pa.cap_pairs, pa.cap_k, b, j)));
      const int64_t gi = peer::entry_index(q0);
      skeys[warp][e] = gi >= 0 ? make_key(__uint_as_float(q0.z), (uint32_t)gi) : 0ull;
    }
    __syncwarp();
    u64 best = 0;
    for (int i = 0; i < total; i += 32) best = warp_offer(best, i + lane < total ? skeys[warp][i + lane] : 0ull, lane);
    if (lane < k) {
      float v = -INFINITY;
      int64_t gi = -1;
      float o9[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) o9[q] = nanv;
      if (best != 0ull) {
        int e = 0;
        while (e < total - 1 && skeys[warp][e] != best) ++e;  // the entry this key came from (payload lookup)
        const uint4* src = reinterpret_cast<const uint4*>(mine + peer::entry_off(par, e / k, pa.cap_pairs, pa.cap_k, b, e % k));
        const uint4 q0 = __ldcv(src), q1 = __ldcv(src + 1), q2 = __ldcv(src + 2);
        v = __uint_as_float(q0.z);
        gi = peer::entry_index(q0);
        peer::unpack_rotation(q0, q1, q2, o9);
      }
      out_val[(size_t)b * k + lane] = ok ? v : nanv;
      out_idx[(size_t)b * k + lane] = ok ? gi : -1;
      if (R_best) {
#pragma unroll
        for (int q = 0; q < 9; ++q) R_best[((size_t)b * k + lane) * 9 + q] = ok ? o9[q] : nanv;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (tid == 0) {
    hdr[0] = seq;
    if (!ok) hdr[1] = 1u;
  }
}

int launch_topk_exchange(const float* val, const int64_t* idx, const float* R, int r_per_pair, int64_t idx_offset,
                         int64_t N, int B, int k, float* out_val, int64_t* out_idx, float* R_best,
                         const peer::Args& pa, cudaStream_t s) {
  if (pa.world < 2 || pa.world > peer::kMaxPeers || pa.rank < 0 || pa.rank >= pa.world) return AHV_EINVAL;
  if (B < 1 || k < 1 || k > kMaxK || B > pa.cap_pairs || k > pa.cap_k) return AHV_EINVAL;
  for (int r = 0; r < pa.world; ++r)
    if (!pa.bufs[r]) return AHV_EINVAL;
  topk_exchange_kernel<<<1, kXchgThreads, 0, s>>>(val, idx, R, r_per_pair, idx_offset, N, B, k, out_val, out_idx, R_best, pa);
  AHV_CUDA_OK(cudaGetLastError());
  return AHV_OK;
}

}  // namespace ahv
