// InfoNCE over hypothesis scores (training variant, modules/model.py:43-63 / model_co3d.py:41-61), fused:
//   positives_b = { n : 180/pi * arccos((clamp(<R[b,n], G[b]>, -1, 3) - 1) / 2) <= ACC_THR }          (:46-49)
//   loss_b = -log( sum_{n in positives_b} exp(s[b,n]/T) / max(sum_n exp(s[b,n]/T), 1e-8) )            (:58-61)
// and, in the same launch, d loss_b / d s[b,n] (what autograd would push into the scores):
//   exp(s/T)/T * ( [sum > 1e-8]/sum_all - [n positive]/sum_pos ).
// The reference builds this from ~15 elementwise / reduction launches plus Python lists of index tensors per
// pair; here it is one CTA per pair, two passes over that pair's N scores and rotations.
#include "ahv_common.cuh"

namespace ahv {

constexpr int kNceThreads = 256;

__device__ __forceinline__ bool nce_positive(const float* __restrict__ r, const float* g, float thr_deg) {
  float dot = 0.0f;
#pragma unroll
  for (int e = 0; e < 9; ++e) dot = fmaf(__ldg(r + e), g[e], dot);
  const float sim = (fminf(fmaxf(dot, -1.0f), 3.0f) - 1.0f) * 0.5f;
  return 180.0f * (acosf(sim) * 0.318309886183790672f) <= thr_deg;  // 180 * arccos(sim) / pi, as the reference writes it
}

__global__ void __launch_bounds__(kNceThreads)
infonce_kernel(const float* __restrict__ scores, const float* __restrict__ R, int r_per_pair,
               const float* __restrict__ gt, float thr_deg, float inv_T, float* __restrict__ loss,
               float* __restrict__ grad, int64_t N) {
  __shared__ double red[2][kNceThreads / 32];
  __shared__ double tot[2];
  const int b = blockIdx.x, t = threadIdx.x;
  const float* s = scores + (size_t)b * N;
  const float* Rb = R + (r_per_pair ? (size_t)b * N * 9 : 0);
  float g[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) g[e] = __ldg(gt + b * 9 + e);
  double sp = 0.0, sa = 0.0;
  for (int64_t n = t; n < N; n += kNceThreads) {
    const float e = expf(s[n] * inv_T);
    sa += e;
    if (nce_positive(Rb + n * 9, g, thr_deg)) sp += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
  }
  if ((t & 31) == 0) { red[0][t >> 5] = sp; red[1][t >> 5] = sa; }
  __syncthreads();
  if (t == 0) {
    double p = 0.0, a = 0.0;
    for (int w = 0; w < kNceThreads / 32; ++w) { p += red[0][w]; a += red[1][w]; }  // fixed order: deterministic
    tot[0] = p; tot[1] = a;
    const float pf = (float)p, af = fmaxf((float)a, 1e-8f);     // .clamp(min=1e-8) (:61)
    loss[b] = -logf(pf / af);
  }
  if (!grad) return;
  __syncthreads();
  const float inv_p = (float)(1.0 / tot[0]);                      // no positive at all: inf, like autograd of the reference
  const float inv_a = (float)tot[1] > 1e-8f ? (float)(1.0 / tot[1]) : 0.0f;
  float* gb = grad + (size_t)b * N;
  for (int64_t n = t; n < N; n += kNceThreads) {
    const float e = expf(s[n] * inv_T) * inv_T;
    const bool pos = nce_positive(Rb + n * 9, g, thr_deg);
    gb[n] = e * (inv_a - (pos ? inv_p : 0.0f));
  }
}

int launch_infonce(const float* scores, const float* R, int r_per_pair, const float* gt, float thr_deg, float temperature,
                   float* loss, float* grad, int B, int64_t N, cudaStream_t s) {
  if (B == 0) return AHV_OK;
  infonce_kernel<<<B, kNceThreads, 0, s>>>(scores, R, r_per_pair, gt, thr_deg, 1.0f / temperature, loss, grad, N);
  AHV_CUDA_OK(cudaGetLastError());
  return AHV_OK;
}

}  // namespace ahv
