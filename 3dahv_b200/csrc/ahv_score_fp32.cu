// fp32 CUDA-core kernels of the path (AHV_MATH_FP32 and the refcompat ops).
//
//  * score_fp32_kernel   — fused modules/model.py:186-193: per (pair,
//    hypothesis) trilinear resample (utils.py:113-131) -> tri-plane head
//    (modules/modules.py:112-124) -> cosine vs target -> mean.  Nothing is
//    materialised in HBM except the 4-byte score.
//  * forward_3d2d_kernel — modules/modules.py:112-124 on plain volumes
//    (target features, refcompat).
//  * rotate_volume_kernel — utils.py:113-131 materialised (refcompat only).
//
// Data layout in shared memory (one CTA = 256 threads, 1 CTA / SM):
//   vol   [1000 lines][16 ch] fp32, zero halo          64000 B
//   w1t   [384 k][32 o]      fp32 (W1 transposed)      49152 B
//   rotA  [16 c][8 d][72]    rotated volume, (h,w) rows padded 64->72
//   rotT  [16 c][8 d][72]    same with h<->w swapped (view x reads along h)
//   h1s   [64 pos][36]       ReLU(conv1) rows padded 32->36
//   w2s   [32 rows][36]      W2, row r(o) = (o%8)*4 + o/8
#include "ahv_common.cuh"

namespace ahv {

constexpr int kThreads = 256;
constexpr int kRotD = 72;          // padded (h,w) plane
constexpr int kRotC = kS * kRotD;  // 576 floats per channel
constexpr int kH1Row = 36;

struct Fp32Smem {
  float vol[kLines * kC];
  float w1t[kK * kO];
  float rotA[kC * kRotC];
  float rotT[kC * kRotC];
  float h1s[kP * kH1Row];
  float w2s[kO * kH1Row];
  float base[8];
  float red[8];
  float Rcur[12];
};

// ---- staging -------------------------------------------------------------
__device__ __forceinline__ void stage_weights(Fp32Smem& sm, const float* __restrict__ W1,
                                              const float* __restrict__ W2,
                                              const float* __restrict__ base) {
  for (int i = threadIdx.x; i < kO * kK; i += kThreads) {
    int o = i / kK, k = i % kK;  // coalesced read of W1[o][k]
    sm.w1t[k * kO + o] = W1[i];
  }
  for (int i = threadIdx.x; i < kO * kO; i += kThreads) {
    int o = i / kO, c = i % kO;
    sm.w2s[((o % 8) * 4 + o / 8) * kH1Row + c] = W2[i];
  }
  if (threadIdx.x < 8) sm.base[threadIdx.x] = base ? base[threadIdx.x] : 0.0f;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// vol_g: [16][512] NCDHW -> halo'd, channel-innermost lines
template <typename T>
__device__ __forceinline__ void stage_volume(float* __restrict__ vol, const T* __restrict__ vol_g) {
  for (int i = threadIdx.x; i < kLines * kC; i += kThreads) vol[i] = 0.0f;
  __syncthreads();
  for (int i = threadIdx.x; i < kC * kVox; i += kThreads) {
    int c = i >> 9, v = i & 511;
    int z = v >> 6, y = (v >> 3) & 7, x = v & 7;
    int line = ((z + 1) * kHalo + (y + 1)) * kHalo + (x + 1);
    vol[line * kC + c] = to_f32<T>(vol_g[i]);
  }
}

// ---- phase G: trilinear gather of one hypothesis into rotA / rotT ---------
// 4 lanes per output voxel (4 channels each), 8 voxels per warp step.
__device__ __forceinline__ void gather_hypothesis(Fp32Smem& sm, const float* R) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = lane & 3, q = lane >> 2;
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const int s = it * 8 + warp;  // (d,h) slab
    const int d = s >> 3, h = s & 7, w = q;
    Tap t = make_tap(R, sm.base[w], sm.base[h], sm.base[d]);
    // Two voxels share a 128-bit shared-memory phase (8 lanes).  Their x-taps
    // sit in adjacent 64 B lines of opposite bank halves; issuing them in
    // parity order makes every phase conflict-free.
    const int swap = (t.line ^ q) & 1;
    const float wx_first = swap ? t.fx : 1.0f - t.fx;
    const float wx_second = swap ? 1.0f - t.fx : t.fx;
    const float* p0 = sm.vol + (t.line + swap) * kC + j * 4;
    const float* p1 = sm.vol + (t.line + 1 - swap) * kC + j * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const float wyz = (dy ? t.fy : 1.0f - t.fy) * (dz ? t.fz : 1.0f - t.fz);
        const int off = (dz * kHalo * kHalo + dy * kHalo) * kC;
        const float4 a = *reinterpret_cast<const float4*>(p0 + off);
        const float4 b = *reinterpret_cast<const float4*>(p1 + off);
        const float wa = wyz * wx_first, wb = wyz * wx_second;
        acc.x = fmaf(wa, a.x, acc.x); acc.y = fmaf(wa, a.y, acc.y);
        acc.z = fmaf(wa, a.z, acc.z); acc.w = fmaf(wa, a.w, acc.w);
        acc.x = fmaf(wb, b.x, acc.x); acc.y = fmaf(wb, b.y, acc.y);
        acc.z = fmaf(wb, b.z, acc.z); acc.w = fmaf(wb, b.w, acc.w);
      }
    const int oa = d * kRotD + h * 8 + w, ot = d * kRotD + w * 8 + h;
    const int c0 = j * 4;
    sm.rotA[(c0 + 0) * kRotC + oa] = acc.x; sm.rotT[(c0 + 0) * kRotC + ot] = acc.x;
    sm.rotA[(c0 + 1) * kRotC + oa] = acc.y; sm.rotT[(c0 + 1) * kRotC + ot] = acc.y;
    sm.rotA[(c0 + 2) * kRotC + oa] = acc.z; sm.rotT[(c0 + 2) * kRotC + ot] = acc.z;
    sm.rotA[(c0 + 3) * kRotC + oa] = acc.w; sm.rotT[(c0 + 3) * kRotC + ot] = acc.w;
  }
}

// plain (un-rotated) volume [16][512] -> rotA / rotT (forward_3d2d on its own)
template <typename T>
__device__ __forceinline__ void load_plain_volume(Fp32Smem& sm, const T* __restrict__ vol_g) {
  for (int i = threadIdx.x; i < kC * kVox; i += kThreads) {
    int c = i >> 9, v = i & 511;
    int d = v >> 6, h = (v >> 3) & 7, w = v & 7;
    float x = to_f32<T>(vol_g[i]);
    sm.rotA[c * kRotC + d * kRotD + h * 8 + w] = x;
    sm.rotT[c * kRotC + d * kRotD + w * 8 + h] = x;
  }
}

// ---- phase C1: tri-plane conv 384->32 + ReLU -> h1s ------------------------
// h1[o,p,q] = sum_{c,k} W1[o,c*8+k] V[c,p,q,k] + W1[o,128+c*8+k] V[c,p,k,q]
//                     + W1[o,256+c*8+k] V[c,k,p,q]      (modules/modules.py:115-118)
// thread tile: 4 positions (p, q0..q0+3) x 2 channels.
__device__ __forceinline__ void conv1_relu(Fp32Smem& sm) {
  const int cg = threadIdx.x & 15, pg = threadIdx.x >> 4;
  const int p = pg >> 1, q0 = (pg & 1) * 4;
  float acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.0f;
  const float* wbase = sm.w1t + 2 * cg;
#pragma unroll 1
  for (int view = 0; view < 3; ++view) {
    const float* a0 = (view == 0 ? sm.rotT : sm.rotA) + (view == 2 ? p * 8 + q0 : p * kRotD + q0);
    const int kstride = (view == 2) ? kRotD : 8;
#pragma unroll 2
    for (int c = 0; c < kC; ++c) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(a0 + c * kRotC + kk * kstride);
        const float2 w = *reinterpret_cast<const float2*>(wbase + (view * 128 + c * 8 + kk) * kO);
        acc[0][0] = fmaf(a.x, w.x, acc[0][0]); acc[0][1] = fmaf(a.x, w.y, acc[0][1]);
        acc[1][0] = fmaf(a.y, w.x, acc[1][0]); acc[1][1] = fmaf(a.y, w.y, acc[1][1]);
        acc[2][0] = fmaf(a.z, w.x, acc[2][0]); acc[2][1] = fmaf(a.z, w.y, acc[2][1]);
        acc[3][0] = fmaf(a.w, w.x, acc[3][0]); acc[3][1] = fmaf(a.w, w.y, acc[3][1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 r = make_float2(fmaxf(acc[i][0], 0.0f), fmaxf(acc[i][1], 0.0f));
    *reinterpret_cast<float2*>(sm.h1s + (p * 8 + q0 + i) * kH1Row + 2 * cg) = r;
  }
}

// ---- phase C2: conv 32->32 + bias; thread = (position, 8 output channels) --
__device__ __forceinline__ void conv2_bias(const Fp32Smem& sm, const float* b2r, float* v) {
  const int pos = threadIdx.x >> 2, cg2 = threadIdx.x & 3;
  float h[kO];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 t = *reinterpret_cast<const float4*>(sm.h1s + pos * kH1Row + 4 * i);
    h[4 * i] = t.x; h[4 * i + 1] = t.y; h[4 * i + 2] = t.z; h[4 * i + 3] = t.w;
  }
#pragma unroll
  for (int oo = 0; oo < 8; ++oo) {
    const float* wr = sm.w2s + (oo * 4 + cg2) * kH1Row;
    float a = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 w = *reinterpret_cast<const float4*>(wr + 4 * i);
      a = fmaf(w.x, h[4 * i], a); a = fmaf(w.y, h[4 * i + 1], a);
      a = fmaf(w.z, h[4 * i + 2], a); a = fmaf(w.w, h[4 * i + 3], a);
    }
    v[oo] = a + b2r[oo];
  }
}

__device__ __forceinline__ float quad_sum(float x) {
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  return x;
}

// ---- fused scoring kernel ---------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
score_fp32_kernel(const T* __restrict__ vol_src, const float* __restrict__ tgt_feat,
                  const float* __restrict__ R, int r_per_pair, const float* __restrict__ W1,
                  const float* __restrict__ W2, const float* __restrict__ b2,
                  const float* __restrict__ base, float* __restrict__ scores, int B, int64_t N) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Fp32Smem& sm = *reinterpret_cast<Fp32Smem*>(smem_raw);
  const int64_t total = (int64_t)B * N;
  const int64_t lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;
  if (lo >= hi) return;
  stage_weights(sm, W1, W2, base);
  const int pos = threadIdx.x >> 2, cg2 = threadIdx.x & 3;
  float b2r[8], tg[8];
#pragma unroll
  for (int oo = 0; oo < 8; ++oo) b2r[oo] = b2[cg2 * 8 + oo];
  int cur_b = -1;
  for (int64_t it = lo; it < hi; ++it) {
    const int b = (int)(it / N);
    const int64_t n = it - (int64_t)b * N;
    if (b != cur_b) {
      __syncthreads();
      stage_volume<T>(sm.vol, vol_src + (size_t)b * kC * kVox);
#pragma unroll
      for (int oo = 0; oo < 8; ++oo) tg[oo] = tgt_feat[((size_t)b * kO + cg2 * 8 + oo) * kP + pos];
      cur_b = b;
    }
    if (threadIdx.x < 9)
      sm.Rcur[threadIdx.x] = R[(r_per_pair ? (size_t)it : (size_t)n) * 9 + threadIdx.x];
    __syncthreads();
    float Rr[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Rr[e] = sm.Rcur[e];
    gather_hypothesis(sm, Rr);
    __syncthreads();
    conv1_relu(sm);
    __syncthreads();
    float v[8];
    conv2_bias(sm, b2r, v);
    float ss = 0.0f, dt = 0.0f;
#pragma unroll
    for (int oo = 0; oo < 8; ++oo) { ss = fmaf(v[oo], v[oo], ss); dt = fmaf(v[oo], tg[oo], dt); }
    ss = quad_sum(ss);
    dt = quad_sum(dt);
    // F.normalize: v / max(||v||, 1e-12); then <.,tgt> summed over channels
    float cosv = (cg2 == 0) ? dt / fmaxf(sqrtf(ss), 1e-12f) : 0.0f;
    cosv = warp_sum(cosv);
    if ((threadIdx.x & 31) == 0) sm.red[threadIdx.x >> 5] = cosv;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += sm.red[w];
      scores[it] = tot * (1.0f / 64.0f);  // .mean(dim=-1) over 64 positions
    }
  }
}

// ---- forward_3d2d on plain volumes ----------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
forward_3d2d_kernel(const float* __restrict__ vol, const float* __restrict__ W1,
                    const float* __restrict__ W2, const float* __restrict__ b2,
                    float* __restrict__ feat, int64_t M) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Fp32Smem& sm = *reinterpret_cast<Fp32Smem*>(smem_raw);
  stage_weights(sm, W1, W2, nullptr);
  const int pos = threadIdx.x >> 2, cg2 = threadIdx.x & 3;
  float b2r[8];
#pragma unroll
  for (int oo = 0; oo < 8; ++oo) b2r[oo] = b2[cg2 * 8 + oo];
  for (int64_t m = blockIdx.x; m < M; m += gridDim.x) {
    __syncthreads();
    load_plain_volume<float>(sm, vol + (size_t)m * kC * kVox);
    __syncthreads();
    conv1_relu(sm);
    __syncthreads();
    float v[8];
    conv2_bias(sm, b2r, v);
    float ss = 0.0f;
#pragma unroll
    for (int oo = 0; oo < 8; ++oo) ss = fmaf(v[oo], v[oo], ss);
    ss = quad_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int oo = 0; oo < 8; ++oo)
      feat[((size_t)m * kO + cg2 * 8 + oo) * kP + pos] = v[oo] * inv;
  }
}

// ---- rotate_volume materialised (refcompat) --------------------------------
__global__ void __launch_bounds__(kThreads, 1)
rotate_volume_kernel(const float* __restrict__ vol, int per_rot, const float* __restrict__ R,
                     const float* __restrict__ base, float* __restrict__ out, int64_t n,
                     int rot_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* svol = reinterpret_cast<float*>(smem_raw);
  float* sbase = svol + kLines * kC;
  float* sR = sbase + 8;
  if (threadIdx.x < 8) sbase[threadIdx.x] = base[threadIdx.x];
  const int64_t first = (int64_t)blockIdx.x * rot_per_cta;
  for (int r = 0; r < rot_per_cta; ++r) {
    const int64_t i = first + r;
    if (i >= n) break;
    if (per_rot || r == 0) {
      __syncthreads();
      stage_volume<float>(svol, vol + (per_rot ? (size_t)i * kC * kVox : 0));
    }
    __syncthreads();
    if (threadIdx.x < 9) sR[threadIdx.x] = R[i * 9 + threadIdx.x];
    __syncthreads();
    float Rr[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Rr[e] = sR[e];
    // task = (channel quad j, voxel v); consecutive lanes -> consecutive voxels
    for (int task = threadIdx.x; task < 4 * kVox; task += kThreads) {
      const int v = task & 511, j = task >> 9;
      const int d = v >> 6, h = (v >> 3) & 7, w = v & 7;
      Tap t = make_tap(Rr, sbase[w], sbase[h], sbase[d]);
      const float* p = svol + t.line * kC + j * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const float wgt = (dx ? t.fx : 1.0f - t.fx) * (dy ? t.fy : 1.0f - t.fy) *
                              (dz ? t.fz : 1.0f - t.fz);
            const float4 a = *reinterpret_cast<const float4*>(
                p + (dz * kHalo * kHalo + dy * kHalo + dx) * kC);
            acc.x = fmaf(wgt, a.x, acc.x); acc.y = fmaf(wgt, a.y, acc.y);
            acc.z = fmaf(wgt, a.z, acc.z); acc.w = fmaf(wgt, a.w, acc.w);
          }
      float* o = out + ((size_t)i * kC + j * 4) * kVox + v;
      o[0] = acc.x; o[kVox] = acc.y; o[2 * kVox] = acc.z; o[3 * kVox] = acc.w;
    }
  }
}

// ---- launchers --------------------------------------------------------------
static int sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return sms;
}

int launch_score_fp32(const void* vol_src, int vol_dtype, const float* tgt_feat, const float* R,
                      int r_per_pair, const float* W1, const float* W2, const float* b2,
                      const float* base, float* scores, int B, int64_t N, cudaStream_t s) {
  const int64_t total = (int64_t)B * N;
  if (total == 0) return AHV_OK;
  const int sms = sm_count();
  if (sms <= 0) return AHV_ECUDA;
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  const size_t smem = sizeof(Fp32Smem);
  if (vol_dtype == AHV_VOL_F32) {
    AHV_CUDA_OK(cudaFuncSetAttribute(score_fp32_kernel<float>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    score_fp32_kernel<float><<<grid, kThreads, smem, s>>>((const float*)vol_src, tgt_feat, R,
                                                          r_per_pair, W1, W2, b2, base, scores, B, N);
  } else {
    AHV_CUDA_OK(cudaFuncSetAttribute(score_fp32_kernel<__nv_bfloat16>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    score_fp32_kernel<__nv_bfloat16><<<grid, kThreads, smem, s>>>(
        (const __nv_bfloat16*)vol_src, tgt_feat, R, r_per_pair, W1, W2, b2, base, scores, B, N);
  }
  AHV_CUDA_OK(cudaGetLastError());
  return AHV_OK;
}

int launch_forward_3d2d(const float* vol, const float* W1, const float* W2, const float* b2,
                        float* feat, int64_t m, cudaStream_t s) {
  if (m == 0) return AHV_OK;
  const int sms = sm_count();
  if (sms <= 0) return AHV_ECUDA;
  const unsigned grid = (unsigned)(m < sms ? m : sms);
  const size_t smem = sizeof(Fp32Smem);
  AHV_CUDA_OK(cudaFuncSetAttribute(forward_3d2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  forward_3d2d_kernel<<<grid, kThreads, smem, s>>>(vol, W1, W2, b2, feat, m);
  AHV_CUDA_OK(cudaGetLastError());
  return AHV_OK;
}

int launch_rotate_volume(const float* vol, int per_rot, const float* R, const float* base,
                         float* out, int64_t n, cudaStream_t s) {
  if (n == 0) return AHV_OK;
  const int rpc = per_rot ? 1 : 8;
  const int64_t grid = (n + rpc - 1) / rpc;
  const size_t smem = (kLines * kC + 8 + 12) * sizeof(float);
  AHV_CUDA_OK(cudaFuncSetAttribute(rotate_volume_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  rotate_volume_kernel<<<(unsigned)grid, kThreads, smem, s>>>(vol, per_rot, R, base, out, n, rpc);
  AHV_CUDA_OK(cudaGetLastError());
  return AHV_OK;
}

}  // namespace ahv
