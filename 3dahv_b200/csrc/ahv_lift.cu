// ResNetBlock_3D(32 -> 16) of the 2D->3D lifting (modules/modules.py:9-47, applied at :100-101): the last step
// before the hypothesis-and-verification path.  Reference (BN=False, stride 1):
//     out = conv2(relu(conv1(x))) + downsample(x)
// conv1 = Conv3d(32,16,3,pad 1,no bias), conv2 = Conv3d(16,16,3,pad 1,no bias), downsample = Conv3d(32,16,1,no bias),
// on [M,32,8,8,8] -> [M,16,8,8,8].  cuDNN runs this as five launches; at B=1 it sits on the per-pair latency path.
//
// One launch: a thread-block CLUSTER of 8 CTAs per volume, CTA r = depth slab d = r.  Each CTA computes conv1+ReLU
// for its own slab into shared memory, the cluster synchronises, conv2 reads the two neighbouring slabs straight
// out of the neighbours' shared memory (DSMEM) - nothing is recomputed and the intermediate never touches HBM.
// Clusters are persistent: a cluster stages the 84 KB of weights once and walks its share of the volumes.
// fp32 FFMA throughout (bit-tight against the reference's fp32 CPU path; tensor cores are not used here because
// this stage is 1 % of the scoring kernel's time and feeds it its input, whose parity is measured in fp32).
//
// Thread map (256 threads), both convs: lane = (icg 8) x (ocg 4), warp = h (8 rows).  A thread accumulates the
// 8 w positions of row h for 4 output channels over ITS input-channel group, reading an input row as 10
// consecutive floats (w-1..w+8) and the weights as float4 over the 4 output channels; the partial sums of the 8
// input-channel groups are reduced with shuffles.
#include <cooperative_groups.h>

#include "ahv_common.cuh"

namespace ahv {
namespace cg = cooperative_groups;

namespace lift {
constexpr int kCin = 32, kCout = 16, kRow = 12, kPlane = 10 * kRow;  // padded planes: [10 h][12 w] (w index 0 = w -1)
constexpr int kThreads = 256;
// shared memory (floats)
// A thread's input-channel group (icg = lane >> 2) selects a slice of the inputs and of the weights; the slices are
// padded by 16 floats so that consecutive groups start 64 B apart modulo 128: the two groups of a quarter-warp then
// hit opposite bank halves (LDS.128: 4 lanes x 16 B of weights per group, one broadcast address of inputs per group).
constexpr int kXsIcg = 4 * 3 * kPlane + 16;     // conv1: 4 input channels per group
constexpr int kW1Icg = 4 * 27 * kCout + 16;
constexpr int kW2Icg = 2 * 27 * kCout + 16;     // conv2: 2 input channels per group
constexpr int kXs = 8 * kXsIcg;                 // input slabs d-1, d, d+1 with zero halo
constexpr int kW1s = 8 * kW1Icg;                // conv1 weights [icg][ic%4][kd][kh][kw][oc]
constexpr int kW2s = 8 * kW2Icg;                // conv2 weights [icg][ic%2][kd][kh][kw][oc]
constexpr int kWds = kCin * kCout;              // downsample   [ic][oc]                               512
constexpr int kHs = kCout * kPlane;             // one conv1 output slab with zero halo (own: DSMEM-visible)  1920
constexpr int kSmemFloats = kXs + kW1s + kW2s + kWds + 3 * kHs;
constexpr int kSmemBytes = kSmemFloats * 4;     // 155 648 B

// acc[w][o] += sum over the thread's channels / taps.  `in` points at channel 0 of the thread's group, plane dz = 0.
template <int kChannels, int kChStride, int kPlaneStride>
__device__ __forceinline__ void conv_rows(float (&acc)[8][4], const float* in, const float* w, int h, int ocg) {
#pragma unroll 1
  for (int c = 0; c < kChannels; ++c) {
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const float* row = in + c * kChStride + kd * kPlaneStride + (h + kh) * kRow;  // rows h-1..h+1 at halo index h..h+2
        const float4 a = *reinterpret_cast<const float4*>(row), b = *reinterpret_cast<const float4*>(row + 4);
        const float2 e = *reinterpret_cast<const float2*>(row + 8);
        const float x[10] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, e.x, e.y};
        const float* wp = w + ((c * 3 + kd) * 3 + kh) * 3 * kCout + ocg * 4;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4 wv = *reinterpret_cast<const float4*>(wp + kw * kCout);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acc[i][0] = fmaf(x[i + kw], wv.x, acc[i][0]);
            acc[i][1] = fmaf(x[i + kw], wv.y, acc[i][1]);
            acc[i][2] = fmaf(x[i + kw], wv.z, acc[i][2]);
            acc[i][3] = fmaf(x[i + kw], wv.w, acc[i][3]);
          }
        }
      }
    }
  }
}

__device__ __forceinline__ void reduce_icg(float (&acc)[8][4]) {  // sum over the 8 lanes that share (ocg, h)
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float v = acc[i][o];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      acc[i][o] = v;
    }
}

__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(kThreads, 1)
resblock3d_kernel(const float* __restrict__ x, const float* __restrict__ Wc1, const float* __restrict__ Wc2,
                  const float* __restrict__ Wd, float* __restrict__ out, int M) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;
  float* w1s = xs + kXs;
  float* w2s = w1s + kW1s;
  float* wds = w2s + kW2s;
  float* hs_own = wds + kWds;     // conv1+ReLU output, slab d (read by the neighbours through DSMEM)
  float* hs_lo = hs_own + kHs;    // local copies of slabs d-1 and d+1
  float* hs_hi = hs_lo + kHs;
  cg::cluster_group cluster = cg::this_cluster();
  const int d = (int)cluster.block_rank();
  const int t = threadIdx.x, lane = t & 31, h = t >> 5, ocg = lane & 3, icg = lane >> 2;

  // ---- once per CTA: zero the padded buffers, stage the weights (loads batched 9 deep: one L2 round trip each) ----
  for (int i = t; i < kXs + 0; i += kThreads) xs[i] = 0.0f;
  for (int i = t; i < 3 * kHs; i += kThreads) hs_own[i] = 0.0f;
  {
    float v[9];
    for (int i0 = 0; i0 < kCin * 27 * kCout; i0 += 9 * kThreads) {  // global [oc][ic][27] -> shared [ic][27][oc]
#pragma unroll
      for (int j = 0; j < 9; ++j) v[j] = i0 + j * kThreads + t < kCin * 27 * kCout ? __ldg(Wc1 + i0 + j * kThreads + t) : 0.0f;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int i = i0 + j * kThreads + t;
        if (i < kCin * 27 * kCout) {
          const int ic = (i / 27) % kCin;
          w1s[(ic >> 2) * kW1Icg + ((ic & 3) * 27 + i % 27) * kCout + i / (27 * kCin)] = v[j];
        }
      }
    }
    for (int i0 = 0; i0 < kCout * 27 * kCout; i0 += 9 * kThreads) {
#pragma unroll
      for (int j = 0; j < 9; ++j) v[j] = i0 + j * kThreads + t < kCout * 27 * kCout ? __ldg(Wc2 + i0 + j * kThreads + t) : 0.0f;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int i = i0 + j * kThreads + t;
        if (i < kCout * 27 * kCout) {
          const int ic = (i / 27) % kCout;
          w2s[(ic >> 1) * kW2Icg + ((ic & 1) * 27 + i % 27) * kCout + i / (27 * kCout)] = v[j];
        }
      }
    }
    for (int i = t; i < kWds; i += kThreads) wds[(i % kCin) * kCout + i / kCin] = __ldg(Wd + i);
  }
  __syncthreads();

  // ---- persistent over volumes: cluster c takes volumes c, c + #clusters, ... (weights staged once) ----
  for (int m = blockIdx.x / 8; m < M; m += gridDim.x / 8) {
  const float* xm = x + (size_t)m * kCin * kVox;
  {  // input slabs d-1, d, d+1 (planes outside the volume stay zero): 24 loads per thread, all in flight
    float v[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) {  // [ic][dz][hw]: 64 consecutive floats per (ic, dz)
      const int i = j * kThreads + t, hw = i & 63, dz = (i >> 6) % 3, ic = i / 192, dd = d + dz - 1;
      v[j] = (dd >= 0 && dd < 8) ? __ldg(xm + ic * kVox + dd * 64 + hw) : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 24; ++j) {
      const int i = j * kThreads + t, hw = i & 63, dz = (i >> 6) % 3, ic = i / 192;
      xs[(ic >> 2) * kXsIcg + ((ic & 3) * 3 + dz) * kPlane + ((hw >> 3) + 1) * kRow + (hw & 7) + 1] = v[j];
    }
  }
  __syncthreads();

  // ---- conv1 + ReLU for slab d: 4 input channels per thread ----
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[i][o] = 0.0f;
  conv_rows<4, 3 * kPlane, kPlane>(acc, xs + icg * kXsIcg, w1s + icg * kW1Icg, h, ocg);
  reduce_icg(acc);
  if (icg == 0) {
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int i = 0; i < 8; ++i) hs_own[(ocg * 4 + o) * kPlane + (h + 1) * kRow + i + 1] = fmaxf(acc[i][o], 0.0f);
  }
  cluster.sync();  // every slab of this volume is in its CTA's shared memory

  // ---- fetch the neighbouring slabs through distributed shared memory ----
  if (d > 0) {
    const float* src = cluster.map_shared_rank(hs_own, d - 1);
    for (int i = t; i < kHs / 4; i += kThreads) reinterpret_cast<float4*>(hs_lo)[i] = reinterpret_cast<const float4*>(src)[i];
  }
  if (d < 7) {
    const float* src = cluster.map_shared_rank(hs_own, d + 1);
    for (int i = t; i < kHs / 4; i += kThreads) reinterpret_cast<float4*>(hs_hi)[i] = reinterpret_cast<const float4*>(src)[i];
  }
  cluster.sync();  // neighbours are done reading before anyone may exit; local copies visible to the CTA

  // ---- conv2 over (d-1, d, d+1) + downsample(x): 2 conv2 input channels and 4 residual channels per thread ----
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[i][o] = 0.0f;
  // slabs are three separate buffers [own, lo, hi]: run the 3x3x3 taps plane by plane
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    const int ic = icg * 2 + c;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const float* plane = (kd == 0 ? hs_lo : kd == 1 ? hs_own : hs_hi) + ic * kPlane;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const float* row = plane + (h + kh) * kRow;
        const float4 a = *reinterpret_cast<const float4*>(row), b = *reinterpret_cast<const float4*>(row + 4);
        const float2 e = *reinterpret_cast<const float2*>(row + 8);
        const float xv[10] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, e.x, e.y};
        const float* wp = w2s + icg * kW2Icg + ((c * 3 + kd) * 3 + kh) * 3 * kCout + ocg * 4;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4 wv = *reinterpret_cast<const float4*>(wp + kw * kCout);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acc[i][0] = fmaf(xv[i + kw], wv.x, acc[i][0]);
            acc[i][1] = fmaf(xv[i + kw], wv.y, acc[i][1]);
            acc[i][2] = fmaf(xv[i + kw], wv.z, acc[i][2]);
            acc[i][3] = fmaf(xv[i + kw], wv.w, acc[i][3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {  // downsample: 1x1x1 conv of the input slab d (plane dz = 1)
    const int ic = icg * 4 + c;
    const float* row = xs + icg * kXsIcg + (c * 3 + 1) * kPlane + (h + 1) * kRow + 1;
    const float4 wv = *reinterpret_cast<const float4*>(wds + ic * kCout + ocg * 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xv = row[i];
      acc[i][0] = fmaf(xv, wv.x, acc[i][0]);
      acc[i][1] = fmaf(xv, wv.y, acc[i][1]);
      acc[i][2] = fmaf(xv, wv.z, acc[i][2]);
      acc[i][3] = fmaf(xv, wv.w, acc[i][3]);
    }
  }
  reduce_icg(acc);
  if (icg == 0) {
    float* om = out + (size_t)m * kCout * kVox + d * 64 + h * 8;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float4* dst = reinterpret_cast<float4*>(om + (ocg * 4 + o) * kVox);
      dst[0] = make_float4(acc[0][o], acc[1][o], acc[2][o], acc[3][o]);
      dst[1] = make_float4(acc[4][o], acc[5][o], acc[6][o], acc[7][o]);
    }
  }
  __syncthreads();  // the input slabs are overwritten by the next volume
  }
}
}  // namespace lift

int launch_resblock3d(const float* x, const float* Wc1, const float* Wc2, const float* Wd, float* out, int64_t m,
                      cudaStream_t s) {
  if (m == 0) return AHV_OK;
  if (m > 0x7fffffffLL) return AHV_EINVAL;
  AHV_CUDA_OK(cudaFuncSetAttribute(lift::resblock3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lift::kSmemBytes));
  int dev = 0, sms = 0;
  AHV_CUDA_OK(cudaGetDevice(&dev));
  AHV_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t resident = sms / 8 > 0 ? sms / 8 : 1;  // one CTA per SM (154 KB): clusters that fit at once
  const int clusters = (int)(m < resident ? m : resident);
  lift::resblock3d_kernel<<<(unsigned)(clusters * 8), lift::kThreads, lift::kSmemBytes, s>>>(x, Wc1, Wc2, Wd, out, (int)m);
  AHV_CUDA_OK(cudaGetLastError());
  return AHV_OK;
}

}  // namespace ahv
