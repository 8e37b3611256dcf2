// SS-form scoring kernel (AHV_TC_VARIANT=ss): both conv1 operand copies (YZ and X) and conv2's A operand in
// shared memory, every tcgen05.mma in the SS form; a pipeline stage is one hypothesis.  Overview, roles and
// precision notes: ahv_score_tc.cu.
#pragma once
#include "ahv_tc_common.cuh"

namespace ahv {
namespace tc {

// A operand of conv1, two fp16 copies of the rotated volume (8-row x 16-byte core matrices):
//   copy YZ: CM(d,h,chalf) = [w][c8] at chalf*yz_ch + d*yz_d + h*yz_h   (views y and z)
//   copy X : CM(d,w,chalf) = [h][c8] at chalf*x_ch  + d*1024 + w*128    (view x)
// Strides are chosen so that the gather's stores are bank-conflict-free for its lane map:
//   fp32 volumes (STS.64 per 4-channel chunk): dense, second channel-half block shifted by 64 B;
//   16-bit volumes (STS.128 per 8-channel chunk): YZ rows padded 128 -> 160 B.
template <bool K16>
struct Map {
  static constexpr int yz_h = K16 ? 160 : 128;
  static constexpr int yz_d = 8 * yz_h;
  static constexpr int yz_ch = K16 ? 8 * yz_d : 8 * yz_d + 64;
  static constexpr int yz_bytes = yz_ch + 8 * yz_d;
  static constexpr int x_ch = 8192 + 64;
  static constexpr int x_bytes = x_ch + 8192;
  static constexpr int stage_bytes = ((yz_bytes + x_bytes + 127) / 128) * 128;
  static constexpr int off_vol = 0;
  static constexpr int off_w1 = off_vol + kVolSmemBytes;          // 64000
  static constexpr int off_w2 = off_w1 + kW1Bytes;
  static constexpr int off_a = off_w2 + kW2Bytes;
  static constexpr int off_a2 = off_a + kStages * stage_bytes;
  static constexpr int off_bar = off_a2 + 2 * 8192;
  static constexpr int off_misc = off_bar + 16 * 8;               // tmem ptr, partial sums, base table
  static constexpr int smem_bytes = off_misc + 256;
  static_assert(off_w1 % 128 == 0 && off_a % 128 == 0 && off_a2 % 128 == 0 && off_bar % 8 == 0, "align");
  static_assert(smem_bytes <= 232448, "shared memory budget");
};

// 16-bit volumes are staged as "x-pair lines": for every (z, y, x0) of the halo'd grid one 64 B line
// holding BOTH x taps (x0, x0+1) x 16 channels as fp16 (exact for scaled bf16 inputs), chunk = tap*2 +
// chalf.  A tap pair is then four LDS.128, the gather reads half the bytes of the fp32 layout and
// interpolates with packed HFMA2 (fp16 accumulation adds <1e-4 relative to the scores; the bf16
// configuration's gate is 1e-2).
static_assert(kHalo * kHalo * 9 * 64 <= kVolSmemBytes, "pair lines (900 x 64 B) fit the volume region");

// ------------------------------------------------------------------------------
template <typename T, bool K16>
__global__ void __launch_bounds__(kThreadsTC, 1)
score_tc_kernel(const T* __restrict__ vol_src, const float* __restrict__ tgt_feat,
                const float* __restrict__ R, int r_per_pair, const float* __restrict__ b2,
                const float* __restrict__ base, const float* __restrict__ W1,
                const float* __restrict__ W2, float* __restrict__ scores,
                u64* __restrict__ best_keys, int B, int64_t N, Finalize fin) {
  extern __shared__ __align__(128) unsigned char smem[];
  using M = Map<K16>;
  constexpr int kOffVol = M::off_vol, kOffW1 = M::off_w1, kOffW2 = M::off_w2, kOffA = M::off_a, kOffA2 = M::off_a2,
                kOffBar = M::off_bar, kOffMisc = M::off_misc, kStageBytes = M::stage_bytes;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  Work work;
  {
    const int64_t total = (int64_t)B * N;
    work.lo = total * blockIdx.x / gridDim.x;
    work.hi = total * (blockIdx.x + 1) / gridDim.x;
    work.N = N;
  }
  if (work.lo >= work.hi) return;

  float* vol = reinterpret_cast<float*>(smem + kOffVol);
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + kOffBar;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffMisc);
  float* partial = reinterpret_cast<float*>(smem + kOffMisc + 16);  // [2 tilebuf][2 slot][4 warps]
  float* sbase = reinterpret_cast<float*>(smem + kOffMisc + 96);    // 8 base coordinates
  uint32_t* l1max_bits = reinterpret_cast<uint32_t*>(smem + kOffMisc + 128);  // max L1 norm of a W1 row (float bits)
  float* red = reinterpret_cast<float*>(smem + kOffMisc + 144);        // 8 per-warp maxima (volume staging)
  float* inv_ring = reinterpret_cast<float*>(smem + kOffMisc + 176);   // 1/scale of pair b at [b & 7] (gather -> epilogue)

  // ---- one-time setup ----
  for (int i = threadIdx.x; i < kVolSmemBytes / 16; i += kThreadsTC)
    reinterpret_cast<uint4*>(vol)[i] = make_uint4(0, 0, 0, 0);  // halo stays zero for the whole kernel
  if (threadIdx.x < 8) sbase[threadIdx.x] = base[threadIdx.x];
  if (threadIdx.x == 8) *l1max_bits = 0u;
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int i = 0; i < 3; ++i) { mbar_init(bar0 + (kFull + i) * 8, kGatherWarps); mbar_init(bar0 + (kEmpty + i) * 8, 1); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar0 + (kD1Full + i) * 8, 1);
        mbar_init(bar0 + (kD1Empty + i) * 8, 4);
        mbar_init(bar0 + (kA2Full + i) * 8, 4);
        mbar_init(bar0 + (kD2Full + i) * 8, 1);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 128);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp >= kGatherWarps) {  // MMA + epilogue warps: weights -> fp16 operand layouts, then release the MMA warp
    pack_weights(smem + kOffW1, W1, W2, threadIdx.x - kGatherWarps * 32);
    fence_proxy_async();  // written through the generic proxy, UMMA reads through the async proxy
    named_bar_sync(3, kPackThreads);
  }

  if (warp < kGatherWarps) {
    // =========================== GATHER ===========================
    // lane -> voxel (d = warp, h = 4e + hh, w): w = lane>>2, hh = lane&3
    const int w = lane >> 2, hh = lane & 3, d = warp;
    const int pf = w & 1;                 // bank parity this lane reads first
    const int rot = ((w & 3) + hh) & 3;   // chunk rotation: lanes of one LDS phase hit 8 distinct bank groups
    const float bx = sbase[w], bz = sbase[d];
    const float by0 = sbase[hh], by1 = sbase[4 + hh];
    uint32_t koff[4], syz[4], sx[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int ck = (rot + t) & 3;       // chunk visited at step t
      koff[t] = ck * 16;
      if constexpr (!K16) {               // fp32: chunk = channels 4ck..4ck+3, held in acc[t]
        syz[t] = (ck >> 1) * M::yz_ch + (ck & 1) * 8 + d * M::yz_d + hh * M::yz_h + w * 16;
        sx[t] = M::yz_bytes + (ck >> 1) * M::x_ch + (ck & 1) * 8 + d * 1024 + w * 128 + hh * 16;
      } else {                            // 16-bit: accumulator t&1 holds channel half (rot+t)&1; t<2 used
        const int ca = (rot + t) & 1;
        syz[t] = ca * M::yz_ch + d * M::yz_d + hh * M::yz_h + w * 16;
        sx[t] = M::yz_bytes + ca * M::x_ch + d * 1024 + w * 128 + hh * 16;
      }
    }
    const unsigned char* volb = smem + kOffVol;
    const int gtid = threadIdx.x;  // 0..255
    TileIter it(work);
    int cur_b = -1;
    uint32_t h = 0;  // hypothesis counter of this CTA (stage = h % 3)
    float Rn[9];
    {
      const int b0 = (int)(work.lo / N);
      const int64_t n0 = work.lo - (int64_t)b0 * N;
      const float* Rg = R + (r_per_pair ? ((size_t)b0 * N + n0) : (size_t)n0) * 9;
#pragma unroll
      for (int e = 0; e < 9; ++e) Rn[e] = __ldg(Rg + e) * opaque_one(r_per_pair);
    }
    while (it.advance()) {
      if (it.b != cur_b) {
        named_bar_sync(1, kGatherWarps * 32);  // everyone is done reading the previous volume
        const T* vg = vol_src + (size_t)it.b * kC * kVox;
        const float inv = stage_pair_volume<T, K16>(smem + kOffVol, vg, l1max_bits, W1, cur_b < 0, red, gtid);
        if (gtid == 0) inv_ring[it.b & 7] = inv;
        named_bar_sync(1, kGatherWarps * 32);
        cur_b = it.b;
      }
      for (int sl = 0; sl < 2; ++sl, ++h) {
        float Rr[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) Rr[e] = Rn[e];
        {  // prefetch the next hypothesis' rotation (hides the L2 round trip behind this gather)
          int nb; uint32_t nn;
          if (sl == 0) { nb = it.b; nn = it.n0 + (it.cnt > 1 ? 1 : 0); }
          else it.peek(nb, nn);
          const float* Rg = R + (r_per_pair ? ((size_t)nb * N + nn) : (size_t)nn) * 9;
#pragma unroll
          for (int e = 0; e < 9; ++e) Rn[e] = __ldg(Rg + e);
        }
        const uint32_t stage = h % kStages, use = h / kStages;
        if (use > 0) mbar_wait(bar0 + (kEmpty + stage) * 8, (use - 1) & 1);
        unsigned char* st = smem + kOffA + stage * kStageBytes;
#pragma unroll 1
        for (int e = 0; e < 2; ++e) {
          const float by = e ? by1 : by0;
          // grid = R @ (x, y, z)  (F.affine_grid, utils.py:126), then grid_sample's un-normalisation
          float ix = unnorm(fmaf(Rr[2], bz, fmaf(Rr[1], by, Rr[0] * bx)));
          float iy = unnorm(fmaf(Rr[5], bz, fmaf(Rr[4], by, Rr[3] * bx)));
          float iz = unnorm(fmaf(Rr[8], bz, fmaf(Rr[7], by, Rr[6] * bx)));
          ix = fminf(fmaxf(ix, -1.0f), 8.0f); iy = fminf(fmaxf(iy, -1.0f), 8.0f); iz = fminf(fmaxf(iz, -1.0f), 8.0f);
          const float x0 = fminf(floorf(ix), 7.0f), y0 = fminf(floorf(iy), 7.0f), z0 = fminf(floorf(iz), 7.0f);
          const float fx = ix - x0, fy = iy - y0, fz = iz - z0;
          if constexpr (!K16) {
            const int line = (((int)z0 + 1) * kHalo + ((int)y0 + 1)) * kHalo + ((int)x0 + 1);
            const int swap = (line ^ pf) & 1;  // first x tap = the one whose 64 B line has bank parity pf
            const float wxa = swap ? fx : 1.0f - fx, wxb = swap ? 1.0f - fx : fx;
            const unsigned char* pa = volb + (line + swap) * 64;
            const unsigned char* pb = volb + (line + 1 - swap) * 64;
            float wa[4], wb[4];
  #pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float wyz = ((c & 1) ? fy : 1.0f - fy) * ((c >> 1) ? fz : 1.0f - fz);
              wa[c] = wyz * wxa;
              wb[c] = wyz * wxb;
            }
            // 4 chunk batches of 8 LDS.128 each, software-pipelined: batch t+1 is in flight
            // while batch t is consumed (two register buffers)
            float4 buf[2][8];
  #pragma unroll
            for (int c = 0; c < 4; ++c) {  // c = dz*2 + dy
              const int off = ((c >> 1) * kHalo * kHalo + (c & 1) * kHalo) * 64;
              buf[0][2 * c] = *reinterpret_cast<const float4*>(pa + koff[0] + off);
              buf[0][2 * c + 1] = *reinterpret_cast<const float4*>(pb + koff[0] + off);
            }
  #pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (t < 3) {
  #pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const int off = ((c >> 1) * kHalo * kHalo + (c & 1) * kHalo) * 64;
                  buf[(t + 1) & 1][2 * c] = *reinterpret_cast<const float4*>(pa + koff[t + 1] + off);
                  buf[(t + 1) & 1][2 * c + 1] = *reinterpret_cast<const float4*>(pb + koff[t + 1] + off);
                }
              }
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  #pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float4 a = buf[t & 1][2 * c], g = buf[t & 1][2 * c + 1];
                acc.x = fmaf(wa[c], a.x, acc.x); acc.y = fmaf(wa[c], a.y, acc.y);
                acc.z = fmaf(wa[c], a.z, acc.z); acc.w = fmaf(wa[c], a.w, acc.w);
                acc.x = fmaf(wb[c], g.x, acc.x); acc.y = fmaf(wb[c], g.y, acc.y);
                acc.z = fmaf(wb[c], g.z, acc.z); acc.w = fmaf(wb[c], g.w, acc.w);
              }
              const __half2 lo = __floats2half2_rn(acc.x, acc.y), hi2 = __floats2half2_rn(acc.z, acc.w);
              uint2 pk;
              pk.x = *reinterpret_cast<const uint32_t*>(&lo);
              pk.y = *reinterpret_cast<const uint32_t*>(&hi2);
              *reinterpret_cast<uint2*>(st + syz[t] + e * (4 * M::yz_h)) = pk;  // h += 4
              *reinterpret_cast<uint2*>(st + sx[t] + e * 64) = pk;
            }
          } else {
            // ---- 16-bit staged volume: pair lines, both x taps per line ----
            const int pline = (((int)z0 + 1) * kHalo + ((int)y0 + 1)) * 9 + ((int)x0 + 1);
            const int swapy = (pline ^ pf) & 1;  // first y tap = the one whose line has bank parity pf (9 is odd)
            const unsigned char* pa = volb + (pline + swapy * 9) * 64;
            const unsigned char* pb = volb + (pline + (1 - swapy) * 9) * 64;
            const float wya = swapy ? fy : 1.0f - fy, wyb = swapy ? 1.0f - fy : fy;
            // tap weights as replicated half2: w[t][c] = wx(tap of chunk t) * wy(order c>>1) * wz(c&1)
            const __half2 wy2[2] = {__float2half2_rn(wya), __float2half2_rn(wyb)};
            const __half2 wz2[2] = {__float2half2_rn(1.0f - fz), __float2half2_rn(fz)};
            const __half2 wx2[2] = {__float2half2_rn(1.0f - fx), __float2half2_rn(fx)};
            __half2 w4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) w4[c] = __hmul2(wy2[c >> 1], wz2[c & 1]);
            // chunk visited at step t is (rot+t)&3 = tap*2+chalf, so its tap is ((rot+t)&3)>>1
            __half2 wxt[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) wxt[t] = (((rot + t) & 3) >> 1) ? wx2[1] : wx2[0];
            __half2 acc[2][4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) acc[0][e2] = acc[1][e2] = __float2half2_rn(0.0f);
            uint4 buf[2][4];
            constexpr int kDz = kHalo * 9 * 64;
            buf[0][0] = *reinterpret_cast<const uint4*>(pa + koff[0]);
            buf[0][1] = *reinterpret_cast<const uint4*>(pa + koff[0] + kDz);
            buf[0][2] = *reinterpret_cast<const uint4*>(pb + koff[0]);
            buf[0][3] = *reinterpret_cast<const uint4*>(pb + koff[0] + kDz);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (t < 3) {
                buf[(t + 1) & 1][0] = *reinterpret_cast<const uint4*>(pa + koff[t + 1]);
                buf[(t + 1) & 1][1] = *reinterpret_cast<const uint4*>(pa + koff[t + 1] + kDz);
                buf[(t + 1) & 1][2] = *reinterpret_cast<const uint4*>(pb + koff[t + 1]);
                buf[(t + 1) & 1][3] = *reinterpret_cast<const uint4*>(pb + koff[t + 1] + kDz);
              }
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const __half2 wg = __hmul2(wxt[t], w4[c]);
                const uint4 q4 = buf[t & 1][c];
                const uint32_t wd[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2)
                  acc[t & 1][k2] = __hfma2(wg, *reinterpret_cast<const __half2*>(&wd[k2]), acc[t & 1][k2]);
              }
            }
#pragma unroll
            for (int a2i = 0; a2i < 2; ++a2i) {
              const uint4 q4 = make_uint4(*reinterpret_cast<const uint32_t*>(&acc[a2i][0]),
                                          *reinterpret_cast<const uint32_t*>(&acc[a2i][1]),
                                          *reinterpret_cast<const uint32_t*>(&acc[a2i][2]),
                                          *reinterpret_cast<const uint32_t*>(&acc[a2i][3]));
              *reinterpret_cast<uint4*>(st + syz[a2i] + e * (4 * M::yz_h)) = q4;  // h += 4
              *reinterpret_cast<uint4*>(st + sx[a2i] + e * 64) = q4;
            }
          }
        }
        fence_proxy_async();  // make this thread's A-operand stores visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + (kFull + stage) * 8);
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA ISSUER (converged warp) ===========================
    {
      constexpr uint32_t idesc1 = instr_desc(64, 32), idesc2 = instr_desc(128, 32);
      const uint32_t w1s = s_base + kOffW1, w2s = s_base + kOffW2;
      TileIter it(work);
      uint32_t h = 0, g = 0;
      auto conv2 = [&](uint32_t gg) {
        const uint32_t gb = gg & 1, u = gg >> 1;
        mbar_wait(bar0 + (kA2Full + gb) * 8, u & 1);
        tc_fence_after();
        const uint32_t a2 = s_base + kOffA2 + gb * kA2Bytes;
#pragma unroll
        for (int i = 0; i < 2; ++i)
          umma_f16(tmem + 64 + gb * 32, smem_desc(a2 + i * 4096, 2048, 128), smem_desc(w2s + i * 1024, 512, 128), idesc2, i);
        umma_commit(bar0 + (kD2Full + gb) * 8);
      };
      while (it.advance()) {
        const uint32_t gb = g & 1, u = g >> 1;
        if (u > 0) mbar_wait(bar0 + (kD1Empty + gb) * 8, (u - 1) & 1);
        for (int sl = 0; sl < 2; ++sl, ++h) {
          const uint32_t stage = h % kStages, use = h / kStages;
          mbar_wait(bar0 + (kFull + stage) * 8, use & 1);
          tc_fence_after();
          const uint32_t a = s_base + kOffA + stage * kStageBytes;
          const uint32_t d1 = tmem + ((uint32_t)(16 * sl) << 16) + gb * 32;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // view x: rows (d,h), K slice = (w=kk, c)
            umma_f16(d1, smem_desc(a + M::yz_bytes + kk * 128, M::x_ch, 1024), smem_desc(w1s + kk * 1024, 512, 128), idesc1, kk);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // view y: rows (d,w), K slice = (h=kk, c)
            umma_f16(d1, smem_desc(a + kk * M::yz_h, M::yz_ch, M::yz_d), smem_desc(w1s + (8 + kk) * 1024, 512, 128), idesc1, 1);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // view z: rows (h,w), K slice = (d=kk, c)
            umma_f16(d1, smem_desc(a + kk * M::yz_d, M::yz_ch, M::yz_h), smem_desc(w1s + (16 + kk) * 1024, 512, 128), idesc1, 1);
          umma_commit(bar0 + (kEmpty + stage) * 8);  // A stage may be overwritten once these MMAs retire
        }
        umma_commit(bar0 + (kD1Full + gb) * 8);
        if (g > 0) conv2(g - 1);
        ++g;
      }
      conv2(g - 1);
    }
    __syncwarp();
  } else {
    // =========================== EPILOGUE ===========================
    epilogue_role<false>(work, warp - kEpiWarp0, lane, tmem, bar0, smem + kOffA2, 0, partial, tgt_feat, b2, inv_ring,
                         scores, best_keys, N, B, R, r_per_pair, fin);
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

}  // namespace tc
}  // namespace ahv
