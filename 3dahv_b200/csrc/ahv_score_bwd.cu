// Fused backward of the verification scores (training variant).
//
// The reference trains through `infoNCE_loss` (modules/model.py:43-63) with PyTorch autograd over
// rotate_volume -> forward_3d2d -> similarity, saving ~420 KB of activations per hypothesis.  Here one
// kernel recomputes the forward per (pair, hypothesis) in shared memory and pushes the upstream gradient
// g = dL/dscore[b,n] back to everything the score depends on, materialising nothing:
//
//   forward (as score_fp32_kernel)   X = rotate(V_b, R_n);  A = tri-plane(X);  H1 = relu(A W1^T);
//                                    H2 = H1 W2^T + b2;  F = H2 / max(|H2|, eps);  s = mean_p <F_p, T_p>
//   backward                         dH2 = (g/64) (T - F <F,T>) / |H2|          (per position p)
//                                    dT_b += (g/64) F;   db2 += sum_p dH2;   dW2 += dH2^T H1
//                                    dH1 = (dH2 W2) * [H1 > 0];   dW1 += dH1^T A;   dA = dH1 W1
//                                    dX = fold(dA)  (three views onto the rotated volume)
//                                    dV_b += rotate^T(dX, R_n)   (adjoint of the trilinear resampling)
//
// fp32 FFMA throughout (gradients match fp64 autograd of the oracle to ~1e-6 of their maximum).
// One CTA = 256 threads, 1 CTA / SM, contiguous range of (pair, hypothesis) items.  Persistent
// per-thread accumulators: dW1 (48), dW2 (4), db2 (8), dT (8, flushed per pair), dV (32, flushed per
// pair); they leave the CTA through global atomics (caller zero-initialises the outputs).
//
// The adjoint of the resampling is a GATHER, not a scatter (shared-memory atomics would cost ~130k cycles
// per hypothesis): every thread owns two input voxels, pulls them back through R^T, walks the 4 x 4 (d, h)
// columns around that point, intersects on each column the three |i_a - v_a| < 1 intervals in w, and for
// the few candidates left re-uses the taps the forward gather recorded, so the weights are the forward's
// bit for bit.  A matrix that is not a rotation falls back to scanning all 512 output voxels.
#include "ahv_head_fp32.cuh"

namespace ahv {

struct BwdSmem {
  Fp32Smem f;
  float dh2[kP * kH1Row];  // dL/dH2 [pos][36]
  float4 taps[kVox];       // (corner line, fx, fy, fz) of every output voxel of the current hypothesis
};
static_assert(sizeof(BwdSmem) <= 232448, "shared memory budget");

// contribution of output voxel `vo` (tap t) to input voxel with halo line `lin`: trilinear weight or 0
__device__ __forceinline__ float tap_weight(const float4 t, int lin) {
  const int delta = lin - __float_as_int(t.x);  // = 100 dz + 10 dy + dx with dx,dy,dz in {0,1} iff the voxel is a tap
  // |dx|,|dy| <= 8 < 10, so the decomposition is unique: test the 8 admissible values
  const int dz = delta >= 100, r1 = delta - 100 * dz;
  const int dy = r1 >= 10, dx = r1 - 10 * dy;
  if ((unsigned)dx > 1u || delta < 0 || delta > 111) return 0.0f;
  return (dx ? t.y : 1.0f - t.y) * (dy ? t.z : 1.0f - t.z) * (dz ? t.w : 1.0f - t.w);
}

__global__ void __launch_bounds__(kThreads, 1)
score_bwd_kernel(const float* __restrict__ vol_src, const float* __restrict__ tgt_feat,
                 const float* __restrict__ R, int r_per_pair, const float* __restrict__ W1,
                 const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ base,
                 const float* __restrict__ grad_scores, float* __restrict__ g_vol, float* __restrict__ g_tgt,
                 float* __restrict__ g_W1, float* __restrict__ g_W2, float* __restrict__ g_b2, int B, int64_t N) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& bs = *reinterpret_cast<BwdSmem*>(smem_raw);
  Fp32Smem& sm = bs.f;
  const int64_t total = (int64_t)B * N;
  const int64_t lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;
  if (lo >= hi) return;
  const int t = threadIdx.x;
  stage_weights(sm, W1, W2, base);
  // conv2 / normalise mapping: thread = (position, 8 output channels cg2*8 ..)
  const int pos = t >> 2, cg2 = t & 3;
  float b2r[8], tg[8];
#pragma unroll
  for (int oo = 0; oo < 8; ++oo) b2r[oo] = b2[cg2 * 8 + oo];
  // persistent accumulators
  float aW1[3][8][2];  // dW1[o = 2cg (+1)][k = view*128 + c*8 + kk], cg = t & 15, c = t >> 4
  float aW2[4];        // dW2[o = t >> 3][i = (t & 7) * 4 ..]
  float ab2[8];        // db2[cg2*8 + oo] (this thread's position only)
  float aT[8];         // dT_b[cg2*8 + oo][pos]
  float aV[2][kC];     // dV_b[c][voxel t + 256 j]
#pragma unroll
  for (int v = 0; v < 3; ++v)
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) aW1[v][kk][0] = aW1[v][kk][1] = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) aW2[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) ab2[i] = aT[i] = 0.0f;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int c = 0; c < kC; ++c) aV[j][c] = 0.0f;

  auto flush_pair = [&](int b) {  // dT and dV of pair b leave the CTA
#pragma unroll
    for (int oo = 0; oo < 8; ++oo) {
      if (aT[oo] != 0.0f) atomicAdd(g_tgt + ((size_t)b * kO + cg2 * 8 + oo) * kP + pos, aT[oo]);
      aT[oo] = 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        if (aV[j][c] != 0.0f) atomicAdd(g_vol + ((size_t)b * kC + c) * kVox + t + 256 * j, aV[j][c]);
        aV[j][c] = 0.0f;
      }
  };

  int cur_b = -1;
  for (int64_t it = lo; it < hi; ++it) {
    const int b = (int)(it / N);
    const int64_t n = it - (int64_t)b * N;
    if (b != cur_b) {
      if (cur_b >= 0) flush_pair(cur_b);
      __syncthreads();
      stage_volume<float>(sm.vol, vol_src + (size_t)b * kC * kVox);
#pragma unroll
      for (int oo = 0; oo < 8; ++oo) tg[oo] = tgt_feat[((size_t)b * kO + cg2 * 8 + oo) * kP + pos];
      cur_b = b;
    }
    if (t < 9) sm.Rcur[t] = R[(r_per_pair ? (size_t)it : (size_t)n) * 9 + t];
    __syncthreads();
    float Rr[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Rr[e] = sm.Rcur[e];
    const float g64 = grad_scores[it] * (1.0f / 64.0f);  // d mean over the 64 positions

    // ---------------- forward recompute ----------------
    gather_hypothesis<true>(sm, Rr, bs.taps);
    __syncthreads();
    conv1_relu(sm);
    __syncthreads();
    float v[8];
    conv2_bias(sm, b2r, v);
    float ss = 0.0f, ft = 0.0f;
#pragma unroll
    for (int oo = 0; oo < 8; ++oo) { ss = fmaf(v[oo], v[oo], ss); ft = fmaf(v[oo], tg[oo], ft); }
    ss = quad_sum(ss);
    ft = quad_sum(ft);
    {
      // F = v / max(|v|, eps) (modules/modules.py:122).  d<F,T>/dv = (T - F <F,T>) / |v| above the clamp, T / eps below
      const float nraw = sqrtf(ss), nr = fmaxf(nraw, 1e-12f), inv = 1.0f / nr;
      const float fdot = ft * inv;  // <F, T>
      const bool clamped = nraw < 1e-12f;
#pragma unroll
      for (int oo = 0; oo < 8; ++oo) {
        const float F = v[oo] * inv;
        const float d2 = g64 * inv * (clamped ? tg[oo] : tg[oo] - F * fdot);
        aT[oo] = fmaf(g64, F, aT[oo]);
        ab2[oo] += d2;
        v[oo] = d2;
      }
      float* dst = bs.dh2 + pos * kH1Row + cg2 * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();

    // ---------------- dW2 += dH2^T H1 ; dH1 = (dH2 W2) * [H1 > 0] ----------------
    {
      const int o = t >> 3, i0 = (t & 7) * 4;
#pragma unroll 8
      for (int p = 0; p < kP; ++p) {
        const float d = bs.dh2[p * kH1Row + o];
        const float4 h = *reinterpret_cast<const float4*>(sm.h1s + p * kH1Row + i0);
        aW2[0] = fmaf(d, h.x, aW2[0]); aW2[1] = fmaf(d, h.y, aW2[1]);
        aW2[2] = fmaf(d, h.z, aW2[2]); aW2[3] = fmaf(d, h.w, aW2[3]);
      }
    }
    float dh1[8];
    {
      const int ig = cg2;  // input channels ig*8 .. +7 of position `pos`
#pragma unroll
      for (int j = 0; j < 8; ++j) dh1[j] = 0.0f;
#pragma unroll 4
      for (int o = 0; o < kO; ++o) {
        const float d = bs.dh2[pos * kH1Row + o];
        const float* wr = sm.w2s + ((o % 8) * 4 + o / 8) * kH1Row + ig * 8;  // W2[o][ig*8 ..]
        const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
        dh1[0] = fmaf(d, w0.x, dh1[0]); dh1[1] = fmaf(d, w0.y, dh1[1]);
        dh1[2] = fmaf(d, w0.z, dh1[2]); dh1[3] = fmaf(d, w0.w, dh1[3]);
        dh1[4] = fmaf(d, w1.x, dh1[4]); dh1[5] = fmaf(d, w1.y, dh1[5]);
        dh1[6] = fmaf(d, w1.z, dh1[6]); dh1[7] = fmaf(d, w1.w, dh1[7]);
      }
      const float* hp = sm.h1s + pos * kH1Row + ig * 8;  // ReLU mask (modules/modules.py:68)
#pragma unroll
      for (int j = 0; j < 8; ++j) dh1[j] = hp[j] > 0.0f ? dh1[j] : 0.0f;
    }
    __syncthreads();  // every dW2 read of h1s is done: overwrite it in place with dH1
    {
      float* hp = sm.h1s + pos * kH1Row + cg2 * 8;
      *reinterpret_cast<float4*>(hp) = make_float4(dh1[0], dh1[1], dh1[2], dh1[3]);
      *reinterpret_cast<float4*>(hp + 4) = make_float4(dh1[4], dh1[5], dh1[6], dh1[7]);
    }
    __syncthreads();

    // ---------------- dW1 += dH1^T A  (A = tri-plane views of the rotated volume) ----------------
    {
      const int cg = t & 15, c = t >> 4;
#pragma unroll 1
      for (int pq = 0; pq < 16; ++pq) {
        const int p = pq >> 1, q0 = (pq & 1) * 4;
        float2 d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = *reinterpret_cast<const float2*>(sm.h1s + (p * 8 + q0 + j) * kH1Row + 2 * cg);
#pragma unroll
        for (int view = 0; view < 3; ++view) {
          const float* a0 = (view == 0 ? sm.rotT : sm.rotA) + c * kRotC + (view == 2 ? p * 8 + q0 : p * kRotD + q0);
          const int kstride = (view == 2) ? kRotD : 8;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(a0 + kk * kstride);
            float s0 = aW1[view][kk][0], s1 = aW1[view][kk][1];
            s0 = fmaf(a.x, d[0].x, s0); s1 = fmaf(a.x, d[0].y, s1);
            s0 = fmaf(a.y, d[1].x, s0); s1 = fmaf(a.y, d[1].y, s1);
            s0 = fmaf(a.z, d[2].x, s0); s1 = fmaf(a.z, d[2].y, s1);
            s0 = fmaf(a.w, d[3].x, s0); s1 = fmaf(a.w, d[3].y, s1);
            aW1[view][kk][0] = s0; aW1[view][kk][1] = s1;
          }
        }
      }
    }
    __syncthreads();  // rotA / rotT are free: rotA becomes dX[c][d][h][w]

    // ---------------- dA = dH1 W1, folded view by view into dX ----------------
    {
      const int pl = t & 15, c = t >> 4;  // positions pl + 16 j, channel c, all (view, kk)
      float* dX = sm.rotA + c * kRotC;
#pragma unroll 1
      for (int view = 0; view < 3; ++view) {
        float acc[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) acc[j][kk] = 0.0f;
        const float* wv = sm.w1t + (view * 128 + c * 8) * kW1Pitch;
#pragma unroll 2
        for (int iq = 0; iq < 8; ++iq) {
          float4 d[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) d[j] = *reinterpret_cast<const float4*>(sm.h1s + (pl + 16 * j) * kH1Row + iq * 4);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const float2 wa = *reinterpret_cast<const float2*>(wv + kk * kW1Pitch + iq * 4);
            const float2 wb = *reinterpret_cast<const float2*>(wv + kk * kW1Pitch + iq * 4 + 2);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              acc[j][kk] = fmaf(d[j].x, wa.x, fmaf(d[j].y, wa.y, fmaf(d[j].z, wb.x, fmaf(d[j].w, wb.y, acc[j][kk]))));
          }
        }
        // fold: every dX element gets exactly one contribution per view, so no races inside a view
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ps = pl + 16 * j, p = ps >> 3, q = ps & 7;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (view == 0) dX[p * kRotD + q * 8 + kk] = acc[j][kk];          // V'[c, d=p, h=q, w=kk]
            else if (view == 1) dX[p * kRotD + kk * 8 + q] += acc[j][kk];    // V'[c, d=p, h=kk, w=q]
            else dX[kk * kRotD + p * 8 + q] += acc[j][kk];                   // V'[c, d=kk, h=p, w=q]
          }
        }
        __syncthreads();
      }
    }

    // ---------------- dV_b += rotate^T(dX): adjoint of utils.py:113-131 as a gather ----------------
    {
      // is R a rotation?  then the output voxels that can touch an input voxel sit within sqrt(3) of R^T x
      float dev = 0.0f;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float gij = Rr[i] * Rr[j] + Rr[3 + i] * Rr[3 + j] + Rr[6 + i] * Rr[6 + j];  // (R^T R)_ij
          dev = fmaxf(dev, fabsf(gij - (i == j ? 1.0f : 0.0f)));
        }
      const bool is_rot = dev < 1e-3f;  // NaN compares false -> full scan
      // per axis a: the w-coefficient of the sample point and its reciprocal (one division per axis and item)
      float inv_r[3];
      bool flat[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        flat[a] = !(fabsf(Rr[3 * a]) > 1e-6f);
        inv_r[a] = flat[a] ? 0.0f : 1.0f / Rr[3 * a];
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int vi = t + 256 * j;
        const int z = vi >> 6, y = (vi >> 3) & 7, x = vi & 7;
        const int lin = ((z + 1) * kHalo + (y + 1)) * kHalo + (x + 1);
        auto visit = [&](int vo) {
          const float w = tap_weight(bs.taps[vo], lin);
          if (w != 0.0f) {
            const float* src = sm.rotA + (vo >> 6) * kRotD + (vo & 63);
#pragma unroll
            for (int c = 0; c < kC; ++c) aV[j][c] = fmaf(w, src[c * kRotC], aV[j][c]);
          }
        };
        if (is_rot) {
          // input voxel centre in normalised coordinates, pulled back: out = R^T in (grid = R out, utils.py:126)
          const float gx = (2 * x + 1) * 0.125f - 1.0f, gy = (2 * y + 1) * 0.125f - 1.0f, gz = (2 * z + 1) * 0.125f - 1.0f;
          const float ox = unnorm(Rr[0] * gx + Rr[3] * gy + Rr[6] * gz);
          const float oy = unnorm(Rr[1] * gx + Rr[4] * gy + Rr[7] * gz);
          const float oz = unnorm(Rr[2] * gx + Rr[5] * gy + Rr[8] * gz);
          (void)ox;
          const int h0 = (int)floorf(oy) - 1, d0 = (int)floorf(oz) - 1;
          // In index space the sample point of output voxel o' is i = R (o' - 3.5) + 3.5, and o' touches this
          // voxel iff |i_a - v_a| < 1 on all three axes.  Along a (d, h) column that is three intervals in w:
          // intersect them (with a margin; tap_weight() makes the exact decision) instead of testing 4 values.
          const float vx = (float)x, vy = (float)y, vz = (float)z;
#pragma unroll 1
          for (int dd = 0; dd < 4; ++dd) {
            const int d = d0 + dd;
            if ((unsigned)d > 7u) continue;
#pragma unroll 1
            for (int hh = 0; hh < 4; ++hh) {
              const int h = h0 + hh;
              if ((unsigned)h > 7u) continue;
              const float uh = (float)h - 3.5f, ud = (float)d - 3.5f;
              float ulo = -3.5f, uhi = 3.5f;  // u = w - 3.5
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                const float ca = fmaf(Rr[3 * a + 1], uh, fmaf(Rr[3 * a + 2], ud, 3.5f - (a == 0 ? vx : (a == 1 ? vy : vz))));
                if (!flat[a]) {
                  const float ir = inv_r[a];
                  const float e0 = (-1.0f - ca) * ir, e1 = (1.0f - ca) * ir;
                  ulo = fmaxf(ulo, fminf(e0, e1));
                  uhi = fminf(uhi, fmaxf(e0, e1));
                } else if (fabsf(ca) >= 1.001f) {
                  uhi = -10.0f;  // the whole column misses on this axis
                }
              }
              const int wlo = max(0, (int)ceilf(ulo + 3.5f - 1e-3f)), whi = min(7, (int)floorf(uhi + 3.5f + 1e-3f));
              for (int w = wlo; w <= whi; ++w) visit(d * 64 + h * 8 + w);
            }
          }
        } else {
#pragma unroll 1
          for (int vo = 0; vo < kVox; ++vo) visit(vo);
        }
      }
    }
    __syncthreads();  // rotA (dX) and taps are re-used by the next hypothesis
  }
  flush_pair(cur_b);
  // weight gradients: one atomic per accumulator and CTA
  {
    const int cg = t & 15, c = t >> 4;
#pragma unroll
    for (int view = 0; view < 3; ++view)
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const int k = view * 128 + c * 8 + kk;
        atomicAdd(g_W1 + (2 * cg) * kK + k, aW1[view][kk][0]);
        atomicAdd(g_W1 + (2 * cg + 1) * kK + k, aW1[view][kk][1]);
      }
    const int o = t >> 3, i0 = (t & 7) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(g_W2 + o * kO + i0 + i, aW2[i]);
    // db2: sum over the 64 positions inside the CTA first (threads with equal cg2 = t & 3)
    __syncthreads();
    float* red = sm.h1s;  // [4 cg2][8 oo][64 pos]
#pragma unroll
    for (int oo = 0; oo < 8; ++oo) red[(cg2 * 8 + oo) * kP + pos] = ab2[oo];
    __syncthreads();
    if (t < kO) {
      float s = 0.0f;
      for (int p = 0; p < kP; ++p) s += red[t * kP + p];
      atomicAdd(g_b2 + t, s);
    }
  }
}

int launch_score_bwd(const float* vol_src, const float* tgt_feat, const float* R, int r_per_pair,
                     const float* W1, const float* W2, const float* b2, const float* base,
                     const float* grad_scores, float* g_vol, float* g_tgt, float* g_W1, float* g_W2,
                     float* g_b2, int B, int64_t N, cudaStream_t s) {
  const int64_t total = (int64_t)B * N;
  if (total == 0) return AHV_OK;
  int dev = 0, sms = 0;
  AHV_CUDA_OK(cudaGetDevice(&dev));
  AHV_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  const size_t smem = sizeof(BwdSmem);
  AHV_CUDA_OK(cudaFuncSetAttribute(score_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  score_bwd_kernel<<<grid, kThreads, smem, s>>>(vol_src, tgt_feat, R, r_per_pair, W1, W2, b2, base, grad_scores,
                                                g_vol, g_tgt, g_W1, g_W2, g_b2, B, N);
  AHV_CUDA_OK(cudaGetLastError());
  return AHV_OK;
}

}  // namespace ahv
