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
// The adjoint of the resampling is a GATHER, not a scatter (floating-point shared-memory atomics would cost
// ~130k cycles per hypothesis): every INPUT voxel gets the list of the (output voxel, weight) contributions that reach
// it (counting sort over the <= 4096 contributions with integer shared-memory atomics and prefix sums), and every
// thread walks the lists of its two input voxels - exactly the output voxels that touch them, whatever the matrix R
// is - reading dX, which lies voxel-major with the channels innermost, as four LDS.128 per contribution.
#include <cstddef>

#include "ahv_head_fp32.cuh"

namespace ahv {

struct BwdSmem {
  Fp32Smem f;
  float dh2[kP * kH1Row];  // dL/dH2 [pos][36]
  float4 taps[kVox];       // (corner line, fx, fy, fz) of every output voxel of the current hypothesis
};
static_assert(sizeof(BwdSmem) <= 232448, "shared memory budget");

// dX (gradient w.r.t. the rotated volume) overlays rotA / rotT once dW1 is done, voxel-major with the 16 channels
// innermost: element (d, h, w, c) at d*kDxD + h*kDxH + w*kDxW + c floats - the adjoint reads a voxel's record as four
// LDS.128.  The pitches are 4*odd mod 32: the fold's lanes (which run over h or over w) meet distinct bank groups.  The
// adjoint's work list (<= 4096 entries of 8 bytes) follows it, into the dH1 buffer, which is dead by then as well.
constexpr int kDxW = 20, kDxH = 164, kDxD = 1312, kDxFloats = 8 * kDxD;
static_assert(offsetof(Fp32Smem, rotT) == offsetof(Fp32Smem, rotA) + sizeof(float) * kC * kRotC &&
              offsetof(Fp32Smem, h1s) == offsetof(Fp32Smem, rotT) + sizeof(float) * kC * kRotC, "rotA, rotT, h1s are contiguous");
static_assert((kDxFloats + 2 * kVox * 8) <= 2 * kC * kRotC + kP * kH1Row, "dX and the work list fit the dead buffers");
static_assert(2 * kVox * sizeof(int) <= sizeof(float) * kP * kH1Row, "counters fit the dH2 buffer");

// output voxel (z*64 + y*8 + x) a thread lists for the adjoint: lane bits -> x2 x1 y2 y1 z2, warp bits and j -> the rest
__device__ __forceinline__ int adjoint_out_voxel(int t, int j) {
  const int x = ((t & 3) << 1) | ((t >> 5) & 1), y = (((t >> 2) & 3) << 1) | ((t >> 6) & 1);
  const int z = (((t >> 4) & 1) << 2) | (((t >> 7) & 1) << 1) | j;
  return z * 64 + y * 8 + x;
}

__global__ void __launch_bounds__(kThreads, 1)
score_bwd_kernel(const float* __restrict__ vol_src, const float* __restrict__ tgt_feat,
                 const float* __restrict__ R, int r_per_pair, const float* __restrict__ W1,
                 const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ base,
                 const float* __restrict__ grad_scores, float* __restrict__ g_vol, float* __restrict__ g_tgt,
                 float* __restrict__ g_W1, float* __restrict__ g_W2, float* __restrict__ g_b2, int B, int64_t N,
                 const __half* __restrict__ h1_saved, const float* __restrict__ pair_inv) {
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
    if (h1_saved) {
      // saved-activation form: H1 = relu(conv1(A)) of this item was kept by the training forward (fp16, in the pair's
      // scaled units) - 4 KB read instead of 786 k FMA recomputed; A itself is still needed (dW1) and was just gathered
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(h1_saved + ((size_t)it * kP + pos) * kO + cg2 * 8));
      const float inv = __ldg(pair_inv + b);
      const __half2* hp = reinterpret_cast<const __half2*>(&q);
      float* dst = sm.h1s + pos * kH1Row + cg2 * 8;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f2 = __half22float2(hp[e]);
        dst[2 * e] = f2.x * inv;
        dst[2 * e + 1] = f2.y * inv;
      }
      __syncthreads();
    } else {
      __syncthreads();
      conv1_relu(sm);
      __syncthreads();
    }
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
    __syncthreads();  // rotA / rotT are free: they become dX, voxel-major with the channels innermost

    // ---------------- dA = dH1 W1, folded view by view into dX ----------------
    {
      const int pl = t & 15, c = t >> 4;  // positions pl + 16 j, channel c, all (view, kk)
      float* dX = sm.rotA + c;            // element (d, h, w, c) at d*kDxD + h*kDxH + w*kDxW + c
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
            if (view == 0) dX[p * kDxD + q * kDxH + kk * kDxW] = acc[j][kk];          // V'[c, d=p, h=q, w=kk]
            else if (view == 1) dX[p * kDxD + kk * kDxH + q * kDxW] += acc[j][kk];    // V'[c, d=p, h=kk, w=q]
            else dX[kk * kDxD + p * kDxH + q * kDxW] += acc[j][kk];                   // V'[c, d=kk, h=p, w=q]
          }
        }
        __syncthreads();
      }
    }

    // ---------------- dV_b += rotate^T(dX): adjoint of utils.py:113-131 as a gather ----------------
    // Output voxel vo feeds the 8 input voxels of its tap, corner(vo) + {0,1}^3, with the forward's weights.  Turned round:
    // every INPUT voxel gets the list of its (vo, weight) contributions - a counting sort over the <= 4096 contributions
    // that land inside the volume (shared-memory integer atomics whose return value is the place in the list, prefix
    // sums, 8-byte entries behind dX in the dead rotated-volume / dH1 buffers) - and the adjoint is one flat loop per
    // input voxel: exactly the contributing voxels, no geometry test, valid for any matrix R, fp32 weights recomputed from
    // the taps the forward gather recorded.
    {
      int* cnt = reinterpret_cast<int*>(bs.dh2);   // [512] contributions per input voxel
      int* start = cnt + kVox;                     // [512] exclusive prefix sum
      uint2* ent = reinterpret_cast<uint2*>(smem_raw + offsetof(BwdSmem, f) + offsetof(Fp32Smem, rotA) + sizeof(float) * kDxFloats);   // [<= 4096] (offset of dX[vo], weight), behind dX
      cnt[t] = 0;
      cnt[t + 256] = 0;
      __syncthreads();
      // the thread's two output voxels: two apart in x and y, four in z across the lanes of a warp, so that the taps of one
      // warp's voxels rarely meet in the same counter
      int tgt0[2];
      uint32_t inside[2], rank[2][4];
      float fx[2], fy[2], fz[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int vo = adjoint_out_voxel(t, j);
        const float4 tp = bs.taps[vo];
        const int line = __float_as_int(tp.x);   // ((z0+1)*10 + (y0+1))*10 + (x0+1), corner in [-1, 7]^3
        const int zl = line / (kHalo * kHalo), rl = line - zl * (kHalo * kHalo), yl = rl / kHalo, xl = rl - yl * kHalo;
        const int x0 = xl - 1, y0 = yl - 1, z0 = zl - 1;
        fx[j] = tp.y; fy[j] = tp.z; fz[j] = tp.w;
        tgt0[j] = z0 * 64 + y0 * 8 + x0;
        uint32_t m = 0;
        rank[j][0] = rank[j][1] = rank[j][2] = rank[j][3] = 0;
#pragma unroll
        for (int dlt = 0; dlt < 8; ++dlt) {
          const int x = x0 + (dlt & 1), y = y0 + ((dlt >> 1) & 1), z = z0 + (dlt >> 2);
          if ((unsigned)x < 8u && (unsigned)y < 8u && (unsigned)z < 8u) {
            m |= 1u << dlt;   // the atomic's return value is this contribution's place in its list (<= 511: 16 bits)
            rank[j][dlt >> 1] |= (uint32_t)atomicAdd(&cnt[tgt0[j] + (dlt >> 2) * 64 + ((dlt >> 1) & 1) * 8 + (dlt & 1)], 1) << (16 * (dlt & 1));
          }
        }
        inside[j] = m;
      }
      __syncthreads();
      {  // exclusive scan of the 512 counters: 2 per thread, warp scan, then the 8 warp totals
        const int c0 = cnt[2 * t], c1 = cnt[2 * t + 1];
        const int local = c0 + c1;
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, o);
          if ((t & 31) >= o) incl += up;
        }
        int* wtot = reinterpret_cast<int*>(sm.red);          // 8 warp totals (sm.red is free outside volume staging)
        if ((t & 31) == 31) wtot[t >> 5] = incl;
        __syncthreads();
        int basep = incl - local;
        for (int w = 0; w < (t >> 5); ++w) basep += wtot[w];
        start[2 * t] = basep;
        start[2 * t + 1] = basep + c0;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int vo = adjoint_out_voxel(t, j);
        const uint32_t off = (uint32_t)((vo >> 6) * kDxD + ((vo >> 3) & 7) * kDxH + (vo & 7) * kDxW);
#pragma unroll
        for (int dlt = 0; dlt < 8; ++dlt)
          if ((inside[j] >> dlt) & 1u) {
            const int dx = dlt & 1, dy = (dlt >> 1) & 1, dz = dlt >> 2;
            const float w = (dx ? fx[j] : 1.0f - fx[j]) * (dy ? fy[j] : 1.0f - fy[j]) * (dz ? fz[j] : 1.0f - fz[j]);
            const int v = tgt0[j] + dz * 64 + dy * 8 + dx;
            ent[start[v] + ((rank[j][dlt >> 1] >> (16 * (dlt & 1))) & 0xffffu)] = make_uint2(off, __float_as_uint(w));
          }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int vi = t + 256 * j;
        const int nn = cnt[vi];
        const uint2* e = ent + start[vi];
#pragma unroll 2
        for (int i = 0; i < nn; ++i) {
          const uint2 en = e[i];
          const float w = __uint_as_float(en.y);
          const float* src = sm.rotA + en.x;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 x = *reinterpret_cast<const float4*>(src + 4 * c4);
            aV[j][4 * c4] = fmaf(w, x.x, aV[j][4 * c4]); aV[j][4 * c4 + 1] = fmaf(w, x.y, aV[j][4 * c4 + 1]);
            aV[j][4 * c4 + 2] = fmaf(w, x.z, aV[j][4 * c4 + 2]); aV[j][4 * c4 + 3] = fmaf(w, x.w, aV[j][4 * c4 + 3]);
          }
        }
      }
    }
    __syncthreads();  // dX, the work list and the taps are re-used by the next hypothesis
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
                     float* g_b2, int B, int64_t N, cudaStream_t s, const void* h1_saved, const float* pair_inv) {
  const int64_t total = (int64_t)B * N;
  if (total == 0) return AHV_OK;
  int dev = 0, sms = 0;
  AHV_CUDA_OK(cudaGetDevice(&dev));
  AHV_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  const size_t smem = sizeof(BwdSmem);
  AHV_CUDA_OK(cudaFuncSetAttribute(score_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  score_bwd_kernel<<<grid, kThreads, smem, s>>>(vol_src, tgt_feat, R, r_per_pair, W1, W2, b2, base, grad_scores,
                                                g_vol, g_tgt, g_W1, g_W2, g_b2, B, N,
                                                static_cast<const __half*>(h1_saved), pair_inv);
  AHV_CUDA_OK(cudaGetLastError());
  return AHV_OK;
}

}  // namespace ahv
