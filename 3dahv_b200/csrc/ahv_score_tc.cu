// Fused hypothesis-and-verification kernels on tcgen05 tensor cores (AHV_MATH_TC*).
//
// Replace modules/model.py:186-193 (rotate_volume -> forward_3d2d -> correlate -> mean) for every
// (pair, hypothesis) without materialising anything in HBM: algorithmic HBM traffic is 36 B of
// rotation in and 4 B of score out (nothing at all when only the arg-max is requested).
//
// One persistent CTA per SM, 13 warps, warp-specialised (both kernels):
//   warps 0-7   GATHER   trilinear resampling (utils.py:113-131) of the source volume held in shared
//                        memory (zero halo, channel innermost).  One lane per output voxel, all 16
//                        channels.  The 16-byte chunks of a voxel line are visited in a per-lane
//                        rotated order and the two taps that sit in adjacent lines in bank-parity
//                        order, which makes every LDS.128 phase conflict-free by construction.
//                        Results are rounded to fp16 and become the tri-plane A operand of conv1.
//   warp  12    MMA      one elected thread issues tcgen05.mma (kind::f16, fp32 accumulate in TMEM):
//                        conv1 = 24 x (M=64,N=32,K=16) per hypothesis - slice (view,k) multiplies the
//                        16 channels at a fixed k, so the rearrange/cat of modules/modules.py:115-118
//                        is pure descriptor arithmetic - two hypotheses interleaved in the two 16-lane
//                        halves of each TMEM sub-partition; conv2 = 2 x (M=128,N=32,K=16) per pair.
//   warps 8-11  EPILOGUE tcgen05.ld D1 -> ReLU -> fp16 -> conv2 operand; tcgen05.ld D2 -> +bias ->
//                        L2 norm -> dot with the target features (registers) -> mean over 64
//                        positions -> score (+ running arg-max key, one atomicMax per CTA and pair).
// Pipelines (mbarrier): A-operand stages full/empty (3 deep), TMEM D1 full/empty, A2 full, D2 full.
//
//   score_tc_ts_kernel  (default)  view x of conv1 and conv2's A operand go register -> TMEM
//                        (tcgen05.st) and are consumed in the TS form of tcgen05.mma; only the YZ
//                        operand copy lives in shared memory; a stage is a hypothesis pair.
//   score_tc_kernel     (AHV_TC_VARIANT=ss)  both operand copies (YZ and X) and A2 in shared memory,
//                        SS form only; a stage is one hypothesis.
// Both exist for fp32-staged volumes (fp32 FFMA interpolation) and for 16-bit staged volumes
// (K16: bf16 inputs or AHV_MATH_TC_F16GATHER; x-pair lines, packed HFMA2 interpolation).
//
// Precision: normalisation and correlation are fp32; the two 1x1 convs use fp16 operands (10-bit
// mantissa = TF32-equivalent) with fp32 accumulation.  The source volume is pre-scaled per pair by a
// power of two (exact) so fp16 can neither overflow nor go subnormal; the scale is undone after
// conv2 (ReLU and the bias-free conv1 are positively homogeneous).
#include <cstdlib>

#include "ahv_head_fp32.cuh"

namespace ahv {

namespace tc {

// Optional in-kernel timeline (build with -DAHV_TIMELINE, read with ahv_diag_timeline): globaltimer stamps
// of the TS kernel's setup / first-tile milestones per CTA, used to attribute the fixed cost of a launch.
#ifdef AHV_TIMELINE
__device__ unsigned long long g_timeline[160][16];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define AHV_TL(i) g_timeline[blockIdx.x][i] = gtimer()
#else
#define AHV_TL(i) ((void)0)
#endif

constexpr int kGatherWarps = 8;
constexpr int kEpiWarp0 = 8;
constexpr int kMmaWarp = 12;
constexpr int kThreadsTC = 13 * 32;
constexpr int kStages = 3;

// ---- shared memory map (bytes) ----
constexpr int kW1Bytes = 24 * 1024;  // 24 MMA slices x [2 chalf][4 ngroup][8][8] fp16
constexpr int kW2Bytes = 2 * 1024;
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
constexpr int kA2Bytes = 8192;  // [4 kc][16 rowgroup][8][8] fp16

// 16-bit volumes are staged as "x-pair lines": for every (z, y, x0) of the halo'd grid one 64 B line
// holding BOTH x taps (x0, x0+1) x 16 channels as fp16 (exact for scaled bf16 inputs), chunk = tap*2 +
// chalf.  A tap pair is then four LDS.128, the gather reads half the bytes of the fp32 layout and
// interpolates with packed HFMA2 (fp16 accumulation adds <1e-4 relative to the scores; the bf16
// configuration's gate is 1e-2).
static_assert(kHalo * kHalo * 9 * 64 <= kVolSmemBytes, "pair lines (900 x 64 B) fit the volume region");

enum Bar { kFull = 0, kEmpty = 3, kD1Full = 6, kD1Empty = 8, kA2Full = 10, kD2Full = 12 };

// ---- PTX wrappers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {  // try_wait suspends the warp (no issue slots burnt) until the phase flips or the hint expires
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// The MMA warp runs converged (all 32 lanes execute the issue loop with warp-uniform values) and one
// elected lane issues each tcgen05.mma / commit.  Issuing under `if (lane == 0)` instead makes ptxas
// wrap every MMA in an ELECT / R2UR.BROADCAST waterfall loop (~16 instructions per MMA), which
// overloads the scheduler the MMA warp shares with two gather warps.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (elect_one())
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if (elect_one())
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_NONE, K-major: 8-row x 16-byte core
// matrices; LBO = byte distance between the two K chunks of one MMA, SBO = byte
// distance between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32, A=B=f16, both K-major, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// 1.0f the compiler cannot fold.  The rotation prefetch is loop-carried in registers; multiplying the
// values loaded BEFORE the loop by this makes every loop-entry value ALU-defined, so the first use at
// the loop top carries no scoreboard wait - otherwise that wait (on the scoreboard the just-issued
// prefetch of the NEXT tile also uses) stalls the warp for a full L2 round trip per tile.
__device__ __forceinline__ float opaque_one(int flag01) { return __uint_as_float(0x3f800000u + ((uint32_t)flag01 >> 8)); }

// Fused selection epilogue (modules/model.py:195-196): the last CTA to finish decodes the arg-max keys
// into (score, global index, sampled_R[pred_index]) per pair.  val == nullptr disables it.
struct Finalize {
  float* val;
  int64_t* idx;
  float* R_best;
  int64_t idx_offset;
  unsigned* counter;  // zero at kernel start (cleared with the keys), reset by the last CTA
};

struct Work {  // contiguous range of (pair, hypothesis) items of this CTA
  int64_t lo, hi, N;
};

// tile = two consecutive hypotheses of one pair; advance() returns false when exhausted.  All per-tile
// state is 32-bit (N <= 2^31-1 and the CTA's share of B*N < 2^32 are checked on the host): the iterator
// runs once per tile in every role.
struct TileIter {
  uint32_t n, N, left;  // next hypothesis within the pair, hypotheses per pair, items left for this CTA
  int b;
  uint32_t n0;  // first hypothesis of the tile
  int cnt;      // 1 or 2 valid hypotheses
  __device__ __forceinline__ TileIter(const Work& w) : N((uint32_t)w.N), left((uint32_t)(w.hi - w.lo)), n0(0), cnt(0) {
    b = (int)(w.lo / w.N);
    n = (uint32_t)(w.lo - (int64_t)b * w.N);
  }
  __device__ __forceinline__ bool advance() {
    if (left == 0) return false;
    if (n == N) { ++b; n = 0; }
    n0 = n;
    const uint32_t room = min(N - n, left);
    cnt = room >= 2 ? 2 : 1;
    n += cnt;
    left -= cnt;
    return true;
  }
  // (pair, first hypothesis, count) of the tile the next advance() will produce; at the very end the
  // current tile again (a harmless, valid address for the rotation prefetch)
  __device__ __forceinline__ void peek_tile(int& pb, uint32_t& pn, int& pc) const {
    if (left == 0) { pb = b; pn = n0; pc = cnt > 0 ? cnt : 1; return; }
    pb = b; pn = n;
    if (n == N) { ++pb; pn = 0; }
    pc = min(N - pn, left) >= 2 ? 2 : 1;
  }
  // (pair, hypothesis) of the item following this tile, or the tile's own first item at the very end
  __device__ __forceinline__ void peek(int& pb, uint32_t& pn) const {
    if (left == 0) { pb = b; pn = n0; }
    else if (n == N) { pb = b + 1; pn = 0; }
    else { pb = b; pn = n; }
  }
};

template <typename T>
__device__ __forceinline__ float ld_vol(const T* p);
template <>
__device__ __forceinline__ float ld_vol<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_vol<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// ---- one-time per CTA: W1/W2 (fp32, global) -> fp16 UMMA B-operand layouts in shared memory ------------
// conv1: 24 slices j = (view, kk), each [chalf][ngroup][n%8][c%8]; conv2: [kc][ngroup][8][8] behind them.
// Every CTA does this itself (52 KB of L2 reads) so that the scoring kernel depends on no preparation
// kernel.  It runs on the 5 non-gather warps (MMA + epilogue, tid = 0..159) while the gather warps already
// stage the first volume and resample the first tile; the MMA warp is released by named barrier 3.
constexpr int kPackThreads = kThreadsTC - kGatherWarps * 32;  // 160
__device__ __forceinline__ void pack_weights(unsigned char* wsm, const float* __restrict__ W1,
                                             const float* __restrict__ W2, int tid) {
  __half* w1h = reinterpret_cast<__half*>(wsm);
  constexpr int kV4 = kO * kK / 4;  // float4 = 4 consecutive kk of one (row, view, channel)
#pragma unroll 10
  for (int i = tid; i < kV4; i += kPackThreads) {
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(W1) + i);
    const int n = i / (kK / 4), k = (i - n * (kK / 4)) * 4;
    const int view = k >> 7, c = (k >> 3) & 15, kk = k & 7;
    __half* dst = w1h + (view * 8 + kk) * 512 + (c >> 3) * 256 + (n >> 3) * 64 + (n & 7) * 8 + (c & 7);
    dst[0] = __float2half_rn(w4.x); dst[512] = __float2half_rn(w4.y);
    dst[1024] = __float2half_rn(w4.z); dst[1536] = __float2half_rn(w4.w);
  }
  for (int i = tid; i < kO * kO; i += kPackThreads) {
    const int n = i / kO, k = i % kO;
    w1h[kW1Bytes / 2 + (k >> 3) * 256 + (n >> 3) * 64 + (n & 7) * 8 + (k & 7)] = __float2half_rn(__ldg(W2 + i));
  }
}

// Largest L1 norm of a W1 row (bounds |conv1 output| / max|V| for the pair scale), by the 8 gather warps:
// warp w sums rows 4w..4w+3 (48 coalesced loads per lane in flight, fixed summation order -> the same value
// in every CTA) and folds them into *l1max_bits (a non-negative float's bit pattern orders like the float).
__device__ __forceinline__ void w1_l1max(uint32_t* l1max_bits, const float* __restrict__ W1, int warp, int lane) {
  float wv[4][12];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < 12; ++j) wv[r][j] = __ldg(W1 + (4 * warp + r) * kK + lane + 32 * j);
  float m = 0.0f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float l1 = 0.0f;
#pragma unroll
    for (int j = 0; j < 12; ++j) l1 += fabsf(wv[r][j]);
    m = fmaxf(m, warp_sum(l1));
  }
  if (lane == 0) atomicMax(l1max_bits, __float_as_uint(m));
}

// ---- per pair: stage the source volume, pre-scaled by the pair's power-of-two scale -------------------
// Called by the 256 gather threads between two named barriers.  The volume is read once into registers, its
// max |V| reduced over the 8 warps, the scale s = 2^e chosen so that max|V|*s <= 2^12 and
// max|V|*s*L1max <= 2^14 (fp16 max 65504; exact, undone after conv2), then the scaled values are written in
// the gather's layout: fp32 lines [halo voxel][16 ch], or for 16-bit "x-pair lines" (every voxel is tap 0 of
// pair xh and tap 1 of pair xh-1).  Returns 1/s.
template <typename T, bool K16>
__device__ __forceinline__ float stage_pair_volume(unsigned char* vsm, const T* __restrict__ vg, uint32_t* l1max_bits,
                                                   const float* __restrict__ W1, bool first, float* red, int gtid) {
  float val[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    // fp32: task = gtid + 256*(i>>2) -> (voxel, 4-channel group jj), channel jj*4 + (i&3)
    // 16-bit: task = gtid + 256*(i>>3) -> (voxel, channel half), channel chalf*8 + (i&7)
    const int task = gtid + 256 * (K16 ? (i >> 3) : (i >> 2));
    const int v = task & 511, grp = task >> 9;
    const int ch = K16 ? grp * 8 + (i & 7) : grp * 4 + (i & 3);
    val[i] = ld_vol<T>(vg + ch * kVox + v);
  }
  // first pair of the CTA: the W1 row norms ride the same memory round trip as the volume
  if (first) w1_l1max(l1max_bits, W1, gtid >> 5, gtid & 31);
  float mx = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fabsf(val[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((gtid & 31) == 0) red[gtid >> 5] = mx;
  named_bar_sync(1, kGatherWarps * 32);
  const float l1max = __uint_as_float(*l1max_bits);
  float m = red[0];
#pragma unroll
  for (int w = 1; w < kGatherWarps; ++w) m = fmaxf(m, red[w]);
  float sc = 1.0f;
  if (m > 0.0f && isfinite(m)) {
    const float bound = fminf(4096.0f, 16384.0f / fmaxf(l1max, 1e-20f));
    int e = ilogbf(bound / m);  // floor(log2)
    e = max(-100, min(100, e));
    sc = scalbnf(1.0f, e);
  }
  if constexpr (!K16) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int task = gtid + 256 * it;
      const int v = task & 511, jj = task >> 9;
      const int z = v >> 6, y = (v >> 3) & 7, x = v & 7;
      const int line = ((z + 1) * kHalo + (y + 1)) * kHalo + (x + 1);
      *reinterpret_cast<float4*>(vsm + (line * kC + jj * 4) * 4) =
          make_float4(val[4 * it] * sc, val[4 * it + 1] * sc, val[4 * it + 2] * sc, val[4 * it + 3] * sc);
    }
  } else {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int task = gtid + 256 * it;
      const int v = task & 511, chalf = task >> 9;
      const int zh = (v >> 6) + 1, yh = ((v >> 3) & 7) + 1, xh = (v & 7) + 1;
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {  // bf16 inputs (8-bit mantissa) x power-of-two scale -> fp16 exactly; fp32 inputs are rounded
        const __half2 two = __floats2half2_rn(val[8 * it + 2 * e] * sc, val[8 * it + 2 * e + 1] * sc);
        pk[e] = *reinterpret_cast<const uint32_t*>(&two);
      }
      const uint4 q4 = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      unsigned char* row = vsm + ((zh * kHalo + yh) * 9) * 64;
      *reinterpret_cast<uint4*>(row + xh * 64 + chalf * 16) = q4;              // tap 0 of pair xh (xh <= 8)
      *reinterpret_cast<uint4*>(row + (xh - 1) * 64 + 32 + chalf * 16) = q4;   // tap 1 of pair xh-1
    }
  }
  return 1.0f / sc;
}

// ---- epilogue role, shared by the SS and TS kernels ---------------------------------------------------
// 4 warps, warp s = TMEM sub-partition s.  Per hypothesis pair (tile):
//   phase A: tcgen05.ld D1 -> ReLU (modules/modules.py:68) -> fp16 -> conv2's A operand, either as the
//            K-major core-matrix tile in shared memory (SS kernel) or straight into TMEM (TS kernel);
//   phase B (one tile later, after conv2): tcgen05.ld D2 -> undo the pair scale, + bias -> L2 norm with
//            F.normalize's eps (:122) -> dot with the target features (registers) -> mean over the 64
//            positions (modules/model.py:193) -> score, and the running arg-max key (:195).
template <bool kA2InTmem>
__device__ __forceinline__ void epilogue_role(const Work& work, int s, int lane, uint32_t tmem, uint32_t bar0,
                                              unsigned char* a2_smem, uint32_t tmem_a2, float* partial,
                                              const float* __restrict__ tgt_feat, const float* __restrict__ b2,
                                              const float* inv_ring, float* __restrict__ scores,
                                              u64* __restrict__ best_keys, int64_t N, int B,
                                              const float* __restrict__ R, int r_per_pair, const Finalize& fin) {
  const int slot = lane >> 4;            // which hypothesis of the tile
  const int pos = 16 * s + (lane & 15);  // position p*8+q of the folded plane
  const uint32_t row = 32 * s + lane;    // TMEM lane == row of the conv2 A operand
  float b2r[kO], tg[kO];
#pragma unroll
  for (int o = 0; o < kO; ++o) b2r[o] = __ldg(b2 + o);
  // the target features and the cleared arg-max keys come from the prologue grid; every other role of this
  // kernel is independent of it (no-op when the kernel was not launched programmatically dependent)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (s == 0 && lane == 0) AHV_TL(7);
  TileIter it(work);
  int cur_b = -1;
  float inv_s = 1.0f;
  uint32_t g = 0;
  int prev_b = 0, prev_cnt = 0;
  uint32_t prev_n0 = 0;
  float prev_inv = 1.0f;
  // arg-max fused into the epilogue (torch.max, modules/model.py:195): the two score-writing lanes keep a
  // running best key for the pair they are in and publish it with one atomicMax per (CTA, pair) - keys
  // order by score, ties by lowest index
  int key_b = -1;
  u64 key_best = 0;
  auto phase_b = [&](uint32_t gg, int pb, uint32_t pn0, int pcnt, float pinv) {
    const uint32_t gb = gg & 1, u = gg >> 1;
    mbar_wait(bar0 + (kD2Full + gb) * 8, u & 1);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(32 * s) << 16) + 64 + gb * 32, r);
    tmem_ld_wait();
    float ss = 0.0f, dt = 0.0f;
#pragma unroll
    for (int o = 0; o < kO; ++o) {
      const float v = fmaf(__uint_as_float(r[o]), pinv, b2r[o]);  // undo the pair scale, add bias
      ss = fmaf(v, v, ss);
      dt = fmaf(v, tg[o], dt);
    }
    float cosv = dt / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps (modules/modules.py:122)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) cosv += __shfl_xor_sync(0xffffffffu, cosv, o);  // 16 positions of this slot
    if ((lane & 15) == 0) partial[(gb * 2 + slot) * 4 + s] = cosv;
    tc_fence_before();
    named_bar_sync(2, 128);
    if (s == 0 && lane < pcnt) {
      const float* pp = partial + (gb * 2 + lane) * 4;
      const float tot = ((pp[0] + pp[1]) + pp[2]) + pp[3];  // fixed order: deterministic
      const float sc = tot * (1.0f / 64.0f);                // .mean(dim=-1)
      if (scores) scores[(size_t)pb * N + pn0 + lane] = sc;
      if (best_keys) {
        const u64 key = make_key(sc, (uint32_t)(pn0 + lane));
        if (pb != key_b) {
          if (key_b >= 0) atomicMax(best_keys + key_b, key_best);
          key_b = pb;
          key_best = key;
        } else if (key > key_best) {
          key_best = key;
        }
      }
    }
  };
  while (it.advance()) {
    const uint32_t gb = g & 1, u = g >> 1;
    // ---- phase A ----
    mbar_wait(bar0 + (kD1Full + gb) * 8, u & 1);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(32 * s) << 16) + gb * 32, r);
    tmem_ld_wait();
    uint32_t wq[16];  // ReLU -> fp16: the 32 channels of this thread's row
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const __half2 hh = __floats2half2_rn(fmaxf(__uint_as_float(r[2 * e]), 0.0f), fmaxf(__uint_as_float(r[2 * e + 1]), 0.0f));
      wq[e] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    if constexpr (kA2InTmem) {
      tmem_st16(tmem + ((uint32_t)(32 * s) << 16) + tmem_a2 + gb * 16, wq);  // 16 TMEM columns of conv2's A operand
      tmem_st_wait();
    } else {
      unsigned char* a2 = a2_smem + gb * kA2Bytes + row * 16;  // [kc][rowgroup][8][8] fp16 core matrices
#pragma unroll
      for (int kc = 0; kc < 4; ++kc)
        *reinterpret_cast<uint4*>(a2 + kc * 2048) = make_uint4(wq[4 * kc], wq[4 * kc + 1], wq[4 * kc + 2], wq[4 * kc + 3]);
      fence_proxy_async();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(bar0 + (kA2Full + gb) * 8);
      mbar_arrive(bar0 + (kD1Empty + gb) * 8);
    }
    if (s == 0 && lane == 0 && g == 0) AHV_TL(8);
    // ---- phase B of the previous tile ----
    if (g > 0) phase_b(g - 1, prev_b, prev_n0, prev_cnt, prev_inv);
    if (it.b != cur_b) {  // target features / scale of the tile just handed to conv2
      cur_b = it.b;
      inv_s = inv_ring[cur_b & 7];  // written by the gather role when it staged this pair
#pragma unroll
      for (int o = 0; o < kO; ++o) tg[o] = __ldg(tgt_feat + ((size_t)cur_b * kO + o) * kP + pos);
    }
    prev_b = it.b; prev_n0 = it.n0; prev_cnt = it.cnt; prev_inv = inv_s;
    ++g;
  }
  phase_b(g - 1, prev_b, prev_n0, prev_cnt, prev_inv);
  if (s == 0 && lane == 0) AHV_TL(9);
  if (best_keys && key_b >= 0) atomicMax(best_keys + key_b, key_best);
  if (best_keys && fin.val && s == 0) {
    __syncwarp();
    unsigned last = 0;
    if (lane == 0) {
      __threadfence();  // this CTA's keys before its ticket
      last = atomicAdd(fin.counter, 1u) == gridDim.x - 1;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {  // every other CTA has published its keys
      __threadfence();
      for (int b = lane; b < B; b += 32) {
        const u64 key = __ldcg(best_keys + b);
        const uint32_t n = key_index(key);
        fin.val[b] = key_score(key);
        fin.idx[b] = (int64_t)n + fin.idx_offset;
        if (fin.R_best) {
          const float* src = R + ((r_per_pair ? (size_t)b * N : 0) + (size_t)n) * 9;
#pragma unroll
          for (int e = 0; e < 9; ++e) fin.R_best[b * 9 + e] = __ldg(src + e);
        }
      }
      if (lane == 0) *fin.counter = 0u;
    }
  }
}

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

// ==============================================================================================
// TS variant (fp32 volumes): view x of conv1 never touches shared memory.  Each gather lane owns one
// accumulator row (slot, d, h) - i.e. one TMEM lane - walks the 8 voxels along w, and writes the
// fp16 channels of every voxel (a) once into the YZ copy (views y and z, K-major core matrices,
// 144-byte row pitch so the STS.64 are conflict-free for this lane map) and (b) with tcgen05.st
// straight into TMEM as the A operand of view x, which the MMA warp consumes in the TS form
// (tcgen05.mma [d], [a_tmem], b_desc).  A and D of one M=64 tile share the lane offset (0 or 16),
// verified by experiments/ts_mma_probe.cu.  Compared with the SS kernel this removes one 16 KB
// operand copy (stores) and one 16 KB operand read per hypothesis from the shared-memory pipe.
struct MapTS {
  static constexpr int yz_h = 144, yz_d = 8 * yz_h, yz_ch = 8 * yz_d, yz_bytes = 2 * yz_ch;  // 18432 per hypothesis
  static constexpr int tile_bytes = 2 * yz_bytes;
  static constexpr int off_vol = 0;
  static constexpr int off_w1 = off_vol + kVolSmemBytes;
  static constexpr int off_w2 = off_w1 + kW1Bytes;
  static constexpr int off_a = off_w2 + kW2Bytes;
  static constexpr int off_bar = off_a + kStages * tile_bytes;
  static constexpr int off_misc = off_bar + 16 * 8;
  static constexpr int smem_bytes = off_misc + 256;
  static constexpr int tmem_cols = 512;      // D1[2] 0..63, D2[2] 64..127, A_x[3] 128 + 64*stage, A2[2] 448 + 16*buf
  static constexpr int tmem_ax = 128;
  static constexpr int tmem_a2 = 448;
  static_assert(off_a % 128 == 0 && off_bar % 8 == 0, "align");
  static_assert(smem_bytes <= 232448, "shared memory budget");
};

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (elect_one())
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

template <typename T, bool K16>
__global__ void __launch_bounds__(kThreadsTC, 1)
score_tc_ts_kernel(const T* __restrict__ vol_src, const float* __restrict__ tgt_feat,
                   const float* __restrict__ R, int r_per_pair, const float* __restrict__ b2,
                   const float* __restrict__ base, const float* __restrict__ W1,
                   const float* __restrict__ W2, float* __restrict__ scores,
                   u64* __restrict__ best_keys, int B, int64_t N, Finalize fin) {
  extern __shared__ __align__(128) unsigned char smem[];
  using M = MapTS;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  Work work;
  {
    const int64_t total = (int64_t)B * N;
    work.lo = total * blockIdx.x / gridDim.x;
    work.hi = total * (blockIdx.x + 1) / gridDim.x;
    work.N = N;
  }
  if (work.lo >= work.hi) return;
  if (threadIdx.x == 0) AHV_TL(0);

  float* vol = reinterpret_cast<float*>(smem + M::off_vol);
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + M::off_bar;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + M::off_misc);
  float* partial = reinterpret_cast<float*>(smem + M::off_misc + 16);  // [2 tilebuf][2 slot][4 warps]
  float* sbase = reinterpret_cast<float*>(smem + M::off_misc + 96);    // 8 base coordinates
  uint32_t* l1max_bits = reinterpret_cast<uint32_t*>(smem + M::off_misc + 128);  // max L1 norm of a W1 row (float bits)
  float* red = reinterpret_cast<float*>(smem + M::off_misc + 144);        // 8 per-warp maxima (volume staging)
  float* inv_ring = reinterpret_cast<float*>(smem + M::off_misc + 176);   // 1/scale of pair b at [b & 7] (gather -> epilogue)

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
    tmem_alloc(smem_u32(tmem_slot), M::tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) AHV_TL(1);
  if (warp >= kGatherWarps) {  // MMA + epilogue warps: weights -> fp16 operand layouts, then release the MMA warp
    pack_weights(smem + M::off_w1, W1, W2, threadIdx.x - kGatherWarps * 32);
    fence_proxy_async();  // written through the generic proxy, UMMA reads through the async proxy
    named_bar_sync(3, kPackThreads);
    if (threadIdx.x == 256) AHV_TL(5);
  }

  if (warp < kGatherWarps) {
    // =========================== GATHER ===========================
    // lane -> accumulator row: slot = lane>>4 (hypothesis of the tile), d = 2*sub + ((lane>>3)&1), h = lane&7;
    // the warp's TMEM sub-partition is warp&3, warps w and w+4 split the w axis
    const int sub = warp & 3, whalf = warp >> 2;
    const int slot = lane >> 4, dlo = (lane >> 3) & 1, h_ = lane & 7, d = 2 * sub + dlo;
    const int pf = (lane >> 2) & 1;            // bank parity this lane reads first
    const int rot = ((lane & 3) + dlo) & 3;    // chunk rotation
    const float by = sbase[h_], bz = sbase[d];
    uint32_t koff[4], syz[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      // chunk visited at step t: a rotation for fp32, an XOR swizzle for 16-bit lines (chunk = tap*2 + chalf, so
      // the tap of step t is (t>>1)^(rot>>1): steps 0,1 share one x tap, steps 2,3 the other - compile-time)
      const int ck = K16 ? (rot ^ t) : ((rot + t) & 3);
      koff[t] = ck * 16;
      if constexpr (!K16)  // fp32: chunk ck = channels 4ck..4ck+3 (8 B of fp16)
        syz[t] = slot * M::yz_bytes + (ck >> 1) * M::yz_ch + (ck & 1) * 8 + d * M::yz_d + h_ * M::yz_h;
      else                 // 16-bit: accumulator t&1 holds channel half (rot^t)&1 (16 B of fp16); t < 2 used
        syz[t] = slot * M::yz_bytes + ((rot ^ t) & 1) * M::yz_ch + d * M::yz_d + h_ * M::yz_h;
    }
    const unsigned char* volb = smem + M::off_vol;
    const int gtid = threadIdx.x;
    TileIter it(work);
    int cur_b = -1;
    uint32_t g = 0;
    float Rn[9];
    auto fetch_R = [&](int fb, uint32_t fn) {
      const float* Rg = R + (r_per_pair ? ((size_t)fb * N + fn) : (size_t)fn) * 9;
#pragma unroll
      for (int e = 0; e < 9; ++e) Rn[e] = __ldg(Rg + e);
    };
    {
      int fb; uint32_t fn; int fc;
      it.peek_tile(fb, fn, fc);
      fetch_R(fb, fn + (slot < fc ? slot : 0));
      // first pair of this CTA: its volume (and the W1 row norms) ride the same memory round trip as the
      // first rotations
      cur_b = fb;
      const float inv = stage_pair_volume<T, K16>(smem + M::off_vol, vol_src + (size_t)fb * kC * kVox, l1max_bits, W1,
                                                  true, red, gtid);
      if (gtid == 0) inv_ring[fb & 7] = inv;
      named_bar_sync(1, kGatherWarps * 32);
      if (threadIdx.x == 0) AHV_TL(3);
#pragma unroll
      for (int e = 0; e < 9; ++e) Rn[e] *= opaque_one(r_per_pair);
    }
    // ---- 16-bit path: coordinates/weights of a voxel pair (w0, w0+1), computed one step AHEAD of its loads
    // (software pipeline across voxel pairs and across tiles) so that the serial coordinate chains of one
    // pair overlap the shared-memory traffic of the previous one ----
    struct VoxPair16 {
      const unsigned char* pa[2];
      const unsigned char* pb[2];
      __half2 wg[2][2][4];  // [voxel][x tap of steps {0,1} / {2,3}][y-z corner]
    };
    auto coords16 = [&](const float (&Rq)[9], int w0, VoxPair16& o) {
      // grid = R @ (x, y, z): y and z are fixed per lane, x walks with w
      const float pgx = fmaf(Rq[2], bz, Rq[1] * by), pgy = fmaf(Rq[5], bz, Rq[4] * by), pgz = fmaf(Rq[8], bz, Rq[7] * by);
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const float bx = sbase[w0 + v];
        float ix = unnorm(fmaf(Rq[0], bx, pgx)), iy = unnorm(fmaf(Rq[3], bx, pgy)), iz = unnorm(fmaf(Rq[6], bx, pgz));
        ix = fminf(fmaxf(ix, -1.0f), 8.0f); iy = fminf(fmaxf(iy, -1.0f), 8.0f); iz = fminf(fmaxf(iz, -1.0f), 8.0f);
        const float x0 = fminf(floorf(ix), 7.0f), y0 = fminf(floorf(iy), 7.0f), z0 = fminf(floorf(iz), 7.0f);
        const float fx = ix - x0, fy = iy - y0, fz = iz - z0;
        const int pline = (((int)z0 + 1) * kHalo + ((int)y0 + 1)) * 9 + ((int)x0 + 1);
        const int swapy = (pline ^ pf) & 1;  // first y tap = the one whose line has bank parity pf (9 is odd)
        o.pa[v] = volb + (pline + swapy * 9) * 64;
        o.pb[v] = volb + (pline + (1 - swapy) * 9) * 64;
        const float wya = swapy ? fy : 1.0f - fy, wyb = swapy ? 1.0f - fy : fy;
        const __half2 wy2[2] = {__float2half2_rn(wya), __float2half2_rn(wyb)};
        const __half2 wz2[2] = {__float2half2_rn(1.0f - fz), __float2half2_rn(fz)};
        // x tap read at steps {0,1} is (rot>>1), at steps {2,3} the other one
        const __half2 wxf = __float2half2_rn((rot & 2) ? fx : 1.0f - fx), wxs = __float2half2_rn((rot & 2) ? 1.0f - fx : fx);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const __half2 w4 = __hmul2(wy2[c >> 1], wz2[c & 1]);
          o.wg[v][0][c] = __hmul2(wxf, w4);
          o.wg[v][1][c] = __hmul2(wxs, w4);
        }
      }
    };
    // trilinear taps of the pair (16 LDS.128 per voxel, packed HFMA2), then the operand stores
    auto sample16 = [&](const VoxPair16& q, int w0, unsigned char* st, uint32_t ax) {
      constexpr int kDz = kHalo * 9 * 64;
      __half2 acc[2][2][4];
#pragma unroll
      for (int v = 0; v < 2; ++v)
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) acc[v][0][e2] = acc[v][1][e2] = __float2half2_rn(0.0f);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        uint4 buf[2][4];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          buf[v][0] = *reinterpret_cast<const uint4*>(q.pa[v] + koff[t]);
          buf[v][1] = *reinterpret_cast<const uint4*>(q.pa[v] + koff[t] + kDz);
          buf[v][2] = *reinterpret_cast<const uint4*>(q.pb[v] + koff[t]);
          buf[v][3] = *reinterpret_cast<const uint4*>(q.pb[v] + koff[t] + kDz);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const __half2 wg = q.wg[v][t >> 1][c];
            const uint4 q4 = buf[v][c];
            const uint32_t wd[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2)
              acc[v][t & 1][k2] = __hfma2(wg, *reinterpret_cast<const __half2*>(&wd[k2]), acc[v][t & 1][k2]);
          }
      }
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int w = w0 + v;
        uint32_t a0[4], a1[4];
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          a0[k2] = *reinterpret_cast<const uint32_t*>(&acc[v][0][k2]);
          a1[k2] = *reinterpret_cast<const uint32_t*>(&acc[v][1][k2]);
        }
        *reinterpret_cast<uint4*>(st + syz[0] + w * 16) = make_uint4(a0[0], a0[1], a0[2], a0[3]);
        *reinterpret_cast<uint4*>(st + syz[1] + w * 16) = make_uint4(a1[0], a1[1], a1[2], a1[3]);
        // view x: acc[.][0] holds channel half (rot&1); put the halves in channel order and store to TMEM
        const bool sw = rot & 1;
        uint32_t regs[8];
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          regs[k2] = sw ? a1[k2] : a0[k2];
          regs[4 + k2] = sw ? a0[k2] : a1[k2];
        }
        tmem_st8(ax + w * 8, regs);
      }
    };
    VoxPair16 vp_a, vp_b;
    if constexpr (K16) coords16(Rn, whalf * 4, vp_a);  // first pair of the first tile
    while (it.advance()) {
      if (it.b != cur_b) {
        named_bar_sync(1, kGatherWarps * 32);
        const T* vg = vol_src + (size_t)it.b * kC * kVox;
        const float inv = stage_pair_volume<T, K16>(smem + M::off_vol, vg, l1max_bits, W1, false, red, gtid);
        if (gtid == 0) inv_ring[it.b & 7] = inv;
        named_bar_sync(1, kGatherWarps * 32);
        cur_b = it.b;
      }
      float Rr[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) Rr[e] = Rn[e];
      {  // prefetch this lane's rotation of the next tile
        int nb; uint32_t nn; int nc;
        it.peek_tile(nb, nn, nc);
        fetch_R(nb, nn + (slot < nc ? slot : 0));
      }
      const uint32_t stage = g % kStages, use = g / kStages;
      if (use > 0) mbar_wait(bar0 + (kEmpty + stage) * 8, (use - 1) & 1);
      unsigned char* st = smem + M::off_a + stage * M::tile_bytes;
      const uint32_t ax = tmem + ((uint32_t)(32 * sub) << 16) + M::tmem_ax + stage * 64;
      if constexpr (!K16) {
        // grid = R @ (x, y, z): y and z are fixed per lane, x walks with w
        const float pgx = fmaf(Rr[2], bz, Rr[1] * by), pgy = fmaf(Rr[5], bz, Rr[4] * by), pgz = fmaf(Rr[8], bz, Rr[7] * by);
#pragma unroll 1
        for (int wi = 0; wi < 4; ++wi) {
          const int w = whalf * 4 + wi;
          const float bx = sbase[w];
          float ix = unnorm(fmaf(Rr[0], bx, pgx)), iy = unnorm(fmaf(Rr[3], bx, pgy)), iz = unnorm(fmaf(Rr[6], bx, pgz));
          ix = fminf(fmaxf(ix, -1.0f), 8.0f); iy = fminf(fmaxf(iy, -1.0f), 8.0f); iz = fminf(fmaxf(iz, -1.0f), 8.0f);
          const float x0 = fminf(floorf(ix), 7.0f), y0 = fminf(floorf(iy), 7.0f), z0 = fminf(floorf(iz), 7.0f);
          const float fx = ix - x0, fy = iy - y0, fz = iz - z0;
          const int line = (((int)z0 + 1) * kHalo + ((int)y0 + 1)) * kHalo + ((int)x0 + 1);
          const int swap = (line ^ pf) & 1;
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
          float4 buf[2][8];
  #pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int off = ((c >> 1) * kHalo * kHalo + (c & 1) * kHalo) * 64;
            buf[0][2 * c] = *reinterpret_cast<const float4*>(pa + koff[0] + off);
            buf[0][2 * c + 1] = *reinterpret_cast<const float4*>(pb + koff[0] + off);
          }
          uint2 pk[4];
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
              const float4 a = buf[t & 1][2 * c], q = buf[t & 1][2 * c + 1];
              acc.x = fmaf(wa[c], a.x, acc.x); acc.y = fmaf(wa[c], a.y, acc.y);
              acc.z = fmaf(wa[c], a.z, acc.z); acc.w = fmaf(wa[c], a.w, acc.w);
              acc.x = fmaf(wb[c], q.x, acc.x); acc.y = fmaf(wb[c], q.y, acc.y);
              acc.z = fmaf(wb[c], q.z, acc.z); acc.w = fmaf(wb[c], q.w, acc.w);
            }
            const __half2 lo = __floats2half2_rn(acc.x, acc.y), hi2 = __floats2half2_rn(acc.z, acc.w);
            pk[t].x = *reinterpret_cast<const uint32_t*>(&lo);
            pk[t].y = *reinterpret_cast<const uint32_t*>(&hi2);
            *reinterpret_cast<uint2*>(st + syz[t] + w * 16) = pk[t];  // YZ copy: row w of core matrix (d,h,chalf)
          }
          // view x: un-rotate the chunks (chunk c was produced at step (c - rot) & 3) and store the 16 channels
          // of this voxel as K slice w of this lane's accumulator row in TMEM
          uint2 s1[4], o4[4];
          const bool r1 = rot & 1, r2 = rot & 2;
  #pragma unroll
          for (int c = 0; c < 4; ++c) {
            s1[c].x = r1 ? pk[(c + 3) & 3].x : pk[c].x;
            s1[c].y = r1 ? pk[(c + 3) & 3].y : pk[c].y;
          }
  #pragma unroll
          for (int c = 0; c < 4; ++c) {
            o4[c].x = r2 ? s1[(c + 2) & 3].x : s1[c].x;
            o4[c].y = r2 ? s1[(c + 2) & 3].y : s1[c].y;
          }
          const uint32_t regs[8] = {o4[0].x, o4[0].y, o4[1].x, o4[1].y, o4[2].x, o4[2].y, o4[3].x, o4[3].y};
          tmem_st8(ax + w * 8, regs);
        }
      } else {
        // 16-bit staged volume (x-pair lines, packed HFMA2): pair 0 was prepared during the previous tile
        sample16(vp_a, whalf * 4, st, ax);
        coords16(Rr, whalf * 4 + 2, vp_b);   // overlaps the loads of pair 0
        sample16(vp_b, whalf * 4 + 2, st, ax);
        coords16(Rn, whalf * 4, vp_a);       // next tile's pair 0 (Rn = its prefetched rotation) overlaps pair 1
      }
      tmem_st_wait();
      fence_proxy_async();  // YZ stores -> async proxy
      tc_fence_before();    // TMEM stores -> tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + (kFull + stage) * 8);
      if (threadIdx.x == 0 && g == 0) AHV_TL(4);
      ++g;
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA ISSUER (converged warp) ===========================
    {
      constexpr uint32_t idesc1 = instr_desc(64, 32), idesc2 = instr_desc(128, 32);
      const uint32_t w1s = s_base + M::off_w1, w2s = s_base + M::off_w2;
      TileIter it(work);
      uint32_t g = 0;
      auto conv2 = [&](uint32_t gg) {
        const uint32_t gb = gg & 1, u = gg >> 1;
        mbar_wait(bar0 + (kA2Full + gb) * 8, u & 1);
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < 2; ++i)  // conv2: A = ReLU(conv1) rows in TMEM (written by the epilogue), M = 128
          umma_f16_ts(tmem + 64 + gb * 32, tmem + M::tmem_a2 + gb * 16 + i * 8, smem_desc(w2s + i * 1024, 512, 128), idesc2, i);
        umma_commit(bar0 + (kD2Full + gb) * 8);
      };
      while (it.advance()) {
        const uint32_t gb = g & 1, u = g >> 1;
        if (u > 0) mbar_wait(bar0 + (kD1Empty + gb) * 8, (u - 1) & 1);
        const uint32_t stage = g % kStages, use = g / kStages;
        mbar_wait(bar0 + (kFull + stage) * 8, use & 1);
        tc_fence_after();
        // view x from TMEM, both hypotheses of the tile in one M=128 MMA per K slice (w=kk, c): A and D of a
        // TS-form MMA share the lane, so row i of the M=128 tile is TMEM lane i - exactly the rows the
        // two interleaved M=64 tiles below accumulate into.  Halves view x's W1 (B operand) reads.
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16_ts(tmem + gb * 32, tmem + M::tmem_ax + stage * 64 + kk * 8, smem_desc(w1s + kk * 1024, 512, 128), idesc2, kk);
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const uint32_t lane_off = (uint32_t)(16 * sl) << 16;
          const uint32_t d1 = tmem + lane_off + gb * 32;
          const uint32_t a = s_base + M::off_a + stage * M::tile_bytes + sl * M::yz_bytes;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // view y: rows (d,w), K slice = (h=kk, c)
            umma_f16(d1, smem_desc(a + kk * M::yz_h, M::yz_ch, M::yz_d), smem_desc(w1s + (8 + kk) * 1024, 512, 128), idesc1, 1);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // view z: rows (h,w), K slice = (d=kk, c)
            umma_f16(d1, smem_desc(a + kk * M::yz_d, M::yz_ch, M::yz_h), smem_desc(w1s + (16 + kk) * 1024, 512, 128), idesc1, 1);
        }
        umma_commit(bar0 + (kEmpty + stage) * 8);
        umma_commit(bar0 + (kD1Full + gb) * 8);
        if (lane == 0 && g == 0) AHV_TL(6);
        if (g > 0) conv2(g - 1);
        ++g;
      }
      conv2(g - 1);
    }
    __syncwarp();
  } else {
    // =========================== EPILOGUE ===========================
    epilogue_role<true>(work, warp - kEpiWarp0, lane, tmem, bar0, nullptr, M::tmem_a2, partial, tgt_feat, b2, inv_ring,
                        scores, best_keys, N, B, R, r_per_pair, fin);
  }

  if (threadIdx.x == 0) AHV_TL(10);
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, M::tmem_cols);
  }
  if (threadIdx.x == 0) AHV_TL(11);
}

// ---- prologue: target features forward_3d2d(vol_tgt[b]) (modules/model.py:191) in fp32, + arg-max key clear ----
// grid (16, B), 128 threads: CTA (j, b) computes the 4 positions (p = j>>1, q = 4*(j&1) .. +3) of pair b for
// all 32 channels.  6.7 KB of static shared memory and <= 96 registers per thread, so these CTAs co-reside
// with the scoring kernel's CTAs (201 KB, 416 x 128 registers): the scoring kernel is launched
// programmatically dependent and only its epilogue warps wait for this grid.
//   A[q][k]   tri-plane operand of the 4 positions (modules/modules.py:115-118):
//             k = c*8+kk       -> V[c, p, q, kk]   (view x)
//             k = 128+c*8+kk   -> V[c, p, kk, q]   (view y)
//             k = 256+c*8+kk   -> V[c, kk, p, q]   (view z)
//   conv1: warp w owns output rows o = w, w+4, ..., w+28; lanes stride k (coalesced W1 row reads, the next
//          row's 12 values in flight while this row is reduced); ReLU -> h1[q][o].
//   conv2 + bias + L2 normalise: warp = position, lane = output channel.
constexpr int kTgtThreads = 128, kTgtCtasPerPair = 16;
__global__ void __launch_bounds__(kTgtThreads, 4)
tc_tgt_feat_kernel(const float* __restrict__ vol_tgt, const float* __restrict__ W1, const float* __restrict__ W2,
                   const float* __restrict__ b2, u64* __restrict__ best_keys, unsigned* __restrict__ done_counter,
                   float* __restrict__ tgt_feat, int B) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // let the scoring kernel start its setup now
  __shared__ float A[4][kK];
  __shared__ float h1[4][kO + 1];
  const int p = blockIdx.x >> 1, q0 = (blockIdx.x & 1) * 4, b = blockIdx.y;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  if (best_keys && blockIdx.x == 0 && t == 0) best_keys[b] = 0ull;
  if (done_counter && blockIdx.x == 0 && b == 0 && t == 32) *done_counter = 0u;
  const float* V = vol_tgt + (size_t)b * kC * kVox;
  float wr[2][12];
#pragma unroll
  for (int j = 0; j < 12; ++j) wr[0][j] = __ldg(W1 + warp * kK + lane + 32 * j);
#pragma unroll
  for (int i = 0; i < 4 * kK / kTgtThreads; ++i) {
    const int e = t + kTgtThreads * i;
    const int ql = e / kK, k = e - ql * kK, q = q0 + ql;
    const int view = k >> 7, c = (k >> 3) & 15, kk = k & 7;
    const int d = view == 2 ? kk : p, h = view == 0 ? q : (view == 1 ? kk : p), w = view == 0 ? kk : q;
    A[ql][k] = __ldg(V + c * kVox + d * 64 + h * 8 + w);
  }
  __syncthreads();
#pragma unroll 1
  for (int r = 0; r < 8; r += 2) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {  // two rows per trip so the double buffer is statically indexed
      const int o = warp + 4 * (r + u);
      if (r + u + 1 < 8) {
#pragma unroll
        for (int j = 0; j < 12; ++j) wr[u ^ 1][j] = __ldg(W1 + (o + 4) * kK + lane + 32 * j);
      }
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int j = 0; j < 12; ++j)
#pragma unroll
        for (int ql = 0; ql < 4; ++ql) acc[ql] = fmaf(wr[u][j], A[ql][lane + 32 * j], acc[ql]);
#pragma unroll
      for (int ql = 0; ql < 4; ++ql) {
        const float v = warp_sum(acc[ql]);
        if (lane == ql) h1[ql][o] = fmaxf(v, 0.0f);  // ReLU (modules/modules.py:68)
      }
    }
  }
  __syncthreads();
  {
    const int ql = warp, o2 = lane;  // 4 warps = 4 positions, lanes = output channels
    float v = __ldg(b2 + o2);
#pragma unroll
    for (int o = 0; o < kO; o += 4) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(W2 + o2 * kO + o));
      v = fmaf(w4.x, h1[ql][o], v); v = fmaf(w4.y, h1[ql][o + 1], v);
      v = fmaf(w4.z, h1[ql][o + 2], v); v = fmaf(w4.w, h1[ql][o + 3], v);
    }
    const float ss = warp_sum(v * v);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps (modules/modules.py:122)
    tgt_feat[((size_t)b * kO + o2) * kP + p * 8 + q0 + ql] = v * inv;
  }
}

struct Scratch {  // carve-up of the tensor-core path's workspace
  __half* w_packed;
  float2* pair_scale;
  u64* best_keys;
  unsigned* counter;
  float* tgt_feat;
};
__host__ inline size_t scratch_bytes(int B) {
  return (size_t)(kW1Bytes + kW2Bytes) + (size_t)B * (sizeof(float2) + sizeof(u64)) + (size_t)B * kO * kP * 4 + 256;
}
__host__ inline Scratch carve(void* ws, int B) {
  unsigned char* p = static_cast<unsigned char*>(ws);
  Scratch sc;
  sc.w_packed = reinterpret_cast<__half*>(p);
  p += kW1Bytes + kW2Bytes;
  sc.tgt_feat = reinterpret_cast<float*>(p);
  p += (size_t)B * kO * kP * 4;
  sc.pair_scale = reinterpret_cast<float2*>(p);
  p += (size_t)B * sizeof(float2);
  sc.best_keys = reinterpret_cast<u64*>(p);
  sc.counter = reinterpret_cast<unsigned*>(sc.best_keys + B);  // inside the 256-byte tail of scratch_bytes
  return sc;
}

template <typename KernelT, typename... Args>
static int launch_pdl(KernelT kernel, unsigned grid, size_t smem, cudaStream_t s, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreadsTC);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  AHV_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, args...));
  return AHV_OK;
}

// Launches: [target features + key clear] -> scoring.  The scoring kernel packs the weights and derives
// the per-pair scales itself; when the prologue runs, the scoring kernel is launched programmatically
// dependent on it (its setup, weight packing, volume staging and first gathers overlap the prologue; only
// the epilogue warps wait for the target features).
template <typename T, bool K16>
int launch_typed(const T* vol_src, const float* vol_tgt, const float* tgt_feat_in, const float* R,
                 int r_per_pair, const float* W1, const float* W2, const float* b2, const float* base,
                 float* scores, bool want_argmax, int B, int64_t N, const Scratch& sc, Finalize fin, cudaStream_t s) {
  int dev = 0, sms = 0;
  AHV_CUDA_OK(cudaGetDevice(&dev));
  AHV_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const bool prologue = vol_tgt != nullptr;
  if (prologue) {
    const int st = launch_tgt_feat(vol_tgt, W1, W2, b2, want_argmax ? sc.best_keys : nullptr, sc.tgt_feat, B, s);
    if (st != AHV_OK) return st;
  } else if (want_argmax) {
    AHV_CUDA_OK(cudaMemsetAsync(sc.best_keys, 0, (size_t)B * sizeof(u64) + sizeof(unsigned), s));
  }
  fin.counter = sc.counter;
  // a tile is two hypotheses: do not spread tiny problems over more CTAs than tiles
  const int64_t tiles = ((int64_t)B * N + 1) / 2;
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  if (((int64_t)B * N) / grid >= 0xffffffffLL) return AHV_EINVAL;  // per-CTA tile iterator is 32-bit
  const float* tgt = prologue ? sc.tgt_feat : tgt_feat_in;
  u64* keys = want_argmax ? sc.best_keys : nullptr;
  static const bool use_ts = [] { const char* e = getenv("AHV_TC_VARIANT"); return !(e && e[0] == 's'); }();
  static const bool use_pdl = [] { const char* e = getenv("AHV_PDL"); return !(e && e[0] == '0'); }();
  if (use_ts) {
    AHV_CUDA_OK(cudaFuncSetAttribute(score_tc_ts_kernel<T, K16>, cudaFuncAttributeMaxDynamicSharedMemorySize, MapTS::smem_bytes));
    return launch_pdl(score_tc_ts_kernel<T, K16>, grid, MapTS::smem_bytes, s, prologue && use_pdl, vol_src, tgt, R, r_per_pair,
                      b2, base, W1, W2, scores, keys, B, N, fin);
  }
  constexpr int kSmemBytes = Map<K16>::smem_bytes;
  AHV_CUDA_OK(cudaFuncSetAttribute(score_tc_kernel<T, K16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  return launch_pdl(score_tc_kernel<T, K16>, grid, kSmemBytes, s, prologue && use_pdl, vol_src, tgt, R, r_per_pair, b2, base,
                    W1, W2, scores, keys, B, N, fin);
}

}  // namespace tc

// forward_3d2d for B plain volumes (+ optional clear of B arg-max keys): the fused step's prologue, also
// behind ahv_forward_3d2d for per-pair sizes
int launch_tgt_feat(const float* vol_tgt, const float* W1, const float* W2, const float* b2,
                    unsigned long long* clear_keys, float* feat, int B, cudaStream_t s) {
  // same shared-memory carve-out as the scoring kernel: an SM cannot change its L1/shared split while a CTA
  // is resident, so with the default (small) carve-out the dependent scoring CTAs could not join these SMs
  AHV_CUDA_OK(cudaFuncSetAttribute(tc::tc_tgt_feat_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   (int)cudaSharedmemCarveoutMaxShared));
  for (int b0 = 0; b0 < B; b0 += 65535) {  // gridDim.y limit
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    tc::tc_tgt_feat_kernel<<<dim3(tc::kTgtCtasPerPair, nb), tc::kTgtThreads, 0, s>>>(
        vol_tgt + (size_t)b0 * kC * kVox, W1, W2, b2, clear_keys ? clear_keys + b0 : nullptr,
        clear_keys && b0 == 0 ? reinterpret_cast<unsigned*>(clear_keys + B) : nullptr,  // Scratch::counter
        feat + (size_t)b0 * kO * kP, nb);
    AHV_CUDA_OK(cudaGetLastError());
  }
  return AHV_OK;
}

size_t score_tc_workspace_bytes(int B, int64_t N) {
  (void)N;
  return tc::scratch_bytes(B);
}

float* scratch_tgt_feat(void* ws, int B) { return tc::carve(ws, B).tgt_feat; }

// dispatch on (volume dtype, gather precision): bf16 volumes always use the 16-bit staged gather
// (exact staging); fp32 volumes use it only when the caller opts in (AHV_MATH_TC_F16GATHER)
static int dispatch(const void* vol_src, int vol_dtype, bool f16_gather, const float* vol_tgt,
                    const float* tgt_feat, const float* R, int r_per_pair, const float* W1, const float* W2,
                    const float* b2, const float* base, float* scores, bool want_argmax, int B, int64_t N,
                    const tc::Scratch& sc, const tc::Finalize& fin, cudaStream_t s) {
  if (vol_dtype == AHV_VOL_BF16)
    return tc::launch_typed<__nv_bfloat16, true>((const __nv_bfloat16*)vol_src, vol_tgt, tgt_feat, R, r_per_pair, W1,
                                                 W2, b2, base, scores, want_argmax, B, N, sc, fin, s);
  if (f16_gather)
    return tc::launch_typed<float, true>((const float*)vol_src, vol_tgt, tgt_feat, R, r_per_pair, W1, W2, b2, base,
                                         scores, want_argmax, B, N, sc, fin, s);
  return tc::launch_typed<float, false>((const float*)vol_src, vol_tgt, tgt_feat, R, r_per_pair, W1, W2, b2, base,
                                        scores, want_argmax, B, N, sc, fin, s);
}

// scores only (target features supplied by the caller)
int launch_score_tc(const void* vol_src, int vol_dtype, const float* tgt_feat, const float* R,
                    int r_per_pair, const float* W1, const float* W2, const float* b2,
                    const float* base, float* scores, int B, int64_t N, void* ws, size_t ws_bytes,
                    cudaStream_t s, bool f16_gather) {
  if ((int64_t)B * N == 0) return AHV_OK;
  if (ws_bytes < tc::scratch_bytes(B)) return AHV_EWORKSPACE;
  return dispatch(vol_src, vol_dtype, f16_gather, nullptr, tgt_feat, R, r_per_pair, W1, W2, b2, base, scores, false,
                  B, N, tc::carve(ws, B), tc::Finalize{nullptr, nullptr, nullptr, 0, nullptr}, s);
}

// the whole verification step with arg-max selection in two launches: prologue (target features, key
// clear) -> fused scoring + arg-max, whose last CTA also decodes the winners (no finalize launch)
int launch_verify_tc_argmax(const void* vol_src, int vol_dtype, const float* vol_tgt, const float* R,
                            int r_per_pair, const float* W1, const float* W2, const float* b2,
                            const float* base, float* scores, float* best_val, int64_t* best_idx,
                            float* R_best, int64_t idx_offset, int B, int64_t N, void* ws, size_t ws_bytes,
                            cudaStream_t s, bool f16_gather) {
  if ((int64_t)B * N == 0) return AHV_OK;
  if (ws_bytes < tc::scratch_bytes(B)) return AHV_EWORKSPACE;
  const tc::Scratch sc = tc::carve(ws, B);
  return dispatch(vol_src, vol_dtype, f16_gather, vol_tgt, nullptr, R, r_per_pair, W1, W2, b2, base, scores, true,
                  B, N, sc, tc::Finalize{best_val, best_idx, R_best, idx_offset, nullptr}, s);
}

}  // namespace ahv

#ifdef AHV_TIMELINE
extern "C" AHV_API int ahv_diag_timeline(unsigned long long* host_out /*[160][16]*/) {
  return cudaMemcpyFromSymbol(host_out, ahv::tc::g_timeline, sizeof(unsigned long long) * 160 * 16) == cudaSuccess ? AHV_OK : AHV_ECUDA;
}
#endif
