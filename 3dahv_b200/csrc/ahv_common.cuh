// Shared constants and small device helpers for lib3dahv_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ahv_b200.h"

namespace ahv {

constexpr int kC = 16;      // volume channels            modules/modules.py:64
constexpr int kS = 8;       // D = H = W                  modules/modules.py:97-100
constexpr int kVox = 512;   // voxel samples per hypothesis
constexpr int kK = 384;     // tri-plane channels          modules/modules.py:67
constexpr int kO = 32;      // head output channels        modules/model.py:34
constexpr int kP = 64;      // positions of the folded 8x8 plane
constexpr int kMaxK = 32;   // largest supported top-k

// Source volume staged in shared memory with a one-voxel zero halo, channel
// innermost: line L = ((z+1)*10 + (y+1))*10 + (x+1), 16 fp32 (64 B) per line.
// Out-of-range taps of padding_mode='zeros' (utils.py:129) then read zeros and
// need no per-tap bounds test.
constexpr int kHalo = 10;
constexpr int kLines = kHalo * kHalo * kHalo;  // 1000
constexpr int kVolSmemBytes = kLines * kC * 4; // 64000

#define AHV_CUDA_OK(expr)                          \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) return AHV_ECUDA;       \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sample position of output voxel (x,y,z base coordinates) under rotation R
// (F.affine_grid, utils.py:126), un-normalised as grid_sample does with
// align_corners=False: i = ((g+1)*8-1)/2.  ((g+1)*8 is exact, so one FFMA
// (g+1)*4-0.5 reproduces ATen's two-step rounding bit for bit.)
struct Tap {
  int line;          // halo line index of the (x0,y0,z0) corner
  float fx, fy, fz;  // fractional parts
};

__device__ __forceinline__ float unnorm(float g) { return fmaf(__fadd_rn(g, 1.0f), 4.0f, -0.5f); }

__device__ __forceinline__ Tap make_tap(const float* R, float x, float y, float z) {
  // grid = R @ (x, y, z); x -> W, y -> H, z -> D
  float gx = fmaf(R[2], z, fmaf(R[1], y, R[0] * x));
  float gy = fmaf(R[5], z, fmaf(R[4], y, R[3] * x));
  float gz = fmaf(R[8], z, fmaf(R[7], y, R[6] * x));
  // clamp to [-1, 8]: anything outside only ever touches the zero halo
  float ix = fminf(fmaxf(unnorm(gx), -1.0f), 8.0f);
  float iy = fminf(fmaxf(unnorm(gy), -1.0f), 8.0f);
  float iz = fminf(fmaxf(unnorm(gz), -1.0f), 8.0f);
  float x0 = fminf(floorf(ix), 7.0f), y0 = fminf(floorf(iy), 7.0f), z0 = fminf(floorf(iz), 7.0f);
  Tap t;
  t.fx = ix - x0;
  t.fy = iy - y0;
  t.fz = iz - z0;
  t.line = (((int)z0 + 1) * kHalo + ((int)y0 + 1)) * kHalo + ((int)x0 + 1);
  return t;
}

// (score, index) -> one ordered 64-bit key: larger key = higher score, ties -> lower index
typedef unsigned long long u64;

__device__ __forceinline__ u64 make_key(float score, uint32_t idx) {
  uint32_t u = __float_as_uint(score + 0.0f);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((u64)u << 32) | (u64)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float key_score(u64 key) {
  uint32_t u = (uint32_t)(key >> 32);
  u = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
  return __uint_as_float(u);
}
__device__ __forceinline__ uint32_t key_index(u64 key) { return 0xFFFFFFFFu - (uint32_t)key; }


namespace peer { struct Args; }  // ahv_peer.cuh

// Launch-side helpers implemented in the .cu files --------------------------
int launch_so3_from_normals(const float* normals, float* R, int64_t n, cudaStream_t s);
int launch_so3_sample(uint64_t seed, int64_t first, float* R, int64_t n, cudaStream_t s);
int launch_so3_grid(int64_t n_total, int64_t first, float* R, int64_t count, cudaStream_t s);
int launch_so3_perturb(const float* centers, int64_t n, int m, float max_angle_deg, uint64_t seed, float* out,
                       cudaStream_t s);
int launch_rotate_volume(const float* vol, int per_rot, const float* R, const float* base,
                         float* out, int64_t n, cudaStream_t s);
int launch_rotate_volume_bwd(const float* grad_out, int per_rot, const float* R, const float* base,
                             float* grad_vol, int64_t n, cudaStream_t s);
int launch_score_bwd(const float* vol_src, const float* tgt_feat, const float* R, int r_per_pair,
                     const float* W1, const float* W2, const float* b2, const float* base,
                     const float* grad_scores, float* g_vol, float* g_tgt, float* g_W1, float* g_W2,
                     float* g_b2, int B, int64_t N, cudaStream_t s, const void* h1_saved = nullptr,
                     const float* pair_inv = nullptr);
int launch_score_bwd_tc(const float* vol_src, const float* tgt_feat, const float* R, int r_per_pair,
                        const float* W1, const float* W2, const float* b2, const float* base,
                        const float* grad_scores, const void* h1_saved, const float* pair_inv, float* g_vol,
                        float* g_tgt, float* g_W1, float* g_W2, float* g_b2, int B, int64_t N, cudaStream_t s);
int launch_score_tc_train(const float* vol_src, const float* tgt_feat, const float* R, int r_per_pair, const float* W1,
                          const float* W2, const float* b2, const float* base, float* scores, void* h1_out,
                          float* pair_inv_out, int B, int64_t N, void* ws, size_t ws_bytes, cudaStream_t s);
int launch_resblock3d(const float* x, const float* Wc1, const float* Wc2, const float* Wd, float* out, int64_t m,
                      cudaStream_t s);
int launch_infonce(const float* scores, const float* R, int r_per_pair, const float* gt, float thr_deg, float temperature,
                   float* loss, float* grad, int B, int64_t N, cudaStream_t s);
int launch_forward_3d2d(const float* vol, const float* W1, const float* W2, const float* b2,
                        float* feat, int64_t m, cudaStream_t s);
int launch_tgt_feat(const float* vol_tgt, const float* W1, const float* W2, const float* b2,
                    unsigned long long* clear_keys, float* feat, int B, cudaStream_t s);
int launch_score_fp32(const void* vol_src, int vol_dtype, const float* tgt_feat, const float* R,
                      int r_per_pair, const float* W1, const float* W2, const float* b2,
                      const float* base, float* scores, int B, int64_t N, cudaStream_t s);
int launch_score_tc(const void* vol_src, int vol_dtype, const float* tgt_feat, const float* R,
                    int r_per_pair, const float* W1, const float* W2, const float* b2,
                    const float* base, float* scores, int B, int64_t N, void* ws, size_t ws_bytes,
                    cudaStream_t s, bool f16_gather);
size_t score_tc_workspace_bytes(int B, int64_t N);
float* scratch_tgt_feat(void* ws, int B);  // [B,32,64] slot inside the tensor-core scratch
int launch_verify_tc_argmax(const void* vol_src, int vol_dtype, const float* vol_tgt, const float* R,
                            int r_per_pair, const float* W1, const float* W2, const float* b2,
                            const float* base, float* scores, float* best_val, int64_t* best_idx,
                            float* R_best, int64_t idx_offset, int B, int64_t N, void* ws, size_t ws_bytes,
                            cudaStream_t s, bool f16_gather, const peer::Args* pa = nullptr);
int launch_topk_exchange(const float* val, const int64_t* idx, const float* R, int r_per_pair, int64_t idx_offset,
                         int64_t N, int B, int k, float* out_val, int64_t* out_idx, float* R_best,
                         const peer::Args& pa, cudaStream_t s);
int launch_topk(const float* scores, int B, int64_t N, int k, int64_t idx_offset, float* val,
                int64_t* idx, void* ws, size_t ws_bytes, cudaStream_t s);
size_t topk_workspace_bytes(int B, int64_t N, int k);
int launch_topk_merge(const float* vals, const int64_t* idx, int parts, int B, int k, float* out_val,
                      int64_t* out_idx, cudaStream_t s);
int launch_gather_rotations(const float* R, int r_per_pair, const int64_t* idx, int64_t idx_offset,
                            int B, int64_t N, int k, float* R_out, cudaStream_t s);

}  // namespace ahv
