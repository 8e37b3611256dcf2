// SO(3) hypothesis generation.
//
// ahv_so3_from_normals: the arithmetic of pytorch3d.transforms.random_rotations
// (call sites modules/model.py:102,131,184; model_co3d.py:86; test_co3d.py:106)
// applied to caller-supplied Gaussian draws, every operation an individually
// rounded IEEE fp32 op (no FMA contraction) so the result is bit-identical to
// oracle/ahv_oracle.c:ahv_oracle_so3_from_normals.
//
// ahv_so3_sample: native counter-based sampler (extension).
#include "ahv_common.cuh"

namespace ahv {

__device__ __forceinline__ void quat_to_matrix_exact(float a, float b, float c, float d,
                                                     float* __restrict__ m) {
  // s = ((a*a + b*b) + c*c) + d*d, sequential like torch's sum over 4 elements
  float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)),
                      __fmul_rn(d, d));
  float den = copysignf(__fsqrt_rn(s), a);
  float r = __fdiv_rn(a, den), i = __fdiv_rn(b, den), j = __fdiv_rn(c, den), k = __fdiv_rn(d, den);
  float qq = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r, r), __fmul_rn(i, i)), __fmul_rn(j, j)),
                       __fmul_rn(k, k));
  float two_s = __fdiv_rn(2.0f, qq);
  float ii = __fmul_rn(i, i), jj = __fmul_rn(j, j), kk = __fmul_rn(k, k);
  float ij = __fmul_rn(i, j), ik = __fmul_rn(i, k), jk = __fmul_rn(j, k);
  float ir = __fmul_rn(i, r), jr = __fmul_rn(j, r), kr = __fmul_rn(k, r);
  m[0] = __fsub_rn(1.0f, __fmul_rn(two_s, __fadd_rn(jj, kk)));
  m[1] = __fmul_rn(two_s, __fsub_rn(ij, kr));
  m[2] = __fmul_rn(two_s, __fadd_rn(ik, jr));
  m[3] = __fmul_rn(two_s, __fadd_rn(ij, kr));
  m[4] = __fsub_rn(1.0f, __fmul_rn(two_s, __fadd_rn(ii, kk)));
  m[5] = __fmul_rn(two_s, __fsub_rn(jk, ir));
  m[6] = __fmul_rn(two_s, __fsub_rn(ik, jr));
  m[7] = __fmul_rn(two_s, __fadd_rn(jk, ir));
  m[8] = __fsub_rn(1.0f, __fmul_rn(two_s, __fadd_rn(ii, jj)));
}

__global__ void __launch_bounds__(256) so3_from_normals_kernel(const float4* __restrict__ normals,
                                                               float* __restrict__ R, int64_t n) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float4 o = normals[t];
  float m[9];
  quat_to_matrix_exact(o.x, o.y, o.z, o.w, m);
#pragma unroll
  for (int e = 0; e < 9; ++e) R[t * 9 + e] = m[e];
}

// Philox-4x32-10 (Salmon et al. 2011), keyed by seed, counter = hypothesis index.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__global__ void __launch_bounds__(256) so3_sample_kernel(uint64_t seed, int64_t first,
                                                         float* __restrict__ R, int64_t n) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint64_t g = (uint64_t)(first + t);
  uint4 u = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), 0x3d41u, 0x4856u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  // two Box-Muller pairs -> four standard normals; (u+0.5)/2^32 lies in (0,1)
  const float inv = 2.3283064365386963e-10f;
  float u0 = ((float)(u.x >> 8) + 0.5f) * (1.0f / 16777216.0f), u1 = (float)u.y * inv;
  float u2 = ((float)(u.z >> 8) + 0.5f) * (1.0f / 16777216.0f), u3 = (float)u.w * inv;
  float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  float m[9];
  quat_to_matrix_exact(r0 * c0, r0 * s0, r1 * c1, r1 * s1, m);
#pragma unroll
  for (int e = 0; e < 9; ++e) R[t * 9 + e] = m[e];
}

// Deterministic near-uniform SO(3) grid: super-Fibonacci spiral (Alexa, CVPR 2022).  Point i of an
// n-point set: s=i+1/2, t=s/n, r=sqrt(t), rr=sqrt(1-t), a=2*pi*s/sqrt(2), b=2*pi*s/psi, psi=1.5337511687...
// q=(r sin a, r cos a, rr sin b, rr cos b) -> matrix by the same map as random_rotations.  Evaluated in fp64
// and rounded once to fp32 so the CPU restatement (oracle/ahv_oracle.c) reproduces it bit for bit.
__global__ void __launch_bounds__(256) so3_grid_kernel(int64_t n_total, int64_t first, float* __restrict__ R,
                                                       int64_t count) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const double s = (double)(first + t) + 0.5;
  const double u = s / (double)n_total;
  const double r = sqrt(u), rr = sqrt(1.0 - u);
  // reduce the angle to [0,1) turns before sinpi/cospi: exact in fp64 for these magnitudes
  double ta = s * 0.70710678118654752440, tb = s * 0.65199624317913454;  // s/sqrt(2), s/psi
  ta -= floor(ta);
  tb -= floor(tb);
  double sa, ca, sb, cb;
  sincospi(2.0 * ta, &sa, &ca);
  sincospi(2.0 * tb, &sb, &cb);
  float m[9];
  quat_to_matrix_exact((float)(r * sa), (float)(r * ca), (float)(rr * sb), (float)(rr * cb), m);
#pragma unroll
  for (int e = 0; e < 9; ++e) R[t * 9 + e] = m[e];
}

// Local refinement set around given rotations (extension, BASELINE config 4 "top-k refinement pass"; the
// reference has no refinement): out[i,0] = center[i]; out[i,j>0] = dR(i,j) @ center[i] where dR is a Haar
// rotation (Philox counter i*m+j, same normal -> quaternion map as the sampler) whose rotation ANGLE theta in
// [0,pi] is rescaled to theta/pi * max_angle about the same axis.  Restated in oracle/ahv_oracle.c.
__global__ void __launch_bounds__(256) so3_perturb_kernel(const float* __restrict__ centers, int64_t n, int m,
                                                          float max_angle_rad, uint64_t seed, float* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * m) return;
  const int64_t i = t / m;
  const int j = (int)(t - i * m);
  float c[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) c[e] = __ldg(centers + i * 9 + e);
  float* o = out + t * 9;
  if (j == 0) {
#pragma unroll
    for (int e = 0; e < 9; ++e) o[e] = c[e];
    return;
  }
  const uint4 u = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)((uint64_t)t >> 32), 0x7065u, 0x7274u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float inv32 = 2.3283064365386963e-10f;
  const float u0 = ((float)(u.x >> 8) + 0.5f) * (1.0f / 16777216.0f), u1 = (float)u.y * inv32;
  const float u2 = ((float)(u.z >> 8) + 0.5f) * (1.0f / 16777216.0f), u3 = (float)u.w * inv32;
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  const float qa = r0 * c0, qb = r0 * s0, qc = r1 * c1, qd = r1 * s1;  // Haar quaternion, unnormalised
  const float vn = sqrtf(qb * qb + qc * qc + qd * qd);
  const float theta = 2.0f * atan2f(vn, fabsf(qa));                    // rotation angle in [0, pi]
  const float phi = theta * (max_angle_rad * 0.318309886183790672f);   // rescaled into the cone
  float sh, ch;
  sincosf(0.5f * phi, &sh, &ch);
  const float k = vn > 0.0f ? sh / vn : 0.0f;
  float d[9];
  quat_to_matrix_exact(ch, k * qb, k * qc, k * qd, d);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q)
      o[r * 3 + q] = fmaf(d[r * 3 + 2], c[6 + q], fmaf(d[r * 3 + 1], c[3 + q], d[r * 3] * c[q]));
}

int launch_so3_perturb(const float* centers, int64_t n, int m, float max_angle_deg, uint64_t seed, float* out,
                       cudaStream_t s) {
  if (n * m == 0) return AHV_OK;
  so3_perturb_kernel<<<(unsigned)((n * m + 255) / 256), 256, 0, s>>>(centers, n, m, max_angle_deg * 0.017453292519943295f,
                                                                    seed, out);
  AHV_CUDA_OK(cudaGetLastError());
  return AHV_OK;
}

int launch_so3_grid(int64_t n_total, int64_t first, float* R, int64_t count, cudaStream_t s) {
  if (count == 0) return AHV_OK;
  so3_grid_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(n_total, first, R, count);
  AHV_CUDA_OK(cudaGetLastError());
  return AHV_OK;
}

int launch_so3_from_normals(const float* normals, float* R, int64_t n, cudaStream_t s) {
  if (n == 0) return AHV_OK;
  int64_t blocks = (n + 255) / 256;
  so3_from_normals_kernel<<<(unsigned)blocks, 256, 0, s>>>((const float4*)normals, R, n);
  AHV_CUDA_OK(cudaGetLastError());
  return AHV_OK;
}

int launch_so3_sample(uint64_t seed, int64_t first, float* R, int64_t n, cudaStream_t s) {
  if (n == 0) return AHV_OK;
  int64_t blocks = (n + 255) / 256;
  so3_sample_kernel<<<(unsigned)blocks, 256, 0, s>>>(seed, first, R, n);
  AHV_CUDA_OK(cudaGetLastError());
  return AHV_OK;
}

}  // namespace ahv
