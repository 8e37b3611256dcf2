/* CPU oracle (plain C) for the 3DAHV hypothesis-and-verification hot path.
 *
 * TEST INFRASTRUCTURE ONLY — never linked into lib3dahv_b200.so and never
 * called from the product path.  Users: tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py.  Parity status: PINNED against
 * tests/golden/ fixtures (outputs of the reference's own Python code, see
 * oracle/make_golden.py) by tests/test_oracle_golden.py.
 *
 * Restates, as explicit scalar arithmetic:
 *   utils.py:113-131              rotate_volume  (affine_grid + grid_sample,
 *                                 trilinear, padding zeros, align_corners=False)
 *   modules/modules.py:112-124    Feature_Aligner.forward_3d2d (tri-plane fold,
 *                                 conv1x1 384->32, ReLU, conv1x1 32->32 + bias,
 *                                 L2 normalise over channels)
 *   modules/model.py:193          (a * b[:,None]).sum(2).mean(-1)
 *   modules/model.py:195          torch.max(dim=1)  (first maximal index)
 *   pytorch3d random_rotations    normals -> unit quaternion -> matrix
 *                                 (third-party, algorithm restated)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AC 16
#define AS 8
#define AVOX 512
#define AK 384
#define AO 32
#define AP 64

/* pytorch3d random_quaternions + quaternion_to_matrix; every op rounded on its
 * own (compile with -ffp-contract=off), sequential 4-term sums. */
void ahv_oracle_so3_from_normals(const float *o, float *R, int64_t n) {
  for (int64_t t = 0; t < n; ++t) {
    const float a = o[4 * t], b = o[4 * t + 1], c = o[4 * t + 2], d = o[4 * t + 3];
    float s = ((a * a + b * b) + c * c) + d * d;
    float den = copysignf(sqrtf(s), a);
    float r = a / den, i = b / den, j = c / den, k = d / den;
    float two_s = 2.0f / (((r * r + i * i) + j * j) + k * k);
    float *m = R + 9 * t;
    m[0] = 1.0f - two_s * (j * j + k * k);
    m[1] = two_s * (i * j - k * r);
    m[2] = two_s * (i * k + j * r);
    m[3] = two_s * (i * j + k * r);
    m[4] = 1.0f - two_s * (i * i + k * k);
    m[5] = two_s * (j * k - i * r);
    m[6] = two_s * (i * k - j * r);
    m[7] = two_s * (j * k + i * r);
    m[8] = 1.0f - two_s * (i * i + j * j);
  }
}

/* Super-Fibonacci SO(3) grid (extension; Alexa 2022), fp64 then one rounding to fp32, then the
 * same quaternion -> matrix map.  sin/cos of 2*pi*frac(x): glibc's fp64 sin/cos and CUDA's
 * sincospi are both < 1 ulp in fp64, so the fp32 roundings agree. */
void ahv_oracle_so3_grid(int64_t n_total, int64_t first, float *R, int64_t count) {
  const double two_pi = 6.283185307179586476925286766559;
  for (int64_t t = 0; t < count; ++t) {
    const double s = (double)(first + t) + 0.5;
    const double u = s / (double)n_total;
    const double r = sqrt(u), rr = sqrt(1.0 - u);
    double ta = s * 0.70710678118654752440, tb = s * 0.65199624317913454;
    ta -= floor(ta);
    tb -= floor(tb);
    float q[4] = {(float)(r * sin(two_pi * ta)), (float)(r * cos(two_pi * ta)),
                  (float)(rr * sin(two_pi * tb)), (float)(rr * cos(two_pi * tb))};
    ahv_oracle_so3_from_normals(q, R + 9 * t, 1);
  }
}

/* utils.py:113-131: one hypothesis.  vol [16][512], out [16][512]. */
void ahv_oracle_rotate_one(const float *vol, const float *R, const float *base, float *out) {
  for (int d = 0; d < AS; ++d)
    for (int h = 0; h < AS; ++h)
      for (int w = 0; w < AS; ++w) {
        const float x = base[w], y = base[h], z = base[d];
        /* affine_grid: grid = theta @ (x, y, z, 1), theta = [R | 0] */
        const float gx = (R[0] * x + R[1] * y) + R[2] * z;
        const float gy = (R[3] * x + R[4] * y) + R[5] * z;
        const float gz = (R[6] * x + R[7] * y) + R[8] * z;
        /* grid_sampler unnormalize, align_corners=False: ((g+1)*size-1)/2 */
        const float ix = ((gx + 1.0f) * 8.0f - 1.0f) / 2.0f;
        const float iy = ((gy + 1.0f) * 8.0f - 1.0f) / 2.0f;
        const float iz = ((gz + 1.0f) * 8.0f - 1.0f) / 2.0f;
        const float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
        const float tx = ix - fx0, ty = iy - fy0, tz = iz - fz0;
        const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
        const int v = (d * AS + h) * AS + w;
        float acc[AC];
        for (int c = 0; c < AC; ++c) acc[c] = 0.0f;
        for (int dz = 0; dz < 2; ++dz)
          for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
              const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
              if (xx < 0 || xx >= AS || yy < 0 || yy >= AS || zz < 0 || zz >= AS) continue;
              const float wgt = (dx ? tx : 1.0f - tx) * (dy ? ty : 1.0f - ty) * (dz ? tz : 1.0f - tz);
              const int src = (zz * AS + yy) * AS + xx;
              for (int c = 0; c < AC; ++c) acc[c] += wgt * vol[c * AVOX + src];
            }
        for (int c = 0; c < AC; ++c) out[c * AVOX + v] = acc[c];
      }
}

/* modules/modules.py:112-124 on one volume [16][8][8][8] -> feat [32][64]. */
void ahv_oracle_forward_3d2d_one(const float *v, const float *W1, const float *W2, const float *b2,
                                 float *feat) {
  float tri[AK][AP];
  float h1[AO][AP], h2[AO][AP];
  for (int c = 0; c < AC; ++c)
    for (int k = 0; k < AS; ++k)
      for (int p = 0; p < AS; ++p)
        for (int q = 0; q < AS; ++q) {
          /* x: 'b c d h w -> b (c w) d h'   y: '(c h) d w'   z: '(c d) h w' */
          tri[c * 8 + k][p * 8 + q] = v[c * AVOX + (p * 8 + q) * 8 + k];
          tri[128 + c * 8 + k][p * 8 + q] = v[c * AVOX + (p * 8 + k) * 8 + q];
          tri[256 + c * 8 + k][p * 8 + q] = v[c * AVOX + (k * 8 + p) * 8 + q];
        }
  for (int o = 0; o < AO; ++o) {
    for (int p = 0; p < AP; ++p) h1[o][p] = 0.0f;
    for (int k = 0; k < AK; ++k) {
      const float wk = W1[o * AK + k];
      for (int p = 0; p < AP; ++p) h1[o][p] += wk * tri[k][p];
    }
    for (int p = 0; p < AP; ++p) h1[o][p] = h1[o][p] > 0.0f ? h1[o][p] : 0.0f;
  }
  for (int o = 0; o < AO; ++o) {
    for (int p = 0; p < AP; ++p) h2[o][p] = 0.0f;
    for (int i = 0; i < AO; ++i) {
      const float wk = W2[o * AO + i];
      for (int p = 0; p < AP; ++p) h2[o][p] += wk * h1[i][p];
    }
    for (int p = 0; p < AP; ++p) h2[o][p] += b2[o];
  }
  for (int p = 0; p < AP; ++p) {
    float s = 0.0f;
    for (int o = 0; o < AO; ++o) s += h2[o][p] * h2[o][p];
    float nrm = sqrtf(s);
    if (nrm < 1e-12f) nrm = 1e-12f; /* F.normalize eps */
    for (int o = 0; o < AO; ++o) feat[o * AP + p] = h2[o][p] / nrm;
  }
}

/* ---- training variant: gradient of the scores (modules/model.py:53-56 under autograd, inside infoNCE_loss :43-63) ----
 * Explicit chain rule in double precision, one (pair, hypothesis) at a time, deliberately in the textbook
 * SCATTER form for the resampling adjoint (the CUDA kernel uses a gather):
 *   X = rotate(V_b, R_n) (utils.py:113-131), A = tri-plane(X), H1 = relu(A W1^T), H2 = H1 W2^T + b2,
 *   F = H2 / max(|H2|, eps) (modules/modules.py:112-124), s = mean_p <F_p, T_p>  (modules/model.py:55)
 * grad_scores [B][N] -> g_vol [B][16][512], g_tgt [B][32][64], g_W1 [32][384], g_W2 [32][32], g_b2 [32]
 * (all double, ADDED into).  tgt = forward_3d2d(vol_tgt) is an input: its own backward is not part of the
 * hot path.  Single-threaded: small cases only. */
void ahv_oracle_score_backward(const float *vol_src, const float *tgt, const float *R, int r_per_pair,
                               const float *W1, const float *W2, const float *b2, const float *base, int B,
                               int64_t N, const float *grad_scores, double *g_vol, double *g_tgt, double *g_W1,
                               double *g_W2, double *g_b2) {
  static double X[AC][AVOX], dX[AC][AVOX], tri[AK][AP], dtri[AK][AP], h1[AO][AP], h2[AO][AP], dh1[AO][AP], dh2[AO][AP];
  for (int b = 0; b < B; ++b)
    for (int64_t n = 0; n < N; ++n) {
      const float *Rn = R + (r_per_pair ? ((int64_t)b * N + n) : n) * 9;
      const float *V = vol_src + (int64_t)b * AC * AVOX;
      const float *T = tgt + (int64_t)b * AO * AP;
      const double g = (double)grad_scores[(int64_t)b * N + n] / (double)AP;
      /* forward, recording each output voxel's taps (same float arithmetic as ahv_oracle_rotate_one) */
      static int tap_src[AVOX][8];
      static double tap_w[AVOX][8];
      for (int d = 0; d < AS; ++d)
        for (int h = 0; h < AS; ++h)
          for (int w = 0; w < AS; ++w) {
            const float x = base[w], y = base[h], z = base[d];
            const float gx = (Rn[0] * x + Rn[1] * y) + Rn[2] * z;
            const float gy = (Rn[3] * x + Rn[4] * y) + Rn[5] * z;
            const float gz = (Rn[6] * x + Rn[7] * y) + Rn[8] * z;
            const float ix = ((gx + 1.0f) * 8.0f - 1.0f) / 2.0f;
            const float iy = ((gy + 1.0f) * 8.0f - 1.0f) / 2.0f;
            const float iz = ((gz + 1.0f) * 8.0f - 1.0f) / 2.0f;
            const float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
            const float tx = ix - fx0, ty = iy - fy0, tz = iz - fz0;
            const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
            const int v = (d * AS + h) * AS + w;
            for (int c = 0; c < AC; ++c) X[c][v] = 0.0;
            for (int t = 0; t < 8; ++t) {
              const int dx = t & 1, dy = (t >> 1) & 1, dz = t >> 2;
              const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
              tap_src[v][t] = -1;
              tap_w[v][t] = 0.0;
              if (xx < 0 || xx >= AS || yy < 0 || yy >= AS || zz < 0 || zz >= AS) continue;
              const float wgt = (dx ? tx : 1.0f - tx) * (dy ? ty : 1.0f - ty) * (dz ? tz : 1.0f - tz);
              tap_src[v][t] = (zz * AS + yy) * AS + xx;
              tap_w[v][t] = (double)wgt;
              for (int c = 0; c < AC; ++c) X[c][v] += (double)wgt * (double)V[c * AVOX + tap_src[v][t]];
            }
          }
      for (int c = 0; c < AC; ++c)
        for (int k = 0; k < AS; ++k)
          for (int p = 0; p < AS; ++p)
            for (int q = 0; q < AS; ++q) {
              tri[c * 8 + k][p * 8 + q] = X[c][(p * 8 + q) * 8 + k];
              tri[128 + c * 8 + k][p * 8 + q] = X[c][(p * 8 + k) * 8 + q];
              tri[256 + c * 8 + k][p * 8 + q] = X[c][(k * 8 + p) * 8 + q];
            }
      for (int o = 0; o < AO; ++o)
        for (int p = 0; p < AP; ++p) {
          double a = 0.0;
          for (int k = 0; k < AK; ++k) a += (double)W1[o * AK + k] * tri[k][p];
          h1[o][p] = a > 0.0 ? a : 0.0;
        }
      for (int o = 0; o < AO; ++o)
        for (int p = 0; p < AP; ++p) {
          double a = (double)b2[o];
          for (int i = 0; i < AO; ++i) a += (double)W2[o * AO + i] * h1[i][p];
          h2[o][p] = a;
        }
      /* backward */
      for (int p = 0; p < AP; ++p) {
        double ss = 0.0, ft = 0.0;
        for (int o = 0; o < AO; ++o) ss += h2[o][p] * h2[o][p];
        const double nraw = sqrt(ss), nr = nraw < 1e-12 ? 1e-12 : nraw;
        for (int o = 0; o < AO; ++o) ft += h2[o][p] / nr * (double)T[o * AP + p];
        for (int o = 0; o < AO; ++o) {
          const double F = h2[o][p] / nr, t = (double)T[o * AP + p];
          dh2[o][p] = g / nr * (nraw < 1e-12 ? t : t - F * ft);
          g_tgt[((int64_t)b * AO + o) * AP + p] += g * F;
          g_b2[o] += dh2[o][p];
        }
      }
      for (int o = 0; o < AO; ++o)
        for (int i = 0; i < AO; ++i) {
          double a = 0.0;
          for (int p = 0; p < AP; ++p) a += dh2[o][p] * h1[i][p];
          g_W2[o * AO + i] += a;
        }
      for (int i = 0; i < AO; ++i)
        for (int p = 0; p < AP; ++p) {
          double a = 0.0;
          for (int o = 0; o < AO; ++o) a += dh2[o][p] * (double)W2[o * AO + i];
          dh1[i][p] = h1[i][p] > 0.0 ? a : 0.0;
        }
      for (int o = 0; o < AO; ++o)
        for (int k = 0; k < AK; ++k) {
          double a = 0.0;
          for (int p = 0; p < AP; ++p) a += dh1[o][p] * tri[k][p];
          g_W1[o * AK + k] += a;
        }
      for (int k = 0; k < AK; ++k)
        for (int p = 0; p < AP; ++p) {
          double a = 0.0;
          for (int o = 0; o < AO; ++o) a += dh1[o][p] * (double)W1[o * AK + k];
          dtri[k][p] = a;
        }
      for (int c = 0; c < AC; ++c)
        for (int v = 0; v < AVOX; ++v) dX[c][v] = 0.0;
      for (int c = 0; c < AC; ++c)
        for (int k = 0; k < AS; ++k)
          for (int p = 0; p < AS; ++p)
            for (int q = 0; q < AS; ++q) {
              dX[c][(p * 8 + q) * 8 + k] += dtri[c * 8 + k][p * 8 + q];
              dX[c][(p * 8 + k) * 8 + q] += dtri[128 + c * 8 + k][p * 8 + q];
              dX[c][(k * 8 + p) * 8 + q] += dtri[256 + c * 8 + k][p * 8 + q];
            }
      for (int v = 0; v < AVOX; ++v)
        for (int t = 0; t < 8; ++t)
          if (tap_src[v][t] >= 0)
            for (int c = 0; c < AC; ++c) g_vol[((int64_t)b * AC + c) * AVOX + tap_src[v][t]] += tap_w[v][t] * dX[c][v];
    }
}

/* modules/model.py:186-193.  vol_src/vol_tgt [B][16][512]; R [N][9] shared
 * (r_per_pair=0) or [B][N][9]; scores [B][N].  The B*N independent
 * (pair, hypothesis) items are split evenly over `nthreads` pthreads. */
typedef struct {
  const float *vol_src, *tgt, *R, *W1, *W2, *b2, *base;
  int r_per_pair;
  int64_t N, lo, hi;
  float *scores;
} ahv_job;

static void *ahv_worker(void *arg) {
  const ahv_job *j = (const ahv_job *)arg;
  float *rot = (float *)malloc(AC * AVOX * sizeof(float));
  float *feat = (float *)malloc(AO * AP * sizeof(float));
  for (int64_t it = j->lo; it < j->hi; ++it) {
    const int64_t b = it / j->N, n = it % j->N;
    const float *Rn = j->R + (j->r_per_pair ? (size_t)it * 9 : (size_t)n * 9);
    ahv_oracle_rotate_one(j->vol_src + (size_t)b * AC * AVOX, Rn, j->base, rot);
    ahv_oracle_forward_3d2d_one(rot, j->W1, j->W2, j->b2, feat);
    const float *t = j->tgt + (size_t)b * AO * AP;
    float tot = 0.0f;
    for (int p = 0; p < AP; ++p) {
      float s = 0.0f; /* .sum(dim=2) over channels first */
      for (int o = 0; o < AO; ++o) s += feat[o * AP + p] * t[o * AP + p];
      tot += s;
    }
    j->scores[it] = tot / (float)AP; /* .mean(dim=-1) */
  }
  free(rot);
  free(feat);
  return NULL;
}

void ahv_oracle_score(const float *vol_src, const float *vol_tgt, const float *R, int r_per_pair,
                      const float *W1, const float *W2, const float *b2, const float *base, int B,
                      int64_t N, float *scores, int nthreads) {
  float *tgt = (float *)malloc((size_t)B * AO * AP * sizeof(float));
  for (int b = 0; b < B; ++b)
    ahv_oracle_forward_3d2d_one(vol_tgt + (size_t)b * AC * AVOX, W1, W2, b2, tgt + (size_t)b * AO * AP);
  const int64_t total = (int64_t)B * N;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if ((int64_t)nthreads > total) nthreads = total > 0 ? (int)total : 1;
  pthread_t th[256];
  ahv_job jobs[256];
  for (int t = 0; t < nthreads; ++t) {
    ahv_job jb = {vol_src, tgt, R, W1, W2, b2, base, r_per_pair, N,
                  total * t / nthreads, total * (t + 1) / nthreads, scores};
    jobs[t] = jb;
    if (t > 0) pthread_create(&th[t], NULL, ahv_worker, &jobs[t]);
  }
  ahv_worker(&jobs[0]);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(tgt);
}

/* modules/model.py:195: first maximal index per pair. */
void ahv_oracle_argmax(const float *scores, int B, int64_t N, float *best, int64_t *idx) {
  for (int b = 0; b < B; ++b) {
    int64_t bi = 0;
    float bv = scores[(size_t)b * N];
    for (int64_t n = 1; n < N; ++n)
      if (scores[(size_t)b * N + n] > bv) { bv = scores[(size_t)b * N + n]; bi = n; }
    best[b] = bv;
    idx[b] = bi;
  }
}
