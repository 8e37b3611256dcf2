/* A non-Python host of lib3dahv_b200: plain C, host buffers in, selection out (include/ahv_b200.h).
 *
 *   gcc -O2 -Iinclude examples/predict_host.c -o examples/predict_host -L3dahv_b200 -l:lib3dahv_b200.so \
 *       -Wl,-rpath,$PWD/3dahv_b200
 *   examples/predict_host inputs.bin outputs.bin
 *
 * inputs.bin  (little endian): int32 B, int32 N, int32 k, then float32 vol_src[B*8192], vol_tgt[B*8192], R[N*9],
 *             W1[32*384], W2[32*32], b2[32], base[8]
 * outputs.bin: int64 idx[B*k], float32 val[B*k], float32 R_best[B*k*9]
 * This is the call sequence a cgo / JNI / N-API binding of the reference's hot path (modules/model.py:184-196) makes:
 * one session, one ahv_predict_host_ex per batch of pairs. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ahv_b200.h"

static float* read_floats(FILE* f, size_t n) {
  float* p = (float*)malloc(n * sizeof(float));
  if (!p || fread(p, sizeof(float), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return p;
}

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s inputs.bin outputs.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  int32_t hdr[3];
  if (fread(hdr, sizeof(int32_t), 3, f) != 3) return 2;
  const int B = hdr[0], k = hdr[2];
  const int64_t N = hdr[1];
  float* vol_src = read_floats(f, (size_t)B * 8192);
  float* vol_tgt = read_floats(f, (size_t)B * 8192);
  float* R = read_floats(f, (size_t)N * 9);
  float* W1 = read_floats(f, 32 * 384);
  float* W2 = read_floats(f, 32 * 32);
  float* b2 = read_floats(f, 32);
  float* base = read_floats(f, 8);
  fclose(f);

  int64_t* idx = (int64_t*)malloc((size_t)B * k * sizeof(int64_t));
  float* val = (float*)malloc((size_t)B * k * sizeof(float));
  float* R_best = (float*)malloc((size_t)B * k * 9 * sizeof(float));
  ahv_host_session* hs = NULL;
  int st = ahv_host_session_create(&hs);
  if (st != AHV_OK) { fprintf(stderr, "session: %s\n", ahv_status_string(st)); return 1; }
  for (int rep = 0; rep < 2; ++rep) { /* the second call reuses the session's device scratch */
    st = ahv_predict_host_ex(hs, vol_src, AHV_VOL_F32, vol_tgt, R, /*r_per_pair=*/0, W1, W2, b2, base, /*scores=*/NULL, val,
                             idx, R_best, k, /*idx_offset=*/0, B, N, AHV_MATH_TC, /*rank=*/0, /*world=*/1, /*peers=*/NULL, 0,
                             0, /*stream=*/NULL);
    if (st != AHV_OK) { fprintf(stderr, "ahv_predict_host_ex: %s\n", ahv_status_string(st)); return 1; }
  }
  ahv_host_session_destroy(hs);

  f = fopen(argv[2], "wb");
  if (!f) { perror(argv[2]); return 2; }
  fwrite(idx, sizeof(int64_t), (size_t)B * k, f);
  fwrite(val, sizeof(float), (size_t)B * k, f);
  fwrite(R_best, sizeof(float), (size_t)B * k * 9, f);
  fclose(f);
  printf("lib3dahv_b200 %d: %d pairs x %lld hypotheses, top-%d; pair 0 best index %lld score %.6f\n", ahv_version(), B,
         (long long)N, k, (long long)idx[0], val[0]);
  return 0;
}
