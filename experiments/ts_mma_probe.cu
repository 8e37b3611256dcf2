// Probe: tcgen05.mma with the A operand in TMEM (TS form), M=64, two tiles interleaved in the two 16-lane
// halves of each sub-partition (lane offset 16), as the scoring kernel lays out its conv1 accumulators.
// Prints how many D entries match the expectation for each (slot) and dumps a few rows.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

__global__ void __launch_bounds__(128, 1) probe(float* out /*[128][32]*/, int a_lane_off_slot1) {
  __shared__ __align__(128) unsigned char bsm[1024];  // B: [chalf][ngroup][8][8] fp16, N=32, K=16
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 512; i += 128) {
    const int n = i / 16, k = i % 16;
    const float v = (float)((n * 5 + k) % 7 - 3);
    reinterpret_cast<__half*>(bsm)[(k >> 3) * 256 + (n >> 3) * 64 + (n & 7) * 8 + (k & 7)] = __float2half(v);
  }
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  // A in TMEM columns [0,8): thread (warp w, lane t) owns TMEM lane 32w+t: slot = t>>4, row m = 16w + (t&15)
  {
    const int slot = lane >> 4, m = 16 * warp + (lane & 15);
    uint32_t r[8];
    for (int c = 0; c < 8; ++c) {
      const float a0 = (float)((m + 3 * (2 * c) + 7 * slot) % 13 - 6), a1 = (float)((m + 3 * (2 * c + 1) + 7 * slot) % 13 - 6);
      __half2 h = __floats2half2_rn(a0, a1);
      r[c] = *reinterpret_cast<uint32_t*>(&h);
    }
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint64_t bdesc = smem_desc(smem_u32(bsm), 512, 128);
    const uint32_t idesc = instr_desc(64, 32);
    for (int slot = 0; slot < 2; ++slot) {
      const uint32_t d = tmem + ((uint32_t)(16 * slot) << 16) + 16;
      const uint32_t a = tmem + ((uint32_t)((slot ? a_lane_off_slot1 : 0)) << 16) + 0;
      asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                   ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t done;
    do {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    } while (!done);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(tmem + ((uint32_t)(32 * warp) << 16) + 16) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < 32; ++n) out[threadIdx.x * 32 + n] = __uint_as_float(r[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
  float* d; cudaMalloc(&d, 128 * 32 * 4);
  float h[128 * 32];
  for (int variant = 0; variant < 2; ++variant) {
    const int off = variant == 0 ? 16 : 0;
    cudaMemset(d, 0, sizeof(h));
    probe<<<1, 128>>>(d, off);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    // expectation: thread (w,t): slot=t>>4, m=16w+(t&15): D[m][n] = sum_k A_slot[m][k]*B[n][k]
    for (int a_slot_hyp = 0; a_slot_hyp < 2; ++a_slot_hyp) {   // which A data did slot 1 actually use?
      int ok[2] = {0, 0}, tot[2] = {0, 0};
      for (int th = 0; th < 128; ++th) {
        const int w = th >> 5, t = th & 31, slot = t >> 4, m = 16 * w + (t & 15);
        const int aslot = slot == 0 ? 0 : a_slot_hyp;
        for (int n = 0; n < 32; ++n) {
          float acc = 0;
          for (int k = 0; k < 16; ++k) acc += (float)((m + 3 * k + 7 * aslot) % 13 - 6) * (float)((n * 5 + k) % 7 - 3);
          tot[slot]++;
          if (h[th * 32 + n] == acc) ok[slot]++;
        }
      }
      printf("variant A-lane-offset(slot1)=%d, hypothesis 'slot1 reads A of slot %d': slot0 %d/%d  slot1 %d/%d\n", off,
             a_slot_hyp, ok[0], tot[0], ok[1], tot[1]);
    }
    printf("  row th=0: %g %g %g %g | th=16: %g %g %g %g | th=33: %g %g %g %g\n", h[0], h[1], h[2], h[3], h[16 * 32], h[16 * 32 + 1],
           h[16 * 32 + 2], h[16 * 32 + 3], h[33 * 32], h[33 * 32 + 1], h[33 * 32 + 2], h[33 * 32 + 3]);
  }
  return 0;
}
