// Micro-benchmark: cycles per tcgen05.mma (M=128, N, K=16, fp16) with both operands in shared memory (SS) versus
// A copied into TMEM by tcgen05.cp right before each MMA (TS).  One CTA per SM, one issuing thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu ; run: ./umma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) { return (sbo_bytes >> 4) | (1u << 14) | (layout_type << 29); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

__global__ void __launch_bounds__(320, 1) probe(int N, int iters, int mode, int taps, long long* out, int spin) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2, bar3[8];
  __shared__ uint32_t tmem_slot;
  // A region: 180 rows x 128 B (like a halo slice), B region: N rows x 128 B x taps
  uint8_t* sA = smem;
  uint8_t* sB = smem + 24 * 1024;
  for (int i = threadIdx.x; i < (24 * 1024 + 9 * 96 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(mode == 2 ? 2 : 1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(1));
    for (int i = 0; i < 8; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar3[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  long long t0 = 0, t1 = 0;
  const int nissue = mode == 2 ? 2 : 1;
  const int who = threadIdx.x >> 5;
  uint32_t elected = 0;
  if (who < nissue) asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
  if (who < nissue && elected) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi_a = desc_hi(10 * 128, 2u), hi_b = desc_hi(8 * 128, 2u);
    const uint32_t a_lo = desc_lo(smem_u32(sA)), b_lo = desc_lo(smem_u32(sB));
    const uint32_t tmem_d = tmem_base + who * 128;   // accumulator columns
    const uint32_t tmem_a = tmem_base + 384;      // A tiles: 8 columns each, 16 slots
    t0 = clock64();
    for (int it = who; it < iters; it += nissue) {
      for (int t9 = 0; t9 < taps; t9++) {
        const uint32_t alo = a_lo + (uint32_t)(((t9 / 3) * 10 + (t9 % 3)) * 8);
        const uint32_t blo = b_lo + (uint32_t)((N <= 96 ? t9 : 0) * ((N * 128) >> 4));
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint64_t da = desc64(hi_a, alo + 2 * k), db = desc64(hi_b, blo + 2 * k);
          if (mode == 3 || mode == 4) {
            const uint32_t td = tmem_base + (uint32_t)(((t9 * 4 + k) % (mode == 3 ? 2 : 4)) * 96);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(td), "l"(da), "l"(db), "r"(idesc) : "memory");
          } else if (mode != 1) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc) : "memory");
          } else {
            const uint32_t ta = tmem_a + (uint32_t)(((t9 * 4 + k) & 15) * 8);
            asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(ta), "l"(da) : "memory");
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem_d), "r"(ta), "l"(db), "r"(idesc) : "memory");
          }
        }
      }
      if (mode >= 5) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar3[it & 7])) : "memory");
        if (mode == 6) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar3[(it + 4) & 7])) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    if (who == 1) { t0 = 0; }
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    t1 = clock64();
    if (blockIdx.x == 0 && who == 0) out[0] = t1 - t0;
    if (who == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar2)) : "memory");
  }
  if (who >= 2 && spin >= 3) {
    // shared-memory traffic generators: spin 3: 16-byte loads only; 4: loads + stores; they stop when bar2 completes
    uint8_t* scratch = smem + 140 * 1024 + (who - 2) * 4096;
    uint4 acc = make_uint4(0, 0, 0, 0);
    uint32_t ok = 0;
    while (!ok) {
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const uint32_t sa = smem_u32(scratch + ((threadIdx.x & 31) * 16 + r * 512));
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa));
        acc.x ^= v.x; acc.y ^= v.y;
        if (spin == 4) asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(acc.x), "r"(acc.y), "r"(acc.z), "r"(acc.w) : "memory");
      }
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar2)), "r"(0) : "memory");
    }
    if (acc.x == 0x12345678u) out[1] = acc.y;
  } else if (who >= 2 && spin && (spin == 2 || (who & 3) == 0)) {   // spin == 1: only the warps on the issuer's sub-partition
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar2)), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 200, taps = 9;
  for (int spin : {0})
  for (int N : {48, 96}) {
    for (int mode : {0, 5, 6}) {
      probe<<<148, 320, 180 * 1024>>>(N > 96 ? 96 : N, 2, mode, taps, d, spin);   // warm-up
      probe<<<148, 320, 180 * 1024>>>(N, iters, mode, taps, d, spin);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("spin=%d N=%3d mode=%s: %s  %.1f cycles / MMA\n", spin, N, mode == 1 ? "TS (tcgen05.cp + A in TMEM)" : (mode == 2 ? "SS, two issuing warps" : (mode == 3 ? "SS, one issuer, 2 accumulators round-robin" : (mode == 4 ? "SS, one issuer, 4 accumulators round-robin" : (mode == 5 ? "SS + 1 commit per 36 MMAs" : (mode == 6 ? "SS + 2 commits per 36 MMAs" : "SS"))))), cudaGetErrorString(e),
             (double)cyc / (iters * taps * 4));
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
