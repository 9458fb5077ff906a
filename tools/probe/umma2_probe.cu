// Micro-benchmark + semantics check of tcgen05.mma.cta_group::2 (CTA pair, M = 256): cycles per MMA for the slab
// kernels' operand shapes (A = 128 rows per CTA, halo-strided 8-row groups, 128B swizzle; B = N/2 rows per CTA), and
// a numerical check of the operand split (CTA r supplies A rows 128r.. and B columns r*N/2..; each CTA's TMEM receives
// its 128 rows x N columns).  One pair per TPC, one issuing thread in the leader CTA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma2_probe umma2_probe.cu ; run: ./umma2_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) { return (sbo_bytes >> 4) | (1u << 14) | (layout_type << 29); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool wait_or_trap(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!try_wait(bar, parity))
    if (clock64() - t0 > 4000000000ll) { printf("umma2_probe: barrier watchdog\n"); __trap(); }
  return true;
}

__host__ __device__ inline float a_val(int rank, int row, int k) { return (float)((row * 7 + k * 3 + rank * 5) % 13 - 6); }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 5 + k * 11) % 9 - 4); }

constexpr int kARows = 180;

// group: 1 or 2 (cta_group); in group 1 both CTAs of the cluster run independent M = 128 MMAs on N columns
template <int GROUP, int N, int ACCS, int CHECK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
probe(int iters, long long* out, float* dout) {
  constexpr int check = CHECK, taps = CHECK ? 1 : 9, accs = ACCS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int nb = GROUP == 2 ? N / 2 : N;            // B rows held by this CTA
  uint8_t* sA = smem;                               // 180 rows x 128 B
  uint8_t* sB = smem + 24 * 1024;                   // taps x nb rows x 128 B
  for (int i = threadIdx.x; i < (24 * 1024 + 9 * 96 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  if (check) {
    // K-major, 128B swizzle: 16-byte chunk c of row r sits at chunk position c ^ (r & 7)
    for (int i = threadIdx.x; i < kARows * 64; i += blockDim.x) {
      const int r = i / 64, k = i % 64;
      reinterpret_cast<__half*>(sA)[r * 64 + (((k >> 3) ^ (r & 7)) << 3) + (k & 7)] = __float2half(a_val(rank, r, k));
    }
    for (int i = threadIdx.x; i < nb * 64; i += blockDim.x) {
      const int r = i / 64, k = i % 64;
      const int ng = GROUP == 2 ? (int)rank * nb + r : r;
      reinterpret_cast<__half*>(sB)[r * 64 + (((k >> 3) ^ (r & 7)) << 3) + (k & 7)] = __float2half(b_val(ng, k));
    }
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if (GROUP == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tmem_slot != 0u) { printf("unexpected TMEM base\n"); __trap(); }
  constexpr uint32_t tmem_base = 0u;
  const int warp = threadIdx.x >> 5;
  uint32_t elected = 0;
  if (warp == 1) asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
  if (warp == 1 && elected && (GROUP == 1 || rank == 0)) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((GROUP == 2 ? 256 : 128) >> 4) << 24);
    const uint32_t hi_a = desc_hi(10 * 128, 2u), hi_b = desc_hi(8 * 128, 2u);
    const uint32_t a_lo = desc_lo(smem_u32(sA)), b_lo = desc_lo(smem_u32(sB));
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int t9 = 0; t9 < taps; t9++) {
        const uint32_t alo = a_lo + (uint32_t)(((t9 / 3) * 10 + (t9 % 3)) * 8);
        const uint32_t blo = b_lo + (uint32_t)((nb <= 96 ? t9 : 0) * ((nb * 128) >> 4));
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint64_t da = desc64(hi_a, alo + 2 * k), db = desc64(hi_b, blo + 2 * k);
          const uint32_t td = tmem_base + (uint32_t)(((t9 * 4 + k) % accs) * N);
          const uint32_t acc = check ? (uint32_t)(k > 0) : 1u;
          if (GROUP == 2)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(td), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(td), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
      }
    }
    if (GROUP == 2)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    wait_or_trap(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  if (warp >= 2 && check) {
    wait_or_trap(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quad = warp & 3, lane = threadIdx.x & 31, m = quad * 32 + lane;
    if (blockIdx.x < 2)
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; j++) dout[((size_t)rank * 128 + m) * N + c0 + j] = __uint_as_float(r[j]);
      }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) {
    if (GROUP == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int GROUP, int N, int ACCS, int CHECK>
int run(int iters, long long* d, float* dout, long long* cyc) {
  cudaFuncSetAttribute(probe<GROUP, N, ACCS, CHECK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<GROUP, N, ACCS, CHECK><<<148, 192, 180 * 1024>>>(iters, d, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("group %d N %d: %s\n", GROUP, N, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(cyc, d, 8, cudaMemcpyDeviceToHost);
  return 0;
}

template <int N>
void timing(int iters, int taps, long long* d, float* dout) {
  long long c1 = 0, c2 = 0, c3 = 0, c4 = 0;
  run<1, N, 1, 0>(2, d, dout, &c1); run<1, N, 1, 0>(iters, d, dout, &c1);
  run<2, N, 1, 0>(2, d, dout, &c2); run<2, N, 1, 0>(iters, d, dout, &c2);
  run<1, N, 2, 0>(2, d, dout, &c3); run<1, N, 2, 0>(iters, d, dout, &c3);
  run<2, N, 2, 0>(2, d, dout, &c4); run<2, N, 2, 0>(iters, d, dout, &c4);
  const double q = (double)iters * taps * 4;
  printf("N=%3d: cta_group::1 (M=128) %.1f cycles/MMA (2 accumulators %.1f), cta_group::2 (M=256) %.1f (2 accumulators %.1f), nominal %.0f\n",
         N, c1 / q, c3 / q, c2 / q, c4 / q, 0.5 * N);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  float* dout; cudaMalloc(&dout, 2 * 128 * 256 * sizeof(float));
  long long cyc = 0;
  // ---- semantics: one tap, K = 64, N = 96
  for (int group = 1; group <= 2; group++) {
    const int N = 96;
    cudaMemset(dout, 0, 2 * 128 * 256 * sizeof(float));
    if (group == 1 ? run<1, 96, 1, 1>(1, d, dout, &cyc) : run<2, 96, 1, 1>(1, d, dout, &cyc)) return 1;
    float* h = (float*)malloc(2 * 128 * N * sizeof(float));
    cudaMemcpy(h, dout, 2 * 128 * N * sizeof(float), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rank = 0; rank < 2; rank++)
      for (int m = 0; m < 128; m++)
        for (int n = 0; n < N; n++) {
          float ref = 0.f;
          const int row = (m / 8) * 10 + m % 8;
          for (int k = 0; k < 64; k++) ref += a_val(rank, row, k) * b_val(n, k);
          const float got = h[((size_t)rank * 128 + m) * N + n];
          if (got != ref && bad++ < 5) printf("  group %d rank %d m %d n %d: got %g want %g\n", group, rank, m, n, got, ref);
        }
    printf("cta_group::%d semantics check (M = %d, N = 96, K = 64): %s (%d mismatches)\n", group, group * 128, bad ? "FAIL" : "ok", bad);
    free(h);
  }
  // ---- timing
  const int iters = 200, taps = 9;
  timing<32>(iters, taps, d, dout); timing<48>(iters, taps, d, dout); timing<64>(iters, taps, d, dout);
  timing<96>(iters, taps, d, dout); timing<128>(iters, taps, d, dout); timing<192>(iters, taps, d, dout);
  timing<256>(iters, taps, d, dout);
  return 0;
}
