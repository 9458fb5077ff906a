// Weight gradient of the stride-1 3x3x3 "same" convolutions (config #4 backward) - the round-1 version on
// mma.sync (m16n8k16, fp16 operands, fp32 accumulate), 84 TFLOP/s.  Superseded by conv_wgrad_tc.cu (tcgen05, 800-1200
// TFLOP/s) wherever that kernel covers the shape; kept as the fallback for W = 48 and as a cross-check in the tests.
//
//   dW[co][ci][kd][kh][kw] = sum over voxels v = (n, d, h, w) of  dY[v][co] * X[n, d+kd-1, h+kh-1, w+kw-1][ci]
//
// Per tap this is a GEMM with M = Cout, N = Cin and K = all voxels.  A CTA owns one 32 (co) x 32 (ci) block of all 27 taps
// and walks a strided set of w-rows (n, d, h): per row it stages dY[row] ([W][32] channels-last) and the 9 halo rows of X
// ([W + 2][32], zero outside the tensor) in shared memory; both operands are K-major in memory (voxel-major, channels
// contiguous), so both fragments come from ldmatrix.trans; the kw shift of a tap is a row offset into the X halo row.
// Warp w accumulates taps w, w + 8, w + 16 (and 24 + w < 27) in registers (<= 4 x 32 x 32 fp32 per warp).
// Split-K: every CTA writes its partial block to the workspace; a second kernel sums the partials in a fixed order
// (bit-reproducible) into the PyTorch weight layout (Cout, Cin, 3, 3, 3) fp32.
#include "common.cuh"
#include "../../include/nm_b200.h"
#include <stdlib.h>

namespace {

constexpr int kWgThreads = 256;
constexpr int kRowHalfs = 40;                 // 32 channels + 8 pad: 80-byte rows -> conflict-free ldmatrix
constexpr int kMaxW = 64;

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_k3_kernel(const __half* __restrict__ x, const __half* __restrict__ gy, int N, int D, int H, int W, int Cin, int Cout,
                     float* __restrict__ partial) {
  extern __shared__ __align__(16) uint8_t smem[];
  __half* sY = reinterpret_cast<__half*>(smem);                         // [W][kRowHalfs]
  __half* sX = sY + kMaxW * kRowHalfs;                                  // [9][W + 2][kRowHalfs]
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.z * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = N * D * H;
  const int xrow = (W + 2) * kRowHalfs;                                 // halfs per X halo row

  float acc[4][2][4][4];                                                // [tap slot][m tile][n tile][c0..c3]
#pragma unroll
  for (int s = 0; s < 4; s++)
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
      for (int n = 0; n < 4; n++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[s][m][n][e] = 0.f;

  const int lj = lane >> 3, li = lane & 7;                              // ldmatrix: lane -> (matrix lj, row li)
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int h = row % H, d = (row / H) % D, n = row / (H * D);
    __syncthreads();                                                    // previous row's fragments are consumed
    // ---- stage dY[row]: W voxels x 4 chunks of 8 channels
    for (int i = threadIdx.x; i < W * 4; i += kWgThreads) {
      const int w = i >> 2, c = i & 3;
      const __half* src = gy + ((long long)row * W + w) * Cout + co0 + c * 8;
      const uint4 v = *reinterpret_cast<const uint4*>(src);
      *reinterpret_cast<uint4*>(sY + w * kRowHalfs + c * 8) = v;
    }
    // ---- stage the 9 halo rows of X: (W + 2) voxels x 4 chunks each, zero outside the tensor
    for (int i = threadIdx.x; i < 9 * (W + 2) * 4; i += kWgThreads) {
      const int c = i & 3, t = i >> 2;
      const int wj = t % (W + 2), r9 = t / (W + 2);
      const int dd = d + r9 / 3 - 1, hh = h + r9 % 3 - 1, ww = wj - 1;
      const bool inside = (unsigned)dd < (unsigned)D && (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (inside) v = *reinterpret_cast<const uint4*>(x + ((((long long)n * D + dd) * H + hh) * W + ww) * Cin + ci0 + c * 8);
      *reinterpret_cast<uint4*>(sX + r9 * xrow + wj * kRowHalfs + c * 8) = v;
    }
    __syncthreads();
    // ---- MMAs: K = the W voxels of the row, 16 per step
    for (int k0 = 0; k0 < W; k0 += 16) {
      uint32_t a[2][4];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        // A = dY^T (co x voxel): stored block [voxel][co]; matrices (m0-7,k0-7), (m8-15,k0-7), (m0-7,k8-15), (m8-15,k8-15)
        const __half* p = sY + (k0 + (lj >> 1) * 8 + li) * kRowHalfs + m * 16 + (lj & 1) * 8;
        ldmatrix_x4_trans((uint32_t)__cvta_generic_to_shared(p), a[m][0], a[m][1], a[m][2], a[m][3]);
      }
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const int tap = warp + 8 * s;
        if (tap < 27) {
          const int kw = tap % 3, r9 = tap / 3;                         // r9 = kd * 3 + kh
          const __half* base = sX + r9 * xrow + (k0 + kw) * kRowHalfs;  // voxel w reads halo column w + kw
#pragma unroll
          for (int np = 0; np < 2; np++) {
            // B = X (voxel x ci): matrices (k0-7,n0-7), (k8-15,n0-7), (k0-7,n8-15), (k8-15,n8-15) of channel block np*16
            const __half* p = base + ((lj & 1) * 8 + li) * kRowHalfs + np * 16 + (lj >> 1) * 8;
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans((uint32_t)__cvta_generic_to_shared(p), b0, b1, b2, b3);
#pragma unroll
            for (int m = 0; m < 2; m++) {
              mma16816(acc[s][m][np * 2], a[m], b0, b1);
              mma16816(acc[s][m][np * 2 + 1], a[m], b2, b3);
            }
          }
        }
      }
    }
  }
  // ---- partial block of this CTA: [chunk][ci block][co block] -> [tap][co 32][ci 32] fp32
  const int g = lane >> 2, t4 = lane & 3;
  float* out = partial + (((long long)blockIdx.x * gridDim.y + blockIdx.y) * gridDim.z + blockIdx.z) * (27 * 1024);
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int tap = warp + 8 * s;
    if (tap < 27) {
#pragma unroll
      for (int m = 0; m < 2; m++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          float* o = out + tap * 1024 + (m * 16 + g) * 32 + nt * 8 + 2 * t4;
          *reinterpret_cast<float2*>(o) = make_float2(acc[s][m][nt][0], acc[s][m][nt][1]);
          *reinterpret_cast<float2*>(o + 8 * 32) = make_float2(acc[s][m][nt][2], acc[s][m][nt][3]);
        }
    }
  }
}

// fixed-order sum over the chunks, written in the PyTorch layout (Cout, Cin, 3, 3, 3)
__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ partial, int chunks, int Cin, int Cout, float out_scale,
                                         float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * Cin * 27) return;
  const int tap = i % 27, ci = (i / 27) % Cin, co = i / (27 * Cin);
  const int cib = Cin / 32, cob = Cout / 32;
  const long long blk = (long long)(ci / 32) * cob + co / 32;
  const long long off = blk * (27 * 1024) + tap * 1024 + (co % 32) * 32 + ci % 32;
  const long long stride = (long long)cib * cob * (27 * 1024);
  float s = 0.f;
  for (int c = 0; c < chunks; c++) s += partial[c * stride + off];
  dw[i] = s * out_scale;
}

int wgrad_chunks(int N, int D, int H, int Cin, int Cout) {
  const int blocks = (Cin / 32) * (Cout / 32);
  const long long rows = (long long)N * D * H;
  long long c = (2LL * nm_num_sms() + blocks - 1) / blocks;
  if (c > rows) c = rows;
  return (int)(c < 1 ? 1 : c);
}

}  // namespace

extern "C" size_t nm_conv3d_k3_wgrad_workspace_bytes(int N, int D, int H, int W, int Cin, int Cout) {
  (void)W;
  if (Cin <= 0 || Cout <= 0 || Cin % 32 || Cout % 32) return 0;
  return (size_t)wgrad_chunks(N, D, H, Cin, Cout) * (Cin / 32) * (Cout / 32) * 27 * 1024 * sizeof(float);
}

extern "C" int nm_conv3d_k3_wgrad(const void* x, const void* grad_out, int N, int D, int H, int W, int Cin, int Cout,
                                  float out_scale, float* dw, void* workspace, void* stream) {
  NM_CHECK_ARG(x && grad_out && dw && workspace, "nm_conv3d_k3_wgrad: null pointer");
  NM_CHECK_ARG(N > 0 && D > 0 && H > 0, "nm_conv3d_k3_wgrad: empty input");
  NM_CHECK_ARG(W % 16 == 0 && W >= 16 && W <= kMaxW, "nm_conv3d_k3_wgrad: W must be 16, 32, 48 or 64 (got %d)", W);
  NM_CHECK_ARG(Cin % 32 == 0 && Cout % 32 == 0 && Cin <= 256 && Cout <= 256,
               "nm_conv3d_k3_wgrad: Cin and Cout must be multiples of 32, <= 256 (got %d -> %d)", Cin, Cout);
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = wgrad_chunks(N, D, H, Cin, Cout);
  const size_t smem = (size_t)(kMaxW + 9 * (W + 2)) * kRowHalfs * sizeof(__half);
  // the attribute is per device: set it on every call (cheap) rather than once per process
  NM_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_k3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const dim3 grid(chunks, Cin / 32, Cout / 32);
  conv_wgrad_k3_kernel<<<grid, kWgThreads, smem, st>>>(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(grad_out),
                                                       N, D, H, W, Cin, Cout, reinterpret_cast<float*>(workspace));
  NM_CHECK_LAUNCH("conv_wgrad_k3_kernel");
  conv_wgrad_reduce_kernel<<<nm_cdiv((long long)Cout * Cin * 27, 256), 256, 0, st>>>(reinterpret_cast<const float*>(workspace),
                                                                                     chunks, Cin, Cout, out_scale, dw);
  NM_CHECK_LAUNCH("conv_wgrad_reduce_kernel");
  return NM_OK;
}
