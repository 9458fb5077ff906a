// Weight gradient of every convolution shape of the detector that the slab kernels do not cover (config #4 backward):
// 1x1 convs, the k2/s2 "pool" convs, ConvTranspose3d(k2, s2), and k3 convs on the small hour-glass grids or with
// 48 / 72 channels.  mma.sync (m16n8k16, fp16 operands, fp32 accumulate), fixed-order split-K -> bit-reproducible.
//
//   P[a][b][tap] = sum over the voxels u of the SMALL-side tensor  S[u][a] * L[u * stride + tap - pad][b]
//
//   nn.Conv3d          (modules/vox_modules.py:12,26,30,39,53):  S = dL/dy, L = x   -> P = dW (Cout, Cin, k, k, k)
//   nn.ConvTranspose3d (modules/vox_modules.py:68):              S = x, L = dL/dy   -> P = dW (Cin, Cout, 2, 2, 2)
//
// A CTA owns a 32 (a) x 32 (b) block of all taps and walks K tiles of 64 (128 for a single tap) consecutive S voxels.
// Per tile it gathers, with 16-byte cp.async (zero fill outside the tensor = the conv's padding), the S rows and - per
// tap - the L rows they pair with; pool / 1x1 convs use every L voxel exactly once, so the gather has no redundancy
// where the bytes matter.  Both operands are voxel-major in shared memory -> ldmatrix.trans fragments as in
// conv_wgrad.cu.  Warp w accumulates taps w, w + 8, ... (k3), tap w (k2), or its own K step (k1).
#include "common.cuh"
#include "../../include/nm_b200.h"

namespace {

constexpr int kGwThreads = 256;
constexpr int kGwStages = 2;                // ring depth of the streaming (few-tap) variants
constexpr int kGwRow = 40;                    // halfs per staged row: 32 channels + 8 pad (80 bytes, conflict-free ldmatrix)

__device__ __forceinline__ void gw_ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void gw_mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void gw_cp_async16(void* dst_smem, const void* src_gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src_gmem),
               "r"(src_bytes) : "memory");
}

struct GwShape {
  int N, Ds, Hs, Ws;            // extent of the small-side tensor
  int Ca, Cb;                   // channels of S / L
  int k, stride, pad;           // taps per axis; L extent = Ds * stride etc.
  long long U;                  // N * Ds * Hs * Ws
  int pow2, lw, lh, ld;         // Ds, Hs, Ws all powers of two: their log2 (voxel decomposition by shifts)
};

// SLOTS = taps per warp (k3: 4, k2: 1, k1: 1); KT = voxels per K tile; KK = taps per axis (compile time: the gather's
// address arithmetic - tap decomposition, voxel decomposition on power-of-two grids - was 5 600 warp instructions per
// 37 KB tile with run-time divisions, which bound the streaming layers at 3 TB/s)
template <int SLOTS, int KT, int KK>
__global__ void __launch_bounds__(kGwThreads, 1)
wgrad_gather_kernel(const __half* __restrict__ Sx, const __half* __restrict__ Lx, GwShape g, float* __restrict__ partial) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int taps = KK * KK * KK;
  __half* sS = reinterpret_cast<__half*>(smem);                          // per buffer: S tile [KT][kGwRow], L tiles [taps][KT][kGwRow]
  const int a0 = blockIdx.y * 32, b0 = blockIdx.z * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Dl = g.Ds * g.stride, Hl = g.Hs * g.stride, Wl = g.Ws * g.stride;

  float acc[SLOTS][2][4][4];
#pragma unroll
  for (int s = 0; s < SLOTS; s++)
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
      for (int n = 0; n < 4; n++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[s][m][n][e] = 0.f;

  const int lj = lane >> 3, li = lane & 7;
  const long long tiles = (g.U + KT - 1) / KT;
  // Few-tap layers (1x1, pool, transposed convs) are pure streaming: their tiles go through a ring of kStages buffers -
  // the gathers of the next kStages - 1 tiles are in flight while a tile feeds the MMAs (one tile ahead left the kernel
  // latency-bound at 3 TB/s).  The 27-tap tile (143 KB) has no room for a second buffer.
  constexpr bool kDouble = SLOTS == 1;
  constexpr int kStages = kDouble ? kGwStages : 1;
  const int buf_halfs = (1 + taps) * KT * kGwRow;
  auto stage = [&](long long tile, int buf) {
    __half* bS = sS + (size_t)buf * buf_halfs;
    __half* bL = bS + KT * kGwRow;
    for (int i = threadIdx.x; i < KT * 4; i += kGwThreads) {
      const int j = i >> 2, c = i & 3;
      const long long u = tile * KT + j;
      const bool uin = u < g.U;
      const int uu = uin ? (int)u : 0;                                   // U < 2^31 (checked by the host): 32-bit index math
      int w, h, d, n;
      if (g.pow2) {
        w = uu & (g.Ws - 1);
        h = (uu >> g.lw) & (g.Hs - 1);
        d = (uu >> (g.lw + g.lh)) & (g.Ds - 1);
        n = uu >> (g.lw + g.lh + g.ld);
      } else {
        const int q1 = uu / g.Ws, q2 = q1 / g.Hs;
        w = uu % g.Ws; h = q1 % g.Hs; d = q2 % g.Ds; n = q2 / g.Ds;
      }
      const bool aok = uin && (a0 + c * 8 < g.Ca);
      gw_cp_async16(bS + j * kGwRow + c * 8, aok ? Sx + (long long)uu * g.Ca + a0 + c * 8 : Sx, aok ? 16 : 0);
      const bool bok = uin && (b0 + c * 8 < g.Cb);
#pragma unroll
      for (int t = 0; t < taps; t++) {
        constexpr int kk2 = KK * KK;
        const int kw = t % KK, kh = (t / KK) % KK, kd = t / kk2;
        const int dd = d * g.stride + kd - g.pad, hh = h * g.stride + kh - g.pad, ww = w * g.stride + kw - g.pad;
        const bool ok = bok && (unsigned)dd < (unsigned)Dl && (unsigned)hh < (unsigned)Hl && (unsigned)ww < (unsigned)Wl;
        const __half* src = ok ? Lx + ((((long long)n * Dl + dd) * Hl + hh) * Wl + ww) * g.Cb + b0 + c * 8 : Lx;
        gw_cp_async16(bL + ((long long)t * KT + j) * kGwRow + c * 8, src, ok ? 16 : 0);
      }
    }
  };
  int buf = 0;
  if (kDouble) {
    // prologue: tiles 0 .. kStages - 2 of this CTA (one commit group per ring slot, empty past the end)
#pragma unroll
    for (int s = 0; s < kStages - 1; s++) {
      const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
      if (t < tiles) stage(t, s);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    if (kDouble) {
      const long long ahead = tile + (long long)(kStages - 1) * gridDim.x;
      if (ahead < tiles) stage(ahead, (buf + kStages - 1) % kStages);    // the slot consumed in the previous iteration
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");   // this tile has landed
    } else {
      __syncthreads();                                                   // the previous tile's fragments are consumed
      stage(tile, 0);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const __half* cS = sS + (size_t)buf * buf_halfs;
    const __half* cL = cS + KT * kGwRow;
#pragma unroll 1
    for (int k0 = 0; k0 < KT; k0 += 16) {
      if (taps == 1 && (k0 >> 4) != warp) continue;                      // k1: one K step per warp
      uint32_t a[2][4];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        const __half* p = cS + (k0 + (lj >> 1) * 8 + li) * kGwRow + m * 16 + (lj & 1) * 8;
        gw_ldmatrix_x4_trans((uint32_t)__cvta_generic_to_shared(p), a[m][0], a[m][1], a[m][2], a[m][3]);
      }
#pragma unroll
      for (int s = 0; s < SLOTS; s++) {
        const int tap = taps == 1 ? 0 : warp + 8 * s;
        if (tap < taps) {
          const __half* base = cL + ((long long)tap * KT + k0) * kGwRow;
#pragma unroll
          for (int np = 0; np < 2; np++) {
            const __half* p = base + ((lj & 1) * 8 + li) * kGwRow + np * 16 + (lj >> 1) * 8;
            uint32_t q0, q1, q2, q3;
            gw_ldmatrix_x4_trans((uint32_t)__cvta_generic_to_shared(p), q0, q1, q2, q3);
#pragma unroll
            for (int m = 0; m < 2; m++) {
              gw_mma16816(acc[s][m][np * 2], a[m], q0, q1);
              gw_mma16816(acc[s][m][np * 2 + 1], a[m], q2, q3);
            }
          }
        }
      }
    }
    if (kDouble) {
      __syncthreads();                                                   // this slot is refilled in the next iteration
      buf = (buf + 1) % kStages;
    }
  }
  // partial[part][a block][b block][tap][32][32]; part = chunk (k2 / k3) or chunk * 8 + warp (k1)
  const int gq = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int s = 0; s < SLOTS; s++) {
    const int tap = taps == 1 ? 0 : warp + 8 * s;
    if (tap < taps) {
      const long long part = taps == 1 ? (long long)blockIdx.x * 8 + warp : blockIdx.x;
      float* out = partial + (((part * gridDim.y + blockIdx.y) * gridDim.z + blockIdx.z) * taps + tap) * 1024;
#pragma unroll
      for (int m = 0; m < 2; m++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          float* o = out + (m * 16 + gq) * 32 + nt * 8 + 2 * t4;
          *reinterpret_cast<float2*>(o) = make_float2(acc[s][m][nt][0], acc[s][m][nt][1]);
          *reinterpret_cast<float2*>(o + 8 * 32) = make_float2(acc[s][m][nt][2], acc[s][m][nt][3]);
        }
    }
  }
}

// fixed-order sum over the parts -> P[a][b][tap] * out_scale (the nn.Conv3d / nn.ConvTranspose3d weight layout).
// blockDim (32, 8): 8 interleaved part groups per output element, combined in order (a serial loop over up to 296 parts
// per thread made the 58 launches of a training step cost 1.6 ms).
__global__ void __launch_bounds__(256)
wgrad_gather_reduce_kernel(const float* __restrict__ partial, int parts, int ablocks, int bblocks, int taps,
                           int Ca, int Cb, float out_scale, float* __restrict__ dw) {
  __shared__ float red[8][32];
  const long long i = (long long)blockIdx.x * 32 + threadIdx.x;
  const bool live = i < (long long)Ca * Cb * taps;
  float s = 0.f;
  if (live) {
    const int tap = (int)(i % taps), b = (int)((i / taps) % Cb), a = (int)(i / ((long long)taps * Cb));
    const long long off = ((((long long)(a / 32) * bblocks + b / 32) * taps) + tap) * 1024 + (a % 32) * 32 + b % 32;
    const long long stride = (long long)ablocks * bblocks * taps * 1024;
    for (int p = threadIdx.y; p < parts; p += 8) s += partial[p * stride + off];
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && live) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; y++) t += red[y][threadIdx.x];
    dw[i] = t * out_scale;
  }
}

int gw_tile(int k) { return k == 1 ? 128 : 64; }

int gw_chunks(const GwShape& g) {
  const int blocks = ((g.Ca + 31) / 32) * ((g.Cb + 31) / 32);
  const long long tiles = (g.U + gw_tile(g.k) - 1) / gw_tile(g.k);
  const long long per_sm = 2;                                          // resident CTAs per SM
  long long c = (per_sm * nm_num_sms() + blocks - 1) / blocks;
  if (c > tiles) c = tiles;
  return (int)(c < 1 ? 1 : c);
}

bool gw_shape(int N, int Ds, int Hs, int Ws, int Ca, int Cb, int k, int stride, GwShape* g) {
  if (N <= 0 || Ds <= 0 || Hs <= 0 || Ws <= 0 || Ca <= 0 || Cb <= 0 || Ca % 8 || Cb % 8 || Ca > 256 || Cb > 256) return false;
  if (!((stride == 1 && (k == 1 || k == 3)) || (stride == 2 && k == 2))) return false;
  if ((long long)N * Ds * Hs * Ws >= (1LL << 31)) return false;
  *g = GwShape{N, Ds, Hs, Ws, Ca, Cb, k, stride, stride == 1 ? (k - 1) / 2 : 0, (long long)N * Ds * Hs * Ws, 0, 0, 0, 0};
  auto lg = [](int v) { int l = 0; while ((1 << l) < v) l++; return l; };
  if (!(Ds & (Ds - 1)) && !(Hs & (Hs - 1)) && !(Ws & (Ws - 1))) { g->pow2 = 1; g->lw = lg(Ws); g->lh = lg(Hs); g->ld = lg(Ds); }
  return true;
}

}  // namespace

extern "C" size_t nm_conv3d_wgrad_gather_workspace_bytes(int N, int Ds, int Hs, int Ws, int Ca, int Cb, int k, int stride) {
  GwShape g;
  if (!gw_shape(N, Ds, Hs, Ws, Ca, Cb, k, stride, &g)) return 0;
  const size_t parts = (size_t)gw_chunks(g) * (k == 1 ? 8 : 1);
  return parts * ((Ca + 31) / 32) * ((Cb + 31) / 32) * (size_t)(k * k * k) * 1024 * sizeof(float);
}

extern "C" int nm_conv3d_wgrad_gather(const void* small_side, const void* large_side, int N, int Ds, int Hs, int Ws, int Ca,
                                      int Cb, int k, int stride, float out_scale, float* dw, void* workspace, void* stream) {
  NM_CHECK_ARG(small_side && large_side && dw && workspace, "nm_conv3d_wgrad_gather: null pointer");
  GwShape g;
  NM_CHECK_ARG(gw_shape(N, Ds, Hs, Ws, Ca, Cb, k, stride, &g),
               "nm_conv3d_wgrad_gather: need channels multiple of 8 (<= 256) and (k, stride) in {(1,1), (3,1), (2,2)}; got %d x %d, k=%d s=%d",
               Ca, Cb, k, stride);
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = gw_chunks(g), taps = k * k * k, KT = gw_tile(k);
  const int ab = (Ca + 31) / 32, bb = (Cb + 31) / 32;
  const size_t smem = (size_t)(1 + taps) * KT * kGwRow * sizeof(__half) * (k == 3 ? 1 : kGwStages);   // few-tap tiles: ring
  NM_CHECK_CUDA(cudaFuncSetAttribute(wgrad_gather_kernel<4, 64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  NM_CHECK_CUDA(cudaFuncSetAttribute(wgrad_gather_kernel<1, 64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  NM_CHECK_CUDA(cudaFuncSetAttribute(wgrad_gather_kernel<1, 128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  const dim3 grid(chunks, ab, bb);
  const __half* S = reinterpret_cast<const __half*>(small_side);
  const __half* L = reinterpret_cast<const __half*>(large_side);
  float* ws = reinterpret_cast<float*>(workspace);
  if (k == 3) wgrad_gather_kernel<4, 64, 3><<<grid, kGwThreads, smem, st>>>(S, L, g, ws);
  else if (k == 2) wgrad_gather_kernel<1, 64, 2><<<grid, kGwThreads, smem, st>>>(S, L, g, ws);
  else wgrad_gather_kernel<1, 128, 1><<<grid, kGwThreads, smem, st>>>(S, L, g, ws);
  NM_CHECK_LAUNCH("wgrad_gather_kernel");
  const int parts = chunks * (k == 1 ? 8 : 1);
  wgrad_gather_reduce_kernel<<<nm_cdiv((long long)Ca * Cb * taps, 32), dim3(32, 8), 0, st>>>(ws, parts, ab, bb, taps, Ca, Cb, out_scale, dw);
  NM_CHECK_LAUNCH("wgrad_gather_reduce_kernel");
  return NM_OK;
}
