// Heat-map heads, soft-argmax keypoints, Gaussian re-rendering and the decoder's
// 1x1 "adjust" convolution over the (never materialised) 179-channel concat.
//
// Reference:
//   heads            model/kypt_detector.py:273-297, 336-343   (LeakyReLU(1x1 conv) -> 2->1 conv -> Softplus;
//                    there is NO softmax: marginals of (hm + 1e-6) are sum-normalised)
//   soft-argmax      utils/kypt_detector_utils.py:28-55
//   Gaussian render  utils/kypt_detector_utils.py:57-90 (called per keypoint, kypt_detector.py:349-353)
//   adjust conv      model/kypt_detector.py:381,404-408 (cat[gauss_t, first_feature, gauss_0, coords] -> 128)
//
// One CTA per frame.  Shared-memory staging of the head weights, warp-shuffle reductions for
// the marginals; the feature map is read once (HBM-bound: g^3*C fp16 per frame).
#include "common.cuh"

namespace {

constexpr int KMAX = 24;   // keypoints (compile-time upper bound for register arrays)

__device__ __forceinline__ float softplus1(float x) {
  // nn.Softplus(beta=1, threshold=20)
  return x > 20.f ? x : log1pf(expf(x));
}

// Phase 1 (parallel over voxels): heat-map values.  grid = (g^3 / 256, n); one voxel per thread.
//   mode 0: heat = lrelu(w1 f + b1);  mode 1: heat = softplus(pw0 * lrelu(w1 f + b1) + pw1 * prev[clip] + pb)
template <int C>
__global__ void __launch_bounds__(256)
head_heat_kernel(const act_t* __restrict__ feature, const float* __restrict__ w1, const float* __restrict__ b1,
                 int K, int S, int mode, const float* __restrict__ prev, int frames_per_clip, float pw0, float pw1,
                 float pb, float* __restrict__ heat) {
  extern __shared__ float smem[];
  float* s_w = smem;                       // [K][C]
  float* s_b = s_w + KMAX * C;             // [K]
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < K * C; i += 256) s_w[i] = w1[i];
  for (int i = threadIdx.x; i < K; i += 256) s_b[i] = b1[i];
  __syncthreads();
  const int s = blockIdx.x * 256 + threadIdx.x;
  if (s >= S) return;
  float acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; k++) acc[k] = 0.f;
  const half8* p = reinterpret_cast<const half8*>(feature + ((long long)n * S + s) * C);
#pragma unroll 2
  for (int c8 = 0; c8 < C / 8; c8++) {
    float v[8];
    nm_unpack8(p[c8], v);
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
      if (k < K) {
        const float4 wa = *reinterpret_cast<const float4*>(s_w + k * C + c8 * 8);
        const float4 wb = *reinterpret_cast<const float4*>(s_w + k * C + c8 * 8 + 4);
        acc[k] = fmaf(v[0], wa.x, fmaf(v[1], wa.y, fmaf(v[2], wa.z, fmaf(v[3], wa.w, acc[k]))));
        acc[k] = fmaf(v[4], wb.x, fmaf(v[5], wb.y, fmaf(v[6], wb.z, fmaf(v[7], wb.w, acc[k]))));
      }
    }
  }
  const float* pv = prev ? prev + (long long)(n / frames_per_clip) * K * S : nullptr;
  float* hm_out = heat + (long long)n * K * S;
#pragma unroll
  for (int k = 0; k < KMAX; k++) {
    if (k < K) {
      float h = nm_lrelu(acc[k] + s_b[k]);
      if (mode == 1) h = softplus1(fmaf(pw0, h, fmaf(pw1, pv[(long long)k * S + s], pb)));
      hm_out[(long long)k * S + s] = h;
    }
  }
}

// Phase 2 (one CTA per frame): marginals of the heat-maps -> soft-argmax keypoints -> Gaussian re-render.
// heat: (n, K, g^3) fp32 written by head_heat_kernel.  g in {8, 16, 32}: a warp's 32 consecutive voxels share x;
// y is constant over runs of min(g, 32) lanes.  No float atomics: the result is bit-reproducible.
__global__ void __launch_bounds__(256)
head_reduce_kernel(int K, int g, const float* __restrict__ lin, float gauss_width, const float* __restrict__ heat,
                   float* __restrict__ keypoints, float* __restrict__ gaussians, float* __restrict__ heat_mean) {
  extern __shared__ float smem[];
  float* s_m = smem;                       // [3][K][32] raw marginal sums along x, y, z
  float* s_kp = s_m + 3 * KMAX * 32;       // [K][4]
  float* s_e = s_kp + KMAX * 4;            // [K][3][32] separable gaussian factors
  float* s_px = s_e + KMAX * 3 * 32;       // [8 warps][K]        per-iteration warp partials (x)
  float* s_py = s_px + 8 * KMAX;           // [8 warps][K][4]     (y: one per run of `run` lanes)
  float* s_pz = s_py + 8 * KMAX * 4;       // [8 warps][K][32]    (z)
  const int n = blockIdx.x;
  const int S = g * g * g;
  for (int i = threadIdx.x; i < 3 * KMAX * 32; i += 256) s_m[i] = 0.f;
  __syncthreads();

  const float* hm_out = heat + (long long)n * K * S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int run = g < 32 ? g : 32;          // lanes that share y
  const int runs = 32 / run;                // y rows per warp

  // voxel s = (x*g + y)*g + z ; consecutive threads -> consecutive s (coalesced heat-map rows)
  for (int s0 = 0; s0 < S; s0 += 256) {
    const int s = s0 + threadIdx.x;        // S is a multiple of 256 for g >= 8
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
      if (k < K) {
        const float h = hm_out[(long long)k * S + s];     // written by head_heat_kernel
        {
          // warp-level partial sums (shuffles only: deterministic)
          float hy = h;                                    // y: segmented sum over runs of `run` lanes
          for (int o = 1; o < run; o <<= 1) hy += __shfl_xor_sync(0xffffffffu, hy, o);
          float hz = h;                                    // z: sum over lanes with equal lane % g
          for (int o = run; o < 32; o <<= 1) hz += __shfl_xor_sync(0xffffffffu, hz, o);
          float hx = hy;                                   // x: whole warp
          for (int o = run; o < 32; o <<= 1) hx += __shfl_xor_sync(0xffffffffu, hx, o);
          if (lane == 0) s_px[warp * KMAX + k] = hx;
          if ((lane % run) == 0) s_py[(warp * KMAX + k) * 4 + lane / run] = hy;
          if (lane < run) s_pz[(warp * KMAX + k) * 32 + lane] = hz;
        }
      }
    }
    {
      // fold the 8 warps' partials into the marginals in a fixed order (no float atomics: results are
      // bit-reproducible run to run, which the reference's callers rely on via cudnn.deterministic)
      __syncthreads();
      for (int i = threadIdx.x; i < K * (1 + 32); i += 256) {
        const int k = i / 33, j = i % 33;
        if (j == 32) {
          for (int w = 0; w < 8; w++) {
            const int sw = s0 + w * 32;                    // first voxel of warp w in this iteration
            s_m[(0 * KMAX + k) * 32 + sw / (g * g)] += s_px[w * KMAX + k];
            for (int r = 0; r < runs; r++)
              s_m[(1 * KMAX + k) * 32 + ((sw + r * run) / g) % g] += s_py[(w * KMAX + k) * 4 + r];
          }
        } else if (j < run) {
          float t = 0.f;
          for (int w = 0; w < 8; w++) t += s_pz[(w * KMAX + k) * 32 + j];
          s_m[(2 * KMAX + k) * 32 + j] += t;
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();

  // ---- soft-argmax (utils/kypt_detector_utils.py:28-55)
  if (threadIdx.x < K) {
    const int k = threadIdx.x;
    float total = 0.f;
    for (int i = 0; i < g; i++) total += s_m[(0 * KMAX + k) * 32 + i];
    s_kp[k * 4 + 3] = total / (float)S;                         // mean (intensity numerator)
    if (heat_mean) heat_mean[(long long)n * K + k] = total / (float)S;   // get_keypoint_sparsity_loss input
    const float plane = (float)(g * g) * 1e-6f;                 // the +1e-6 summed over the other two axes
    for (int a = 0; a < 3; a++) {
      const float* m = s_m + (a * KMAX + k) * 32;
      float den = 0.f;
      for (int i = 0; i < g; i++) den += m[i] + plane;
      float c = 0.f;
      for (int i = 0; i < g; i++) c += ((m[i] + plane) / den) * lin[i];
      s_kp[k * 4 + a] = c;
    }
  }
  __syncthreads();
  float inten = 0.f;
  if (threadIdx.x < K) {
    float mx = -INFINITY;
    for (int k = 0; k < K; k++) mx = fmaxf(mx, s_kp[k * 4 + 3]);
    inten = s_kp[threadIdx.x * 4 + 3] / (mx + 1e-6f);
  }
  __syncthreads();
  if (threadIdx.x < K) s_kp[threadIdx.x * 4 + 3] = inten;
  __syncthreads();
  for (int i = threadIdx.x; i < K * 4; i += 256) keypoints[(long long)n * K * 4 + i] = s_kp[i];
  if (!gaussians) return;
  // ---- Gaussian re-render (utils/kypt_detector_utils.py:57-90): ((1*ex)*ey)*ez*I
  for (int i = threadIdx.x; i < K * 3 * g; i += 256) {
    const int k = i / (3 * g), a = (i / g) % 3, j = i % g;
    const float d = lin[j] - s_kp[k * 4 + a];
    s_e[(k * 3 + a) * 32 + j] = expf(-(d * d) / gauss_width);
  }
  __syncthreads();
  float* go = gaussians + (long long)n * K * S;
  for (int i = threadIdx.x; i < K * S; i += 256) {
    const int k = i / S, s = i % S;
    const int z = s % g, y = (s / g) % g, x = s / (g * g);
    go[i] = (s_e[(k * 3 + 0) * 32 + x] * s_e[(k * 3 + 1) * 32 + y]) * s_e[(k * 3 + 2) * 32 + z] * s_kp[k * 4 + 3];
  }
}

// Standalone render for decode_from_dyna (model/kypt_detector.py:213-231): keypoints (n,K,4) -> (n,K,g^3)
__global__ void __launch_bounds__(256)
gaussian_render_kernel(const float* __restrict__ keypoints, int K, int g, const float* __restrict__ lin,
                       float gauss_width, float* __restrict__ gaussians) {
  __shared__ float s_e[KMAX * 3 * 32];
  __shared__ float s_i[KMAX];
  const int n = blockIdx.x;
  const int S = g * g * g;
  const float* kp = keypoints + (long long)n * K * 4;
  for (int i = threadIdx.x; i < K * 3 * g; i += 256) {
    const int k = i / (3 * g), a = (i / g) % 3, j = i % g;
    const float d = lin[j] - kp[k * 4 + a];
    s_e[(k * 3 + a) * 32 + j] = expf(-(d * d) / gauss_width);
  }
  for (int i = threadIdx.x; i < K; i += 256) s_i[i] = kp[i * 4 + 3];
  __syncthreads();
  float* go = gaussians + (long long)n * K * S;
  for (int i = threadIdx.x; i < K * S; i += 256) {
    const int k = i / S, s = i % S;
    const int z = s % g, y = (s / g) % g, x = s / (g * g);
    go[i] = ((1.0f * s_e[(k * 3 + 0) * 32 + x]) * s_e[(k * 3 + 1) * 32 + y]) * s_e[(k * 3 + 2) * 32 + z] * s_i[k];
  }
}

// ------------------------------------------------------------------ decoder adjust conv
// combined = cat[gauss_t (K), first_feature (F=128), gauss_0 (K), coords (3)]; out = lrelu(W comb + b), W (CO, 2K+F+3).
// Everything but gauss_t is constant over the frames of a clip, so it is hoisted:
//   base[clip][s][co] = b[co] + W[:,K:K+F] ff[clip][s] + W[:,K+F:2K+F] gauss_0[clip][s] + W[:,2K+F:] coords(s)
//   out[frame][s][co] = lrelu(base[clip][s][co] + W[:, :K] gauss_t[frame][s])
// Gaussians are recomputed from the keypoints (separable exps) unless a gaussians tensor is passed.
template <int CO, int FD>
__global__ void __launch_bounds__(256)
adjust_base_kernel(const act_t* __restrict__ ff /* (clips, S, FD) */, const float* __restrict__ kp /* (n,K,4) */,
                   const float* __restrict__ gs /* optional (n,K,S) */, int frames_per_clip,
                   const float* __restrict__ w, const float* __restrict__ bias, int K, int g,
                   const float* __restrict__ lin, float gauss_width, float* __restrict__ base) {
  extern __shared__ float smem[];
  const int ld = 2 * K + FD + 3;
  float* s_w = smem;                 // [CO][FD + K + 3] (columns K .. end of W)
  float* s_e = s_w + CO * (FD + K + 3);  // [K][3][32]
  float* s_i = s_e + KMAX * 3 * 32;  // [K]
  const int clip = blockIdx.y;
  const int S = g * g * g;
  const int wcols = FD + K + 3;
  // gauss_0 = the gaussians of the clip's first frame (kypt_detector.py:406: gaussians[:, 0])
  const float* kp0 = kp ? kp + (long long)clip * frames_per_clip * K * 4 : nullptr;
  const float* g0 = gs ? gs + (long long)clip * frames_per_clip * K * S : nullptr;
  for (int i = threadIdx.x; i < CO * wcols; i += 256) s_w[i] = w[(i / wcols) * ld + K + (i % wcols)];
  if (!g0) {
    for (int i = threadIdx.x; i < K * 3 * g; i += 256) {
      const int k = i / (3 * g), a = (i / g) % 3, j = i % g;
      const float d = lin[j] - kp0[k * 4 + a];
      s_e[(k * 3 + a) * 32 + j] = expf(-(d * d) / gauss_width);
    }
    for (int i = threadIdx.x; i < K; i += 256) s_i[i] = kp0[i * 4 + 3];
  }
  __syncthreads();
  // thread -> (voxel, output-channel quarter): 4 threads per voxel, CO/4 outputs each
  const int q = threadIdx.x & 3;
  const int s = blockIdx.x * 64 + (threadIdx.x >> 2);
  if (s >= S) return;
  const int z = s % g, y = (s / g) % g, x = s / (g * g);
  constexpr int PER = CO / 4;
  float acc[PER];
#pragma unroll
  for (int j = 0; j < PER; j++) acc[j] = bias[q * PER + j];
  const half8* p = reinterpret_cast<const half8*>(ff + ((long long)clip * S + s) * FD);
  for (int c8 = 0; c8 < FD / 8; c8++) {
    float v[8];
    nm_unpack8(p[c8], v);
#pragma unroll
    for (int j = 0; j < PER; j++) {
      const float* wr = s_w + (q * PER + j) * wcols + c8 * 8;
#pragma unroll
      for (int e = 0; e < 8; e++) acc[j] = fmaf(v[e], wr[e], acc[j]);
    }
  }
  for (int k = 0; k < K; k++) {
    float gv;
    if (g0) gv = g0[(long long)k * S + s];
    else gv = ((1.0f * s_e[(k * 3) * 32 + x]) * s_e[(k * 3 + 1) * 32 + y]) * s_e[(k * 3 + 2) * 32 + z] * s_i[k];
#pragma unroll
    for (int j = 0; j < PER; j++) acc[j] = fmaf(gv, s_w[(q * PER + j) * wcols + FD + k], acc[j]);
  }
  const float cx = lin[x], cy = lin[y], cz = lin[z];
#pragma unroll
  for (int j = 0; j < PER; j++) {
    const float* wr = s_w + (q * PER + j) * wcols + FD + K;
    acc[j] = fmaf(cx, wr[0], fmaf(cy, wr[1], fmaf(cz, wr[2], acc[j])));
  }
  float* dst = base + ((long long)clip * S + s) * CO + q * PER;
#pragma unroll
  for (int j = 0; j < PER; j++) dst[j] = acc[j];
}

template <int CO>
__global__ void __launch_bounds__(256)
adjust_frame_kernel(const float* __restrict__ base, const float* __restrict__ kp /* (n,K,4) */,
                    const float* __restrict__ gs /* optional (n,K,S) */, const float* __restrict__ w, int ld,
                    int K, int g, int frames_per_clip, const float* __restrict__ lin, float gauss_width,
                    act_t* __restrict__ out) {
  // block: 64 voxels of one frame; thread <-> (voxel slot tid / 16, 8 output channels tid % 16), four rounds of 16
  // voxels.  W^T [K][CO] and the Gaussian values of the 64 voxels are staged in shared memory; base reads and
  // output stores are 16/32-byte vectors, contiguous across the 16 threads of a voxel.
  __shared__ float s_wt[KMAX * CO];      // [K][CO]
  __shared__ float s_e[KMAX * 3 * 32];
  __shared__ float s_i[KMAX];
  __shared__ float s_g[64 * KMAX];       // [voxel][K]
  const int n = blockIdx.y, clip = n / frames_per_clip;
  const int S = g * g * g;
  for (int i = threadIdx.x; i < CO * K; i += 256) s_wt[(i % K) * CO + i / K] = w[(i / K) * ld + (i % K)];
  if (!gs) {
    for (int i = threadIdx.x; i < K * 3 * g; i += 256) {
      const int k = i / (3 * g), a = (i / g) % 3, j = i % g;
      const float d = lin[j] - kp[((long long)n * K + k) * 4 + a];
      s_e[(k * 3 + a) * 32 + j] = expf(-(d * d) / gauss_width);
    }
    for (int i = threadIdx.x; i < K; i += 256) s_i[i] = kp[((long long)n * K + i) * 4 + 3];
  }
  __syncthreads();
  const int s0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < 64 * K; i += 256) {
    const int v = i / K, k = i % K, s = s0 + v;
    float gv = 0.f;
    if (s < S) {
      if (gs) gv = gs[((long long)n * K + k) * S + s];
      else {
        const int z = s % g, y = (s / g) % g, x = s / (g * g);
        gv = ((1.0f * s_e[(k * 3) * 32 + x]) * s_e[(k * 3 + 1) * 32 + y]) * s_e[(k * 3 + 2) * 32 + z] * s_i[k];
      }
    }
    s_g[v * KMAX + k] = gv;
  }
  __syncthreads();
  const int c8 = threadIdx.x & 15;
#pragma unroll 1
  for (int round = 0; round < 4; round++) {
    const int v = round * 16 + (threadIdx.x >> 4), s = s0 + v;
    if (s >= S) continue;
    float acc[8];
    const float4* bp = reinterpret_cast<const float4*>(base + ((long long)clip * S + s) * CO + c8 * 8);
    const float4 b0 = bp[0], b1 = bp[1];
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    for (int k = 0; k < K; k++) {
      const float gv = s_g[v * KMAX + k];
      const float4 w0 = *reinterpret_cast<const float4*>(s_wt + k * CO + c8 * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(s_wt + k * CO + c8 * 8 + 4);
      acc[0] = fmaf(gv, w0.x, acc[0]); acc[1] = fmaf(gv, w0.y, acc[1]); acc[2] = fmaf(gv, w0.z, acc[2]);
      acc[3] = fmaf(gv, w0.w, acc[3]); acc[4] = fmaf(gv, w1.x, acc[4]); acc[5] = fmaf(gv, w1.y, acc[5]);
      acc[6] = fmaf(gv, w1.z, acc[6]); acc[7] = fmaf(gv, w1.w, acc[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = nm_lrelu(acc[j]);
    *reinterpret_cast<half8*>(out + ((long long)n * S + s) * CO + c8 * 8) = nm_pack8(acc);
  }
}

}  // namespace

extern "C" int nm_heatmap_head(const void* feature, const float* w1, const float* b1, int n, int g, int C, int K,
                               int mode, const float* prev, int frames_per_clip, float pw0, float pw1, float pb,
                               const float* linspace, float gauss_width, float* heat, float* keypoints,
                               float* gaussians, float* heat_mean, void* stream) {
  NM_CHECK_ARG(feature && w1 && b1 && heat && linspace, "nm_heatmap_head: null pointer");
  NM_CHECK_ARG(K <= KMAX && (g == 8 || g == 16 || g == 32), "nm_heatmap_head: K=%d g=%d unsupported", K, g);
  NM_CHECK_ARG(mode == 0 || (prev && keypoints), "nm_heatmap_head: mode 1 needs prev and keypoints");
  if (n == 0) return NM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int S = g * g * g;
  const size_t smem1 = (size_t)(KMAX * C + KMAX) * sizeof(float);
  dim3 grid1(nm_cdiv(S, 256), n);
  if (C == 128) {
    head_heat_kernel<128><<<grid1, 256, smem1, st>>>((const act_t*)feature, w1, b1, K, S, mode, prev, frames_per_clip,
                                                     pw0, pw1, pb, heat);
  } else if (C == 256) {
    head_heat_kernel<256><<<grid1, 256, smem1, st>>>((const act_t*)feature, w1, b1, K, S, mode, prev, frames_per_clip,
                                                     pw0, pw1, pb, heat);
  } else {
    NM_CHECK_ARG(false, "nm_heatmap_head: C=%d unsupported", C);
  }
  NM_CHECK_LAUNCH("heatmap_head(heat)");
  if (mode == 0) return NM_OK;
  const size_t smem2 = (size_t)(3 * KMAX * 32 + KMAX * 4 + KMAX * 3 * 32 + 8 * KMAX * (1 + 4 + 32)) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(head_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr = true;
  }
  head_reduce_kernel<<<n, 256, smem2, st>>>(K, g, linspace, gauss_width, heat, keypoints, gaussians, heat_mean);
  NM_CHECK_LAUNCH("heatmap_head(reduce)");
  return NM_OK;
}

extern "C" int nm_gaussian_render(const float* keypoints, int n, int K, int g, const float* linspace,
                                  float gauss_width, float* gaussians, void* stream) {
  NM_CHECK_ARG(keypoints && linspace && gaussians, "nm_gaussian_render: null pointer");
  NM_CHECK_ARG(K <= KMAX && g <= 32, "nm_gaussian_render: K=%d g=%d unsupported", K, g);
  if (n == 0) return NM_OK;
  const float width = gauss_width;
  gaussian_render_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(keypoints, K, g, linspace, width, gaussians);
  NM_CHECK_LAUNCH("gaussian_render");
  return NM_OK;
}

extern "C" int nm_decoder_adjust(const void* first_feature, const float* keypoints, const float* gaussians,
                                 const float* weight, const float* bias, int n_clips, int frames_per_clip, int g,
                                 int K, const float* linspace, float gauss_width, float* base_ws, void* out,
                                 void* stream) {
  NM_CHECK_ARG(first_feature && weight && bias && base_ws && out && linspace, "nm_decoder_adjust: null pointer");
  NM_CHECK_ARG(keypoints || gaussians, "nm_decoder_adjust: need keypoints or gaussians");
  NM_CHECK_ARG(K <= KMAX && g <= 32, "nm_decoder_adjust: K=%d g=%d unsupported", K, g);
  if (n_clips == 0) return NM_OK;
  constexpr int CO = 128, FD = 128;
  const int S = g * g * g;
  const float width = gauss_width;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(CO * (FD + K + 3) + KMAX * 3 * 32 + KMAX) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(adjust_base_kernel<CO, FD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr = true;
  }
  const int n = n_clips * frames_per_clip;
  adjust_base_kernel<CO, FD><<<dim3(nm_cdiv(S, 64), n_clips), 256, smem, st>>>(
      (const act_t*)first_feature, gaussians ? nullptr : keypoints, gaussians, frames_per_clip, weight, bias, K, g,
      linspace, width, base_ws);
  NM_CHECK_LAUNCH("decoder_adjust(base)");
  adjust_frame_kernel<CO><<<dim3(nm_cdiv(S, 64), n), 256, 0, st>>>(base_ws, gaussians ? nullptr : keypoints, gaussians, weight,
                                                                   2 * K + FD + 3, K, g, frames_per_clip, linspace,
                                                                   width, (act_t*)out);
  NM_CHECK_LAUNCH("decoder_adjust(frame)");
  return NM_OK;
}
