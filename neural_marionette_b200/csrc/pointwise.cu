// HBM-bound glue between the convolutions: GroupNorm statistics / apply, residual
// adds, trilinear up-sampling, layout conversion, the final 1x1 conv + activation +
// BCE, the chamfer volume-fitting loss and the per-clip frame mean.
// All activations are fp16 channels-last (N, S, C) with S = D*H*W; every kernel
// moves 16-byte vectors (8 channels) per thread.
#include "common.cuh"
#include "../../include/nm_b200.h"   // the definitions below must match the public declarations

namespace {

// ------------------------------------------------------------------ GroupNorm statistics
// Reference: nn.GroupNorm(C // 16, C), eps 1e-5 (modules/vox_modules.py:14,28,32,41,55,70).
// Stage 1: per (sample, chunk, channel) partial sum / sum-of-squares in fp32.
// Stage 2 (finalize): fixed-order reduction over chunks and the channels of a group in
// float64 -> per (sample, channel) scale a = gamma*rstd and shift b = beta - mean*a.
// Deterministic (no float atomics): callers set cudnn.deterministic (train.py:136).
constexpr int kStatThreads = 288;  // divisible by C/8 for C in {32,48,64,72,128,256}

__global__ void __launch_bounds__(kStatThreads)
gn_stats_kernel(const act_t* __restrict__ x, int S, int C, int chunks, float* __restrict__ partial) {
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int c8n = C >> 3;
  const int lanes = kStatThreads / c8n;      // voxels handled in parallel
  const int c8 = threadIdx.x % c8n, vl = threadIdx.x / c8n;
  const int per = (S + chunks - 1) / chunks;
  const int s0 = chunk * per, s1 = min(S, s0 + per);
  float sum[8], sq[8];
#pragma unroll
  for (int i = 0; i < 8; i++) sum[i] = sq[i] = 0.f;
  const half8* base = reinterpret_cast<const half8*>(x + (long long)n * S * C);
  if (vl < lanes) {
    for (int s = s0 + vl; s < s1; s += lanes) {
      float f[8];
      nm_unpack8(base[(long long)s * c8n + c8], f);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        sum[i] += f[i];
        sq[i] += f[i] * f[i];
      }
    }
  }
  __shared__ float red[kStatThreads][17];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    red[threadIdx.x][i] = sum[i];
    red[threadIdx.x][8 + i] = sq[i];
  }
  __syncthreads();
  // thread (c8, j<16) reduces statistic j of channel-octet c8 over the voxel lanes
  for (int idx = threadIdx.x; idx < c8n * 16; idx += kStatThreads) {
    const int o = idx / 16, j = idx % 16;
    float acc = 0.f;
    for (int l = 0; l < lanes; l++) acc += red[l * c8n + o][j];
    const int ch = o * 8 + (j & 7);
    partial[(((long long)n * chunks + chunk) * C + ch) * 2 + (j >> 3)] = acc;
  }
}

__global__ void __launch_bounds__(256)
gn_finalize_kernel(const float* __restrict__ partial, int S, int C, int groups, int chunks,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_rstd /* optional (n, groups, 2) */,
                   float* __restrict__ xsum /* optional (n, C): per-channel sums (GroupNorm backward) */) {
  // one block per (sample, group): 256 threads stride over the (chunk, channel-in-group) partials in a fixed
  // order, fp64 accumulation, shuffle tree + fixed-order fold of the 8 warps -> deterministic.  (The conv
  // epilogues emit up to 1024 chunks per sample; a single warp per group was latency-bound at ~20 us.)
  const int n = blockIdx.x, g = blockIdx.y;
  const int cpg = C / groups;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double sum = 0.0, sq = 0.0;
  const int items = chunks * cpg;
  for (int i = threadIdx.x; i < items; i += 256) {
    const int k = i / cpg, c = g * cpg + i % cpg;
    const float2 v = *reinterpret_cast<const float2*>(partial + (((long long)n * chunks + k) * C + c) * 2);
    sum += (double)v.x;
    sq += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  __shared__ double s_part[8][2];
  __shared__ double s_stat[2];
  if (lane == 0) { s_part[warp][0] = sum; s_part[warp][1] = sq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, q = 0.0;
    for (int w = 0; w < 8; w++) { a += s_part[w][0]; q += s_part[w][1]; }
    const double cnt = (double)S * cpg;
    const double mean = a / cnt;
    const double var = fmax(q / cnt - mean * mean, 0.0);
    s_stat[0] = mean;
    s_stat[1] = 1.0 / sqrt(var + (double)eps);
    if (mean_rstd) {
      mean_rstd[((long long)n * groups + g) * 2] = (float)mean;
      mean_rstd[((long long)n * groups + g) * 2 + 1] = (float)s_stat[1];
    }
  }
  if (xsum) {
    // per-channel totals in a fixed order (one thread per channel of the group)
    for (int c = g * cpg + threadIdx.x; c < (g + 1) * cpg; c += 256) {
      double t = 0.0;
      for (int k = 0; k < chunks; k++) t += (double)partial[(((long long)n * chunks + k) * C + c) * 2];
      xsum[(long long)n * C + c] = (float)t;
    }
  }
  __syncthreads();
  for (int c = g * cpg + threadIdx.x; c < (g + 1) * cpg; c += 256) {
    const double a = (double)gamma[c] * s_stat[1];
    scale[(long long)n * C + c] = (float)a;
    shift[(long long)n * C + c] = (float)((double)beta[c] - s_stat[0] * a);
  }
}

// ------------------------------------------------------------------ affine (+act) (+second operand)
// out = act1(x1 * a1 + b1) + (x2 * a2 + b2 | x2 | nothing)
//   Basic/Pool block:      lrelu(GN(x1))
//   Res block output:      GN(res) + GN(skip)  |  GN(res) + x        (final leaky_relu(.,True) == identity)
//   HG up-sample + skip:   lrelu(GN(x1)) + x2
__global__ void __launch_bounds__(256)
affine_act_kernel(const act_t* __restrict__ x1, const float* __restrict__ a1, const float* __restrict__ b1,
                  int act1, const act_t* __restrict__ x2, const float* __restrict__ a2,
                  const float* __restrict__ b2, act_t* __restrict__ out, int S, int C, long long total8) {
  const int c8n = C >> 3;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total8; i += (long long)gridDim.x * 256) {
    const int c8 = (int)(i % c8n);
    const long long n = i / ((long long)S * c8n);
    float f[8], g[8];
    nm_unpack8(reinterpret_cast<const half8*>(x1)[i], f);
    const float4* pa = reinterpret_cast<const float4*>(a1 + n * C + c8 * 8);
    const float4* pb = reinterpret_cast<const float4*>(b1 + n * C + c8 * 8);
    float4 A0 = pa[0], A1 = pa[1], B0 = pb[0], B1 = pb[1];
    const float av[8] = {A0.x, A0.y, A0.z, A0.w, A1.x, A1.y, A1.z, A1.w};
    const float bv[8] = {B0.x, B0.y, B0.z, B0.w, B1.x, B1.y, B1.z, B1.w};
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float v = fmaf(f[k], av[k], bv[k]);
      f[k] = act1 ? nm_lrelu(v) : v;
    }
    if (x2) {
      nm_unpack8(reinterpret_cast<const half8*>(x2)[i], g);
      if (a2) {
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] += fmaf(g[k], a2[n * C + c8 * 8 + k], b2[n * C + c8 * 8 + k]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] += g[k];
      }
    }
    reinterpret_cast<half8*>(out)[i] = nm_pack8(f);
  }
}

// Same operation for C in {8, 16, 32, 64, 128, 256}: grid (voxel chunks, samples), thread = (voxel lane, 16-byte channel
// chunk).  The chunk's scale / shift live in registers and the voxel loop has no index arithmetic beyond one add - the
// flat-index kernel above spends two 64-bit divisions and up to 36 parameter loads per 16 bytes (instruction-bound).
template <bool HAS2, bool AFF2>
__global__ void __launch_bounds__(256)
affine_act_rows_kernel(const act_t* __restrict__ x1, const float* __restrict__ a1, const float* __restrict__ b1,
                       int act1, const act_t* __restrict__ x2, const float* __restrict__ a2,
                       const float* __restrict__ b2, act_t* __restrict__ out, int S, int C) {
  const int c8n = C >> 3, nvl = 256 / c8n;
  const int n = blockIdx.y, c8 = threadIdx.x % c8n, vl = threadIdx.x / c8n;
  float av[8], bv[8], a2v[8], b2v[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    av[k] = a1[(long long)n * C + c8 * 8 + k];
    bv[k] = b1[(long long)n * C + c8 * 8 + k];
    a2v[k] = AFF2 ? a2[(long long)n * C + c8 * 8 + k] : 1.f;
    b2v[k] = AFF2 ? b2[(long long)n * C + c8 * 8 + k] : 0.f;
  }
  const long long base = (long long)n * S * C + c8 * 8;
  const int step = gridDim.x * nvl;
#pragma unroll 2
  for (int v = blockIdx.x * nvl + vl; v < S; v += step) {
    const long long off = base + (long long)v * C;
    float f[8], g[8];
    nm_unpack8(*reinterpret_cast<const half8*>(x1 + off), f);
    if (HAS2) nm_unpack8(*reinterpret_cast<const half8*>(x2 + off), g);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float v1 = fmaf(f[k], av[k], bv[k]);
      f[k] = act1 ? nm_lrelu(v1) : v1;
      if (HAS2) f[k] += AFF2 ? fmaf(g[k], a2v[k], b2v[k]) : g[k];
    }
    *reinterpret_cast<half8*>(out + off) = nm_pack8(f);
  }
}

// ------------------------------------------------------------------ trilinear x2 (align_corners=False)
// Reference: nn.Upsample(scale_factor=2, mode='trilinear') at model/kypt_detector.py:427,441.
// Optional fused prologue: v = lrelu(x*a + b) of the producing GroupNorm.
// One thread per (input cell, 8 channels): it reads the 3x3x3 clamped neighbourhood once (27 16-byte loads)
// and writes the 2x2x2 output block.  For scale 2 the source position of output 2i+a is i-0.25 (a=0) or
// i+0.25 (a=1): weights (0.25, 0.75) on (i-1, i) resp. (0.75, 0.25) on (i, i+1), replicate-clamped at the
// borders (identical to PyTorch's index clamping).  Separable: lerp along w, then h, then accumulate along d.
template <bool AFFINE>
__global__ void __launch_bounds__(128)
upsample2x_kernel(const act_t* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b, int act,
                  act_t* __restrict__ out, int D, int H, int W, int C, long long total8) {
  const int c8n = C >> 3;
  const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
  if (i >= total8) return;
  int c8, w, h, d;
  long long n;
  if (total8 <= 0x7fffffffLL) {                   // 32-bit index arithmetic (a 64-bit division costs ~4x as much)
    unsigned r = (unsigned)i;
    c8 = (int)(r % (unsigned)c8n); r /= (unsigned)c8n;
    w = (int)(r % (unsigned)W); r /= (unsigned)W;
    h = (int)(r % (unsigned)H); r /= (unsigned)H;
    d = (int)(r % (unsigned)D);
    n = r / (unsigned)D;
  } else {
    c8 = (int)(i % c8n);
    long long r = i / c8n;
    w = (int)(r % W); r /= W;
    h = (int)(r % H); r /= H;
    d = (int)(r % D);
    n = r / D;
  }
  float av[8], bv[8];
  if (AFFINE) {
    const float4* pa = reinterpret_cast<const float4*>(a + n * C + c8 * 8);
    const float4* pb = reinterpret_cast<const float4*>(b + n * C + c8 * 8);
    const float4 A0 = pa[0], A1 = pa[1], B0 = pb[0], B1 = pb[1];
    av[0] = A0.x; av[1] = A0.y; av[2] = A0.z; av[3] = A0.w; av[4] = A1.x; av[5] = A1.y; av[6] = A1.z; av[7] = A1.w;
    bv[0] = B0.x; bv[1] = B0.y; bv[2] = B0.z; bv[3] = B0.w; bv[4] = B1.x; bv[5] = B1.y; bv[6] = B1.z; bv[7] = B1.w;
  }
  const half8* base = reinterpret_cast<const half8*>(x) + n * (long long)D * H * W * c8n;
  const int wi[3] = {max(w - 1, 0), w, min(w + 1, W - 1)};
  const int hi[3] = {max(h - 1, 0), h, min(h + 1, H - 1)};
  const int di[3] = {max(d - 1, 0), d, min(d + 1, D - 1)};
  float acc[8][8];       // [output (a,b,c)][channel]
#pragma unroll
  for (int o = 0; o < 8; o++)
#pragma unroll
    for (int k = 0; k < 8; k++) acc[o][k] = 0.f;
#pragma unroll
  for (int dd = 0; dd < 3; dd++) {
    float row[3][2][8];  // after the w-lerp: [hh][c][channel]
#pragma unroll
    for (int hh = 0; hh < 3; hh++) {
      float v[3][8];
#pragma unroll
      for (int ww = 0; ww < 3; ww++) {
        nm_unpack8(base[((long long)(di[dd] * H + hi[hh]) * W + wi[ww]) * c8n + c8], v[ww]);
        if (AFFINE) {
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const float t = fmaf(v[ww][k], av[k], bv[k]);
            v[ww][k] = act ? nm_lrelu(t) : t;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        row[hh][0][k] = fmaf(0.25f, v[0][k], 0.75f * v[1][k]);
        row[hh][1][k] = fmaf(0.25f, v[2][k], 0.75f * v[1][k]);
      }
    }
    // h-lerp -> plane[b][c], then accumulate along d with weights (a=0: .25,.75,0 ; a=1: 0,.75,.25)
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float p0 = fmaf(0.25f, row[0][c][k], 0.75f * row[1][c][k]);   // b = 0
        const float p1 = fmaf(0.25f, row[2][c][k], 0.75f * row[1][c][k]);   // b = 1
        if (dd == 0) { acc[0 + c][k] += 0.25f * p0; acc[2 + c][k] += 0.25f * p1; }
        if (dd == 1) {
          acc[0 + c][k] = fmaf(0.75f, p0, acc[0 + c][k]); acc[2 + c][k] = fmaf(0.75f, p1, acc[2 + c][k]);
          acc[4 + c][k] = fmaf(0.75f, p0, acc[4 + c][k]); acc[6 + c][k] = fmaf(0.75f, p1, acc[6 + c][k]);
        }
        if (dd == 2) { acc[4 + c][k] = fmaf(0.25f, p0, acc[4 + c][k]); acc[6 + c][k] = fmaf(0.25f, p1, acc[6 + c][k]); }
      }
  }
  const int OH = 2 * H, OW = 2 * W;
  half8* ob = reinterpret_cast<half8*>(out) + n * (long long)8 * D * H * W * c8n;
#pragma unroll
  for (int o = 0; o < 8; o++) {
    const int oa = o >> 2, obb = (o >> 1) & 1, oc = o & 1;
    ob[((long long)((2 * d + oa) * OH + 2 * h + obb) * OW + 2 * w + oc) * c8n + c8] = nm_pack8(acc[o]);
  }
}

// Same up-sampling without the fused prologue, computed in packed half2 arithmetic (lerp as x0 + q*(x1-x0)):
// the kernel above is instruction-bound (27 fp32 affine+LeakyReLU vectors and 38 fp32 lerp vectors per
// thread); applying the GroupNorm affine in a separate HBM-bound pass over the 8x smaller input and doing
// the interpolation here on half2 lanes needs ~2.5x fewer instructions.  Decoder path only.
__device__ __forceinline__ uint4 h2_lerp4(const uint4& x0, const uint4& x1, __half2 q) {
  // x0 + q * (x1 - x0), 4 half2 lanes
  uint4 r;
  const __half2* a = reinterpret_cast<const __half2*>(&x0);
  const __half2* b = reinterpret_cast<const __half2*>(&x1);
  __half2* o = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; i++) o[i] = __hfma2(q, __hsub2(b[i], a[i]), a[i]);
  return r;
}

__global__ void __launch_bounds__(128)
upsample2x_h2_kernel(const act_t* __restrict__ x, act_t* __restrict__ out, int D, int H, int W, int C,
                     long long total8) {
  const int c8n = C >> 3;
  const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
  if (i >= total8) return;
  int c8, w, h, d;
  long long n;
  if (total8 <= 0x7fffffffLL) {                   // 32-bit index arithmetic (a 64-bit division costs ~4x as much)
    unsigned r = (unsigned)i;
    c8 = (int)(r % (unsigned)c8n); r /= (unsigned)c8n;
    w = (int)(r % (unsigned)W); r /= (unsigned)W;
    h = (int)(r % (unsigned)H); r /= (unsigned)H;
    d = (int)(r % (unsigned)D);
    n = r / (unsigned)D;
  } else {
    c8 = (int)(i % c8n);
    long long r = i / c8n;
    w = (int)(r % W); r /= W;
    h = (int)(r % H); r /= H;
    d = (int)(r % D);
    n = r / D;
  }
  const uint4* base = reinterpret_cast<const uint4*>(x) + n * (long long)D * H * W * c8n;
  const int wi[3] = {max(w - 1, 0), w, min(w + 1, W - 1)};
  const int hi[3] = {max(h - 1, 0), h, min(h + 1, H - 1)};
  const int di[3] = {max(d - 1, 0), d, min(d + 1, D - 1)};
  const __half2 q = __floats2half2_rn(0.25f, 0.25f);
  uint4 plane[3][2][2];      // [dd][b][c] after the w- and h-lerps
#pragma unroll
  for (int dd = 0; dd < 3; dd++) {
    uint4 row[3][2];
#pragma unroll
    for (int hh = 0; hh < 3; hh++) {
      const uint4* p = base + ((long long)(di[dd] * H + hi[hh]) * W) * c8n + c8;
      const uint4 vm = __ldg(p + (long long)wi[0] * c8n), v0 = __ldg(p + (long long)wi[1] * c8n),
                  vp = __ldg(p + (long long)wi[2] * c8n);
      row[hh][0] = h2_lerp4(v0, vm, q);       // output 2w   : 0.75 x[w] + 0.25 x[w-1]
      row[hh][1] = h2_lerp4(v0, vp, q);       // output 2w+1 : 0.75 x[w] + 0.25 x[w+1]
    }
#pragma unroll
    for (int c = 0; c < 2; c++) {
      plane[dd][0][c] = h2_lerp4(row[1][c], row[0][c], q);
      plane[dd][1][c] = h2_lerp4(row[1][c], row[2][c], q);
    }
  }
  const int OH = 2 * H, OW = 2 * W;
  uint4* ob = reinterpret_cast<uint4*>(out) + n * (long long)8 * D * H * W * c8n;
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
      for (int c = 0; c < 2; c++)
        ob[((long long)((2 * d + a) * OH + 2 * h + b) * OW + 2 * w + c) * c8n + c8] =
            h2_lerp4(plane[1][b][c], plane[a == 0 ? 0 : 2][b][c], q);
}

// ------------------------------------------------------------------ layout conversion
// channels-last fp16 (N, S, C) -> NCDHW fp32 (N, C, S): the first_feature output tensor.
__global__ void __launch_bounds__(256)
ndhwc_to_ncdhw_kernel(const act_t* __restrict__ x, float* __restrict__ out, int S, int C, int n_stride_in) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const act_t* src = x + (long long)n * n_stride_in;
  for (int j = ty; j < 32; j += 8) {
    const int s = s0 + j, c = c0 + tx;
    tile[j][tx] = (s < S && c < C) ? __half2float(src[(long long)s * C + c]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, s = s0 + tx;
    if (s < S && c < C) out[((long long)n * C + c) * S + s] = tile[tx][j];
  }
}

// NCDHW fp32 -> channels-last fp16 (decode_from_dyna receives first_feature from the caller)
__global__ void __launch_bounds__(256)
ncdhw_to_ndhwc_kernel(const float* __restrict__ x, act_t* __restrict__ out, int S, int C) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, s = s0 + tx;
    tile[j][tx] = (s < S && c < C) ? x[((long long)n * C + c) * S + s] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int s = s0 + j, c = c0 + tx;
    if (s < S && c < C) out[((long long)n * S + s) * C + c] = __float2half_rn(tile[tx][j]);
  }
}

// ------------------------------------------------------------------ per-clip frame mean
// seq.mean(dim=1) at model/kypt_detector.py:312: (B, T, S) fp32 -> (B, S) fp32.
__global__ void __launch_bounds__(256)
mean_over_t_kernel(const float* __restrict__ seq, float* __restrict__ out, int T, long long S4) {
  const int b = blockIdx.y;
  const float4* src = reinterpret_cast<const float4*>(seq) + (long long)b * T * S4;
  float4* dst = reinterpret_cast<float4*>(out) + (long long)b * S4;
  const float inv = 1.0f / (float)T;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < S4; i += (long long)gridDim.x * 256) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; t++) {
      const float4 v = src[(long long)t * S4 + i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    dst[i] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}

// ------------------------------------------------------------------ final 1x1 conv + activation (+BCE)
// Reference: decoder tail GN -> LReLU -> Conv3d(32,1,k1) (kypt_detector.py:453-457) then
// sigmoid(sharpness * (tanh(x) + first_frame - translation)) (:410) and
// nn.BCELoss(reduction='none')(recon, seq).mean (:91-92; log clamped at -100).
// One thread per voxel: reads C fp16 channels, writes one fp32.
template <int C>
__global__ void __launch_bounds__(256)
final_recon_kernel(const act_t* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                   const float* __restrict__ w, float bias, const float* __restrict__ bias_dev,
                   const float* __restrict__ first_frame,
                   int frames_per_clip, float sharp, float trans, float* __restrict__ recon,
                   const float* __restrict__ target, float* __restrict__ bce_partial, int S) {
  if (bias_dev) bias = __ldg(bias_dev);            // the live parameter: no host copy (and no host sync) per training step
  // C/8 lanes per voxel: each loads one 16-byte channel chunk (consecutive lanes -> consecutive addresses),
  // partial dot product, shuffle-reduce inside the lane group
  constexpr int LPV = C / 8;                       // lanes per voxel (4 for C = 32)
  const int n = blockIdx.y;
  __shared__ float sa[C], sb[C], sw[C];
  if (threadIdx.x < C) {
    sa[threadIdx.x] = a[(long long)n * C + threadIdx.x];
    sb[threadIdx.x] = b[(long long)n * C + threadIdx.x];
    sw[threadIdx.x] = w[threadIdx.x];
  }
  __syncthreads();
  const int clip = n / frames_per_clip;
  const int sub = threadIdx.x % LPV;
  float loss = 0.f;
  // this lane's 8 channels never change: scale / shift / weight live in registers (24 shared loads per chunk less)
  float ra[8], rb[8], rw[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { ra[k] = sa[sub * 8 + k]; rb[k] = sb[sub * 8 + k]; rw[k] = sw[sub * 8 + k]; }
  const half8* base = reinterpret_cast<const half8*>(x + (long long)n * S * C);
  // A group of LPV lanes handles LPV consecutive voxels per iteration: lane `sub` loads channel chunk `sub` of
  // each of them, the partial dot products are transpose-reduced so that lane j ends up with the full sum of
  // voxel j, and every lane then runs the tanh / sigmoid / log tail for ONE voxel (no divergent scalar tail).
  const int groups_per_block = 256 / LPV;
  for (int v0 = (blockIdx.x * groups_per_block + threadIdx.x / LPV) * LPV; v0 < S; v0 += gridDim.x * groups_per_block * LPV) {
    float part[LPV];
    half8 raw[LPV];
#pragma unroll
    for (int j = 0; j < LPV; j++) raw[j] = base[(long long)(v0 + j) * LPV + sub];   // all loads in flight first
    const float ff = first_frame[(long long)clip * S + v0 + sub];
    const float tt = target ? target[(long long)n * S + v0 + sub] : 0.f;
#pragma unroll
    for (int j = 0; j < LPV; j++) {
      float f[8];
      nm_unpack8(raw[j], f);
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; k++) acc = fmaf(nm_lrelu(fmaf(f[k], ra[k], rb[k])), rw[k], acc);
      part[j] = acc;
    }
    // recursive halving inside the LPV-lane group: after it, lane `sub` holds the total of voxel `sub`
#pragma unroll
    for (int half = LPV / 2; half >= 1; half >>= 1) {
      const bool upper = (sub & half) != 0;
#pragma unroll
      for (int i = 0; i < half; i++) {
        const float send = upper ? part[i] : part[i + half];
        const float recv = __shfl_xor_sync(0xffffffffu, send, half);
        part[i] = (upper ? part[i + half] : part[i]) + recv;
      }
    }
    const int s = v0 + sub;
    const float z = sharp * (tanhf(part[0] + bias) + ff - trans);
    const float r = 1.0f / (1.0f + expf(-z));
    recon[(long long)n * S + s] = r;
    if (target) {
      const float t = tt;
      loss -= t * fmaxf(logf(r), -100.f) + (1.f - t) * fmaxf(logf(1.f - r), -100.f);
    }
  }
  if (bce_partial) {
    __shared__ float red[8];
    loss = nm_warp_sum(loss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int k = 0; k < 8; k++) tot += red[k];
      bce_partial[(long long)n * gridDim.x + blockIdx.x] = tot;
    }
  }
}

__global__ void reduce_rows_kernel(const float* __restrict__ partial, int cols, float scale, float* __restrict__ out) {
  // one warp per row, fixed order -> deterministic
  const int row = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < cols; i += 32) acc += (double)partial[(long long)row * cols + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) out[row] = (float)(acc * (double)scale);
}

// ------------------------------------------------------------------ chamfer volume-fitting loss
// Reference: get_volume_fitting_loss('chamfer'), utils/kypt_detector_utils.py:141-157.  The reference
// materialises (B,K,3,X^3); here: per frame, over occupied voxels only, min_k |coord - kp_k|^2.
__global__ void __launch_bounds__(256)
chamfer_kernel(const float* __restrict__ seq, const float* __restrict__ kp, int K, int G,
               const float* __restrict__ lin, float* __restrict__ partial /* [n][blocks][2] */) {
  const int n = blockIdx.y;
  extern __shared__ float s_kp[];  // K*3
  for (int i = threadIdx.x; i < K * 3; i += 256) s_kp[i] = kp[((long long)n * K + i / 3) * 4 + i % 3];
  __syncthreads();
  const int S = G * G * G;
  float dsum = 0.f, osum = 0.f;
  for (int s = blockIdx.x * 256 + threadIdx.x; s < S; s += gridDim.x * 256) {
    const float o = seq[(long long)n * S + s];
    if (o != 0.f) {
      const float cx = lin[s / (G * G)], cy = lin[(s / G) % G], cz = lin[s % G];
      float best = INFINITY;
      for (int k = 0; k < K; k++) {
        const float dx = cx - s_kp[k * 3], dy = cy - s_kp[k * 3 + 1], dz = cz - s_kp[k * 3 + 2];
        best = fminf(best, dx * dx + dy * dy + dz * dz);
      }
      dsum += best * o;
      osum += o;
    }
  }
  __shared__ float red[8][2];
  dsum = nm_warp_sum(dsum);
  osum = nm_warp_sum(osum);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = dsum; red[threadIdx.x >> 5][1] = osum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < 8; k++) { a += red[k][0]; b += red[k][1]; }
    partial[((long long)n * gridDim.x + blockIdx.x) * 2] = a;
    partial[((long long)n * gridDim.x + blockIdx.x) * 2 + 1] = b;
  }
}

__global__ void chamfer_finalize_kernel(const float* __restrict__ partial, int blocks, float* __restrict__ out) {
  const int n = blockIdx.x;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < blocks; i += 32) {
    a += (double)partial[((long long)n * blocks + i) * 2];
    b += (double)partial[((long long)n * blocks + i) * 2 + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (threadIdx.x == 0) out[n] = (float)(a / b);
}

}  // namespace

// ============================================================================ C ABI
extern "C" int nm_gn_stats_chunks(int S) {
  int c = S / 2048;
  return c < 1 ? 1 : (c > 128 ? 128 : c);
}

extern "C" size_t nm_gn_workspace_bytes(int n, int S, int C) {
  return (size_t)n * nm_gn_stats_chunks(S) * C * 2 * sizeof(float);
}

extern "C" int nm_groupnorm_scale_shift(const void* x, int n, int S, int C, int groups, const float* gamma,
                                        const float* beta, float eps, float* scale, float* shift,
                                        void* workspace, float* mean_rstd, float* xsum, void* stream) {
  NM_CHECK_ARG(x && gamma && beta && scale && shift && workspace, "nm_groupnorm_scale_shift: null pointer");
  NM_CHECK_ARG(C % 8 == 0 && kStatThreads % (C / 8) == 0 && groups > 0 && groups <= 32 && C % groups == 0,
               "nm_groupnorm_scale_shift: unsupported C=%d groups=%d", C, groups);
  if (n == 0) return NM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = nm_gn_stats_chunks(S);
  gn_stats_kernel<<<dim3(chunks, n), kStatThreads, 0, st>>>((const act_t*)x, S, C, chunks, (float*)workspace);
  NM_CHECK_LAUNCH("gn_stats");
  gn_finalize_kernel<<<dim3(n, groups), 256, 0, st>>>((const float*)workspace, S, C, groups, chunks, gamma, beta, eps, scale, shift,
                                                      mean_rstd, xsum);
  NM_CHECK_LAUNCH("gn_finalize");
  return NM_OK;
}

// Finalize GroupNorm scale/shift from partial sums produced elsewhere (the conv epilogues emit them):
// partial [n][chunks][C][2] = (sum, sum of squares) over disjoint pieces of each sample.
extern "C" int nm_groupnorm_finalize(const float* partial, int n, int S, int C, int groups, int chunks,
                                     const float* gamma, const float* beta, float eps, float* scale, float* shift,
                                     float* mean_rstd, float* xsum, void* stream) {
  NM_CHECK_ARG(partial && gamma && beta && scale && shift, "nm_groupnorm_finalize: null pointer");
  NM_CHECK_ARG(groups > 0 && groups <= 32 && C % groups == 0 && chunks > 0, "nm_groupnorm_finalize: bad C=%d groups=%d chunks=%d",
               C, groups, chunks);
  if (n == 0) return NM_OK;
  gn_finalize_kernel<<<dim3(n, groups), 256, 0, (cudaStream_t)stream>>>(partial, S, C, groups, chunks, gamma, beta, eps, scale, shift,
                                                                        mean_rstd, xsum);
  NM_CHECK_LAUNCH("gn_finalize");
  return NM_OK;
}

extern "C" int nm_affine_act(const void* x1, const float* a1, const float* b1, int act1, const void* x2,
                             const float* a2, const float* b2, void* out, int n, int S, int C, void* stream) {
  NM_CHECK_ARG(x1 && a1 && b1 && out, "nm_affine_act: null pointer");
  NM_CHECK_ARG(C % 8 == 0, "nm_affine_act: C=%d not a multiple of 8", C);
  const long long total8 = (long long)n * S * (C / 8);
  if (total8 == 0) return NM_OK;
  const int c8n = C / 8;
  if (256 % c8n == 0 && n <= 65535) {
    const int nvl = 256 / c8n;
    long long chunks = (16LL * nm_num_sms() + n - 1) / n;
    const long long rounds = ((long long)S + nvl - 1) / nvl;
    if (chunks > rounds) chunks = rounds;
    if (chunks < 1) chunks = 1;
    const dim3 grid((unsigned)chunks, (unsigned)n);
    cudaStream_t st = (cudaStream_t)stream;
    if (!x2)
      affine_act_rows_kernel<false, false><<<grid, 256, 0, st>>>((const act_t*)x1, a1, b1, act1, nullptr, nullptr, nullptr, (act_t*)out, S, C);
    else if (a2)
      affine_act_rows_kernel<true, true><<<grid, 256, 0, st>>>((const act_t*)x1, a1, b1, act1, (const act_t*)x2, a2, b2, (act_t*)out, S, C);
    else
      affine_act_rows_kernel<true, false><<<grid, 256, 0, st>>>((const act_t*)x1, a1, b1, act1, (const act_t*)x2, nullptr, nullptr, (act_t*)out, S, C);
    NM_CHECK_LAUNCH("affine_act(rows)");
    return NM_OK;
  }
  const int blocks = (int)min((long long)nm_num_sms() * 16, (total8 + 255) / 256);
  affine_act_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const act_t*)x1, a1, b1, act1, (const act_t*)x2, a2,
                                                              b2, (act_t*)out, S, C, total8);
  NM_CHECK_LAUNCH("affine_act");
  return NM_OK;
}

extern "C" int nm_upsample2x(const void* x, const float* a, const float* b, int act, void* out, int n, int D, int H,
                             int W, int C, void* stream) {
  NM_CHECK_ARG(x && out, "nm_upsample2x: null pointer");
  NM_CHECK_ARG(C % 8 == 0, "nm_upsample2x: C=%d not a multiple of 8", C);
  const long long total8 = (long long)n * D * H * W * (C / 8);      // one thread per input cell x 8 channels
  if (total8 == 0) return NM_OK;
  const int blocks = (int)((total8 + 127) / 128);
  if (a)
    upsample2x_kernel<true><<<blocks, 128, 0, (cudaStream_t)stream>>>((const act_t*)x, a, b, act, (act_t*)out, D, H,
                                                                      W, C, total8);
  else
    upsample2x_h2_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>((const act_t*)x, (act_t*)out, D, H, W, C, total8);
  NM_CHECK_LAUNCH("upsample2x");
  return NM_OK;
}

extern "C" int nm_ndhwc_to_ncdhw_f32(const void* x, float* out, int n, int S, int C, long long in_sample_stride,
                                     void* stream) {
  NM_CHECK_ARG(x && out, "nm_ndhwc_to_ncdhw_f32: null pointer");
  if (n == 0) return NM_OK;
  ndhwc_to_ncdhw_kernel<<<dim3(nm_cdiv(S, 32), nm_cdiv(C, 32), n), 256, 0, (cudaStream_t)stream>>>(
      (const act_t*)x, out, S, C, (int)in_sample_stride);
  NM_CHECK_LAUNCH("ndhwc_to_ncdhw");
  return NM_OK;
}

extern "C" int nm_ncdhw_f32_to_ndhwc(const float* x, void* out, int n, int S, int C, void* stream) {
  NM_CHECK_ARG(x && out, "nm_ncdhw_f32_to_ndhwc: null pointer");
  if (n == 0) return NM_OK;
  ncdhw_to_ndhwc_kernel<<<dim3(nm_cdiv(S, 32), nm_cdiv(C, 32), n), 256, 0, (cudaStream_t)stream>>>(x, (act_t*)out,
                                                                                                 S, C);
  NM_CHECK_LAUNCH("ncdhw_to_ndhwc");
  return NM_OK;
}

extern "C" int nm_mean_over_frames(const float* seq, float* out, int n_clips, int T, long long S, void* stream) {
  NM_CHECK_ARG(seq && out, "nm_mean_over_frames: null pointer");
  NM_CHECK_ARG(S % 4 == 0, "nm_mean_over_frames: S must be a multiple of 4");
  if (n_clips == 0) return NM_OK;
  const long long S4 = S / 4;
  mean_over_t_kernel<<<dim3((int)min((S4 + 255) / 256, 1024LL), n_clips), 256, 0, (cudaStream_t)stream>>>(seq, out,
                                                                                                       T, S4);
  NM_CHECK_LAUNCH("mean_over_frames");
  return NM_OK;
}

constexpr int kReconBlocks = 64;
extern "C" size_t nm_final_recon_workspace_bytes(int n) { return (size_t)n * kReconBlocks * sizeof(float); }

extern "C" int nm_final_recon(const void* x, const float* a, const float* b, const float* w, float bias, const float* bias_dev,
                              const float* first_frame, int frames_per_clip, float sharpness, float translation,
                              float* recon, const float* target, float* bce_mean, void* workspace, int n, int S,
                              int C, void* stream) {
  NM_CHECK_ARG(x && a && b && w && first_frame && recon, "nm_final_recon: null pointer");
  NM_CHECK_ARG(C == 32, "nm_final_recon: C=%d unsupported (decoder tail is 32 channels)", C);
  NM_CHECK_ARG(!target || (bce_mean && workspace), "nm_final_recon: BCE needs bce_mean and workspace");
  if (n == 0) return NM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  final_recon_kernel<32><<<dim3(kReconBlocks, n), 256, 0, st>>>((const act_t*)x, a, b, w, bias, bias_dev, first_frame,
                                                               frames_per_clip, sharpness, translation, recon,
                                                               target, target ? (float*)workspace : nullptr, S);
  NM_CHECK_LAUNCH("final_recon");
  if (target) {
    reduce_rows_kernel<<<n, 32, 0, st>>>((const float*)workspace, kReconBlocks, 1.0f / (float)S, bce_mean);
    NM_CHECK_LAUNCH("final_recon(reduce)");
  }
  return NM_OK;
}

constexpr int kChamferBlocks = 32;
extern "C" size_t nm_chamfer_workspace_bytes(int n) { return (size_t)n * kChamferBlocks * 2 * sizeof(float); }

extern "C" int nm_chamfer_vol_fit(const float* seq, const float* keypoints, const float* linspace, int n, int K,
                                  int G, float* out, void* workspace, void* stream) {
  NM_CHECK_ARG(seq && keypoints && linspace && out && workspace, "nm_chamfer_vol_fit: null pointer");
  if (n == 0) return NM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  chamfer_kernel<<<dim3(kChamferBlocks, n), 256, K * 3 * sizeof(float), st>>>(seq, keypoints, K, G, linspace,
                                                                              (float*)workspace);
  NM_CHECK_LAUNCH("chamfer");
  chamfer_finalize_kernel<<<n, 32, 0, st>>>((const float*)workspace, kChamferBlocks, out);
  NM_CHECK_LAUNCH("chamfer(finalize)");
  return NM_OK;
}
