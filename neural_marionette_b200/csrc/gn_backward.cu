// Backward of z = LeakyReLU_0.01(GroupNorm(x)) (or of GroupNorm alone) on channels-last fp16 tensors - config #4 backward,
// first version: three HBM passes, every reduction in a fixed order (fp32 partials per (sample, chunk,
// channel), fp64 finalize) -> bit-reproducible.
//   pass A: per (sample, channel) sum x, sum x^2                 -> mean, rstd per (sample, group)
//   pass B: dy = dz * (y > 0 ? 1 : 0.01), y = gamma * xhat + beta -> per (sample, channel) sum dy, sum dy * x
//           -> dbeta, dgamma (summed over samples in order) and the two group means of the input gradient
//   pass C: dx = rstd * (gamma * dy - m1 - xhat * m2)
// Reference semantics: torch.autograd through nn.GroupNorm + nn.LeakyReLU (modules/vox_modules.py:8-75).
#include "common.cuh"
#include "../../include/nm_b200.h"

namespace {

constexpr int kGnbThreads = 256;
constexpr int kGnbChunks = 32;                 // CTAs per sample in the reduction passes

struct GnbShape {
  int n, C, groups, cpg, ccs, nvl;             // ccs = C / 8 channel chunks, nvl = voxel lanes per CTA
  long long S;
};

// Block-level fixed-order reduction of per-thread partials v[16] (8 channels x 2 quantities): thread (vl, cc) holds the
// partial of channel chunk cc over its voxels.  Lanes of a warp that share a chunk (32 / ccs of them; ccs is a power of two
// <= 32) are combined by a butterfly, the 8 warps through shared memory in warp order: a fixed order, bit-reproducible.
// (The first version had 64 threads sum nvl = 64 shared-memory values serially: ~1.2 us of every ~9 us CTA.)
__device__ __forceinline__ void gnb_block_reduce(const float* v, float* red, const GnbShape& s, float* out /* [C][2] */) {
  float w[16];
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = v[i];
  for (int o = s.ccs; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] += __shfl_xor_sync(0xffffffffu, w[i], o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < s.ccs) {                                   // lane == channel chunk (32 % ccs == 0)
#pragma unroll
    for (int i = 0; i < 16; i++) red[(warp * s.ccs + lane) * 16 + i] = w[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < s.ccs * 16; i += kGnbThreads) {
    const int c2 = i / 16, q = i % 16;         // q = channel-in-chunk * 2 + quantity
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kGnbThreads / 32; k++) acc += red[(k * s.ccs + c2) * 16 + q];
    out[(c2 * 8 + (q >> 1)) * 2 + (q & 1)] = acc;
  }
}

// pass A (stats == true): partial[n][chunk][c] = (sum x, sum x^2)
// pass B (stats == false): partial[n][chunk][c] = (sum dy, sum dy * x), dy masked by the LeakyReLU derivative
template <bool kStats>
__global__ void __launch_bounds__(kGnbThreads)
gnb_reduce_kernel(const __half* __restrict__ x, const __half* __restrict__ dz, const float* __restrict__ gamma,
                  const float* __restrict__ beta, const float* __restrict__ mean_rstd, GnbShape s, int leaky,
                  float* __restrict__ partial) {
  __shared__ float red[kGnbThreads * 16];
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int cc = threadIdx.x % s.ccs, vl = threadIdx.x / s.ccs;
  const long long per = (s.S + kGnbChunks - 1) / kGnbChunks;
  const long long v0 = chunk * per, v1 = min(s.S, v0 + per);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = 0.f;
  float g[8], b[8], mu = 0.f, rs = 0.f;
  if (!kStats) {
    const int grp = (cc * 8) / s.cpg;          // cpg is a multiple of 8 (checked by the host): one group per chunk
    mu = mean_rstd[(n * s.groups + grp) * 2];
    rs = mean_rstd[(n * s.groups + grp) * 2 + 1];
#pragma unroll
    for (int i = 0; i < 8; i++) { g[i] = gamma[cc * 8 + i]; b[i] = beta[cc * 8 + i]; }
  }
#pragma unroll 2
  for (long long v = v0 + vl; v < v1; v += s.nvl) {
    const long long off = ((long long)n * s.S + v) * s.C + cc * 8;
    float xf[8];
    nm_unpack8(*reinterpret_cast<const half8*>(x + off), xf);
    if (kStats) {
#pragma unroll
      for (int i = 0; i < 8; i++) { acc[2 * i] += xf[i]; acc[2 * i + 1] = fmaf(xf[i], xf[i], acc[2 * i + 1]); }
    } else {
      float df[8];
      nm_unpack8(*reinterpret_cast<const half8*>(dz + off), df);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float y = fmaf(g[i], (xf[i] - mu) * rs, b[i]);
        const float dy = (leaky && !(y > 0.f)) ? 0.01f * df[i] : df[i];
        acc[2 * i] += dy;
        acc[2 * i + 1] = fmaf(dy, xf[i], acc[2 * i + 1]);
      }
    }
  }
  gnb_block_reduce(acc, red, s, partial + ((long long)n * kGnbChunks + chunk) * s.C * 2);
}

// mean / rstd per (sample, group) from the pass-A partials (fp64, fixed order)
__global__ void gnb_stats_finalize_kernel(const float* __restrict__ partial, GnbShape s, float eps, float* __restrict__ mean_rstd,
                                          float* __restrict__ xsum /* [n][C] */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.n * s.groups) return;
  const int n = i / s.groups, grp = i % s.groups;
  double s1 = 0.0, s2 = 0.0;
  for (int c = grp * s.cpg; c < (grp + 1) * s.cpg; c++) {
    double c1 = 0.0;
    for (int ch = 0; ch < kGnbChunks; ch++) {
      const float* p = partial + (((long long)n * kGnbChunks + ch) * s.C + c) * 2;
      c1 += p[0];
      s2 += p[1];
    }
    xsum[(long long)n * s.C + c] = (float)c1;
    s1 += c1;
  }
  const double M = (double)s.cpg * (double)s.S;
  const double mu = s1 / M;
  double var = s2 / M - mu * mu;                                        // biased variance, as nn.GroupNorm
  if (var < 0.0) var = 0.0;
  mean_rstd[i * 2] = (float)mu;
  mean_rstd[i * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// per (sample, channel) totals of pass B -> group means m1, m2 (per sample, group) and dgamma / dbeta contributions.
// One warp per (sample, group): lane = (chunk half, channel of the group) sums its 16 chunk partials in order, the two
// halves and then the channels are combined with butterflies (fixed order; the serial one-thread-per-group version
// took 73 us per launch, 3.5 ms per training step).
__device__ __forceinline__ double gnb_shfl_xor(double v, int m) {
  return __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(v), m), __shfl_xor_sync(0xffffffffu, __double2loint(v), m));
}

__global__ void __launch_bounds__(128)
gnb_grad_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ gamma,
                         const float* __restrict__ mean_rstd, const float* __restrict__ xsum, GnbShape s,
                         float* __restrict__ m12, float* __restrict__ chan /* [n][C][3]: sum dy, sum dy*xhat, sum dx */) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= s.n * s.groups) return;                                       // warp-uniform
  const int n = i / s.groups, grp = i % s.groups;
  const double mu = mean_rstd[i * 2], rs = mean_rstd[i * 2 + 1];
  const double M = (double)s.cpg * (double)s.S;
  double a1 = 0.0, a2 = 0.0;
  // channels of the group in rounds of 16 (cpg is a multiple of 8: 8, 16 or 32 here)
  for (int c0 = 0; c0 < s.cpg; c0 += 16) {
    const int cl = c0 + (lane & 15), half = lane >> 4;
    const bool live = cl < s.cpg;
    const int c = grp * s.cpg + (live ? cl : 0);
    double p1 = 0.0, p2 = 0.0;
    if (live)
      for (int ch = half * (kGnbChunks / 2); ch < (half + 1) * (kGnbChunks / 2); ch++) {
        const float* p = partial + (((long long)n * kGnbChunks + ch) * s.C + c) * 2;
        p1 += p[0];
        p2 += p[1];
      }
    p1 += gnb_shfl_xor(p1, 16);
    p2 += gnb_shfl_xor(p2, 16);
    const double dyxhat = rs * (p2 - mu * p1);                          // sum dy * xhat
    double g1 = live ? (double)gamma[c] * p1 : 0.0, g2 = live ? (double)gamma[c] * dyxhat : 0.0;
    if (live && half == 0) {
      chan[((long long)n * s.C + c) * 3] = (float)p1;
      chan[((long long)n * s.C + c) * 3 + 1] = (float)dyxhat;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      g1 += gnb_shfl_xor(g1, o);
      g2 += gnb_shfl_xor(g2, o);
    }
    a1 += g1;
    a2 += g2;
  }
  const double m1 = a1 / M, m2 = a2 / M;
  if (lane == 0) {
    m12[i * 2] = (float)m1;
    m12[i * 2 + 1] = (float)m2;
  }
  __syncwarp();
  // sum over the voxels of dx = rstd * (gamma * dy - m1 - xhat * m2), in closed form from the channel totals
  for (int cl = lane; cl < s.cpg; cl += 32) {
    const int c = grp * s.cpg + cl;
    const double p1 = chan[((long long)n * s.C + c) * 3];
    const double xhat_sum = rs * ((double)xsum[(long long)n * s.C + c] - (double)s.S * mu);
    chan[((long long)n * s.C + c) * 3 + 2] = (float)(rs * ((double)gamma[c] * p1 - (double)s.S * m1 - xhat_sum * m2));
  }
}

// One warp per channel: lane l sums samples l, l + 32, ... in order, then a butterfly (fixed order, fp64).  The
// one-thread-per-channel version walked the samples serially: 23 us per launch, 76 launches per training step.
__global__ void __launch_bounds__(128)
gnb_param_grad_kernel(const float* __restrict__ chan, int n_samples, int C, float out_scale,
                      float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;                                                     // warp-uniform
  double db = 0.0, dg = 0.0, dx = 0.0;
  for (int n = lane; n < n_samples; n += 32) {
    const float* p = chan + ((long long)n * C + c) * 3;
    db += p[0];
    dg += p[1];
    dx += p[2];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    db += gnb_shfl_xor(db, o);
    dg += gnb_shfl_xor(dg, o);
    dx += gnb_shfl_xor(dx, o);
  }
  if (lane == 0) {
    if (dbeta) dbeta[c] = (float)(db * (double)out_scale);
    if (dgamma) dgamma[c] = (float)(dg * (double)out_scale);
    if (dxsum) dxsum[c] = (float)(dx * (double)out_scale);
  }
}

// Any (C, groups) on small tensors (hour-glass levels with 48 / 72 channels, 18 channels per group): one warp per
// (sample, group), three sweeps over the group's S * cpg values, butterfly reductions (fixed order).
__global__ void __launch_bounds__(32)
gnb_small_kernel(const __half* __restrict__ x, const __half* __restrict__ dz, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int C, int groups, long long S, float eps, int leaky,
                 __half* __restrict__ dx, float* __restrict__ chan /* [n][C][3] */) {
  const int n = blockIdx.y, grp = blockIdx.x, lane = threadIdx.x;
  const int cpg = C / groups, c0 = grp * cpg;
  const long long M = (long long)cpg * S;
  const __half* xs = x + (long long)n * S * C;
  const __half* ds = dz + (long long)n * S * C;
  float s1 = 0.f, s2 = 0.f;
  for (long long i = lane; i < M; i += 32) {
    const float v = __half2float(xs[(i / cpg) * C + c0 + (int)(i % cpg)]);
    s1 += v;
    s2 = fmaf(v, v, s2);
  }
  s1 = nm_warp_sum(s1);
  s2 = nm_warp_sum(s2);
  const float mu = s1 / (float)M;
  const float rs = rsqrtf(fmaxf(s2 / (float)M - mu * mu, 0.f) + eps);
  float a1 = 0.f, a2 = 0.f;
  for (int c = c0; c < c0 + cpg; c++) {
    const float g = gamma[c], b = beta[c];
    float p1 = 0.f, p2 = 0.f, xh = 0.f;
    for (long long v = lane; v < S; v += 32) {
      const float xhat = (__half2float(xs[v * C + c]) - mu) * rs;
      const float y = fmaf(g, xhat, b);
      const float d = __half2float(ds[v * C + c]);
      const float dy = (leaky && !(y > 0.f)) ? 0.01f * d : d;
      p1 += dy;
      p2 = fmaf(dy, xhat, p2);
      xh += xhat;
    }
    p1 = nm_warp_sum(p1);
    p2 = nm_warp_sum(p2);
    xh = nm_warp_sum(xh);
    if (lane == 0) {
      chan[((long long)n * C + c) * 3] = p1;
      chan[((long long)n * C + c) * 3 + 1] = p2;
      chan[((long long)n * C + c) * 3 + 2] = xh;             // completed below once m1, m2 are known
    }
    a1 = fmaf(g, p1, a1);
    a2 = fmaf(g, p2, a2);
  }
  const float m1 = a1 / (float)M, m2 = a2 / (float)M;
  __syncwarp();
  for (int c = c0 + lane; c < c0 + cpg; c += 32) {
    float* q = chan + ((long long)n * C + c) * 3;
    q[2] = rs * (gamma[c] * q[0] - (float)S * m1 - q[2] * m2);
  }
  for (long long i = lane; i < M; i += 32) {
    const long long off = (i / cpg) * C + c0 + (int)(i % cpg);
    const int c = c0 + (int)(i % cpg);
    const float xhat = (__half2float(xs[off]) - mu) * rs;
    const float y = fmaf(gamma[c], xhat, beta[c]);
    const float d = __half2float(ds[off]);
    const float dy = (leaky && !(y > 0.f)) ? 0.01f * d : d;
    dx[(long long)n * S * C + off] = __float2half_rn(rs * (gamma[c] * dy - m1 - xhat * m2));
  }
}

// pass C: dx = rstd * (gamma * dy - m1 - xhat * m2)
__global__ void __launch_bounds__(kGnbThreads)
gnb_dx_kernel(const __half* __restrict__ x, const __half* __restrict__ dz, const float* __restrict__ gamma, const float* __restrict__ beta,
              const float* __restrict__ mean_rstd, const float* __restrict__ m12, GnbShape s, int leaky, __half* __restrict__ dx) {
  // grid (chunks, n); thread = (voxel lane, 16-byte channel chunk): the chunk's parameters live in registers and the voxel
  // loop has no index arithmetic beyond one add (the flat-index version spent two 64-bit divisions per 16 bytes and was
  // instruction-bound at 4.2 TB/s)
  const int n = blockIdx.y;
  const int cc = threadIdx.x % s.ccs, vl = threadIdx.x / s.ccs;
  const int grp = (cc * 8) / s.cpg;
  const float mu = mean_rstd[(n * s.groups + grp) * 2], rs = mean_rstd[(n * s.groups + grp) * 2 + 1];
  const float m1 = m12[(n * s.groups + grp) * 2], m2 = m12[(n * s.groups + grp) * 2 + 1];
  float g[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { g[k] = gamma[cc * 8 + k]; b[k] = beta[cc * 8 + k]; }
  const long long base = (long long)n * s.S * s.C + cc * 8;
  const long long step = (long long)gridDim.x * s.nvl;
#pragma unroll 2
  for (long long v = (long long)blockIdx.x * s.nvl + vl; v < s.S; v += step) {
    const long long off = base + v * s.C;
    float xf[8], df[8], o[8];
    nm_unpack8(*reinterpret_cast<const half8*>(x + off), xf);
    nm_unpack8(*reinterpret_cast<const half8*>(dz + off), df);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float xhat = (xf[k] - mu) * rs;
      const float y = fmaf(g[k], xhat, b[k]);
      const float dy = (leaky && !(y > 0.f)) ? 0.01f * df[k] : df[k];
      o[k] = rs * (g[k] * dy - m1 - xhat * m2);
    }
    *reinterpret_cast<half8*>(dx + off) = nm_pack8(o);
  }
}

// ------------------------------------------------------------------ decoder tail fused with the last GroupNorm backward
// The tail's gradient w.r.t. the activated tensor is rank one per voxel: dz[v][c] = w[c] * dx14[v].  Instead of writing
// it (16.8 MB / frame) and re-reading it twice, pass 1 recomputes the tail, stores the per-voxel scalar dx14 and already
// accumulates the GroupNorm-backward sums (sum dy, sum dy * x with dy = dz * LeakyReLU'), pass 2 forms dz on the fly.
//   forward: y = x*a + b (GroupNorm folded), act = lrelu(y), x14 = w . act + bias, p = sigmoid(sharp (tanh(x14) + ff - trans))
//   dx14 = gbce[n] / S * (p - t) * [p(1-p) / max(p(1-p), 1e-12)] * sharp * (1 - tanh^2(x14))       (model/kypt_detector.py:410,91-92)
//   dw[c] = sum_v act[c] dx14 = a[c] * sum(dx14 m x) + b[c] * sum(dx14 m),  m = LeakyReLU'(y)
template <int C>
__global__ void __launch_bounds__(256)
tail_bwd_pass1_kernel(const __half* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                      const float* __restrict__ w, float bias, const float* __restrict__ bias_dev, float sharp,
                      const float* __restrict__ recon,
                      const float* __restrict__ target, const float* __restrict__ gbce, float grad_scale, long long S,
                      float* __restrict__ dx14_out /* (n, S), times grad_scale */,
                      float* __restrict__ gn_partial /* [n][kGnbChunks][C][2] */, float* __restrict__ tail_partial /* [n][chunks][C + 1] */) {
  static_assert(C == 32, "decoder tail has 32 channels");
  if (bias_dev) bias = __ldg(bias_dev);
  const int n = blockIdx.y;
  __shared__ float sa[C], sb[C], sw[C];
  __shared__ float red[8][2 * C + 1];
  if (threadIdx.x < C) {
    sa[threadIdx.x] = a[(long long)n * C + threadIdx.x];
    sb[threadIdx.x] = b[(long long)n * C + threadIdx.x];
    sw[threadIdx.x] = w[threadIdx.x];
  }
  __syncthreads();
  const float gn = gbce[n] / (float)S;
  float A[C], B[C], db = 0.f;
#pragma unroll
  for (int c = 0; c < C; c++) A[c] = B[c] = 0.f;
  const half8* base = reinterpret_cast<const half8*>(x + (long long)n * S * C);
  // software pipeline: the next voxel's 64 bytes (and its recon / target) are requested before this voxel's ~250
  // instructions of arithmetic (16 warps per SM cannot hide the load latency otherwise: 2.2 TB/s)
  const long long sstep = (long long)gridDim.x * 256;
  long long s = (long long)blockIdx.x * 256 + threadIdx.x;
  half8 nraw[4];
  float np = 0.f, ntg = 0.f;
  if (s < S) {
#pragma unroll
    for (int j = 0; j < 4; j++) nraw[j] = base[s * 4 + j];
    np = recon[(long long)n * S + s];
    ntg = target[(long long)n * S + s];
  }
  for (; s < S; s += sstep) {
    half8 raw[4];
#pragma unroll
    for (int j = 0; j < 4; j++) raw[j] = nraw[j];
    const float p = np, tg = ntg;
    if (s + sstep < S) {
#pragma unroll
      for (int j = 0; j < 4; j++) nraw[j] = base[(s + sstep) * 4 + j];
      np = recon[(long long)n * S + s + sstep];
      ntg = target[(long long)n * S + s + sstep];
    }
    float x14 = bias;
    uint32_t pos = 0;                                  // bit c: y_c > 0
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float f[8];
      nm_unpack8(raw[j], f);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int c = j * 8 + k;
        const float y = fmaf(f[k], sa[c], sb[c]);
        pos |= (y > 0.f ? 1u : 0u) << c;
        x14 = fmaf(y > 0.f ? y : 0.01f * y, sw[c], x14);
      }
    }
    const float t = tanhf(x14);
    const float pq = p * (1.f - p);
    const float dx14 = gn * (p - tg) * (pq / fmaxf(pq, 1e-12f)) * sharp * (1.f - t * t);
    db += dx14;
    dx14_out[(long long)n * S + s] = dx14 * grad_scale;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float f[8];
      nm_unpack8(raw[j], f);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int c = j * 8 + k;
        const float dm = ((pos >> c) & 1u) ? dx14 : 0.01f * dx14;
        A[c] += dm;
        B[c] = fmaf(dm, f[k], B[c]);
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < C; c++) {
    const float va = nm_warp_sum(A[c]), vb = nm_warp_sum(B[c]);
    if (lane == 0) { red[warp][c] = va; red[warp][C + c] = vb; }
  }
  db = nm_warp_sum(db);
  if (lane == 0) red[warp][2 * C] = db;
  __syncthreads();
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    float ta = 0.f, tb = 0.f;
    for (int k = 0; k < 8; k++) { ta += red[k][c]; tb += red[k][C + c]; }
    // GroupNorm-backward partial sums of dy = grad_scale * w[c] * dx14 * m and dy * x
    float* gp = gn_partial + (((long long)n * gridDim.x + blockIdx.x) * C + c) * 2;
    gp[0] = grad_scale * sw[c] * ta;
    gp[1] = grad_scale * sw[c] * tb;
    tail_partial[((long long)n * gridDim.x + blockIdx.x) * (C + 1) + c] = sa[c] * tb + sb[c] * ta;    // dw[c]
  } else if (threadIdx.x == C) {
    float t = 0.f;
    for (int k = 0; k < 8; k++) t += red[k][2 * C];
    tail_partial[((long long)n * gridDim.x + blockIdx.x) * (C + 1) + C] = t;                            // dbias
  }
}

// pass 2: dx = rstd * (gamma * dy - m1 - xhat * m2) with dy = w[c] * dx14[v] * LeakyReLU'(gamma * xhat + beta)
__global__ void __launch_bounds__(kGnbThreads)
gnb_dx_rank1_kernel(const __half* __restrict__ x, const float* __restrict__ dx14, const float* __restrict__ w,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean_rstd,
                    const float* __restrict__ m12, GnbShape s, __half* __restrict__ dx) {
  // grid (chunks, n), thread = (voxel lane, channel chunk) as gnb_dx_kernel
  const int n = blockIdx.y;
  const int cc = threadIdx.x % s.ccs, vl = threadIdx.x / s.ccs;
  const int grp = (cc * 8) / s.cpg;
  const float mu = mean_rstd[(n * s.groups + grp) * 2], rs = mean_rstd[(n * s.groups + grp) * 2 + 1];
  const float m1 = m12[(n * s.groups + grp) * 2], m2 = m12[(n * s.groups + grp) * 2 + 1];
  float g[8], b[8], wk[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { g[k] = gamma[cc * 8 + k]; b[k] = beta[cc * 8 + k]; wk[k] = w[cc * 8 + k]; }
  const long long base = (long long)n * s.S * s.C + cc * 8;
  const long long step = (long long)gridDim.x * s.nvl;
#pragma unroll 2
  for (long long v = (long long)blockIdx.x * s.nvl + vl; v < s.S; v += step) {
    const long long off = base + v * s.C;
    const float d14 = dx14[(long long)n * s.S + v];
    float xf[8], o[8];
    nm_unpack8(*reinterpret_cast<const half8*>(x + off), xf);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float xhat = (xf[k] - mu) * rs;
      const float y = fmaf(g[k], xhat, b[k]);
      const float dy = wk[k] * (y > 0.f ? d14 : 0.01f * d14);
      o[k] = rs * (g[k] * dy - m1 - xhat * m2);
    }
    *reinterpret_cast<half8*>(dx + off) = nm_pack8(o);
  }
}

__global__ void tail_reduce_kernel(const float* __restrict__ partial, long long rows, int C, float* __restrict__ dw, float* __restrict__ dbias) {
  const int j = threadIdx.x;
  if (j > C) return;
  double acc = 0.0;
  for (long long r = 0; r < rows; r++) acc += (double)partial[r * (C + 1) + j];
  if (j < C) dw[j] = (float)acc;
  else dbias[0] = (float)acc;
}

bool gnb_shape(int n, long long S, int C, int groups, GnbShape* s) {
  if (n <= 0 || S <= 0 || C <= 0 || groups <= 0 || C % groups) return false;
  const int cpg = C / groups, ccs = C / 8;
  if (C % 8 || cpg % 8 || ccs > 32 || (kGnbThreads % ccs) != 0) return false;   // C in {8, 16, 32, 64, 128, 256}
  *s = GnbShape{n, C, groups, cpg, ccs, kGnbThreads / ccs, S};
  return true;
}

// grid (voxel chunks, samples) of the dx passes: about 16 CTAs per SM in total, at least one voxel-lane round per CTA
dim3 gnb_dx_grid(const GnbShape& s) {
  const long long rounds = (s.S + s.nvl - 1) / s.nvl;
  long long chunks = (16LL * nm_num_sms() + s.n - 1) / s.n;
  if (chunks > rounds) chunks = rounds;
  if (chunks < 1) chunks = 1;
  return dim3((unsigned)chunks, (unsigned)s.n);
}

size_t gnb_ws_floats(int n, int C, int groups) {
  return (size_t)n * kGnbChunks * C * 2 /* partial */ + (size_t)n * groups * 4 /* mean_rstd, m12 */ +
         (size_t)n * C * 3 /* chan */ + (size_t)n * C /* xsum */;
}

}  // namespace

extern "C" size_t nm_groupnorm_backward_workspace_bytes(int n, int C, int groups) {
  if (n <= 0 || C <= 0 || groups <= 0 || C % groups) return 0;
  return gnb_ws_floats(n, C, groups) * sizeof(float) + 64;
}

extern "C" int nm_groupnorm_backward(const void* x, const void* grad_out, const float* gamma, const float* beta, int n,
                                     long long S, int C, int groups, float eps, int leaky, float out_scale,
                                     const float* fwd_mean_rstd, const float* fwd_xsum, void* grad_in, float* dgamma,
                                     float* dbeta, float* dxsum, void* workspace, void* stream) {
  NM_CHECK_ARG(x && grad_out && gamma && beta && grad_in && workspace, "nm_groupnorm_backward: null pointer");
  NM_CHECK_ARG(n > 0 && S > 0 && C > 0 && groups > 0 && C % groups == 0, "nm_groupnorm_backward: bad shape (C=%d, groups=%d)", C, groups);
  NM_CHECK_ARG(n <= 65535, "nm_groupnorm_backward: at most 65535 samples per call");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(workspace);
  float* mean_rstd = partial + (size_t)n * kGnbChunks * C * 2;
  float* m12 = mean_rstd + (size_t)n * groups * 2;
  float* chan = m12 + (size_t)n * groups * 2;
  float* xsum = chan + (size_t)n * C * 3;
  const __half* xh = reinterpret_cast<const __half*>(x);
  const __half* dz = reinterpret_cast<const __half*>(grad_out);
  GnbShape s;
  if (!gnb_shape(n, S, C, groups, &s)) {
    // shapes outside the streaming kernels (48 / 72 channels): the small-tensor kernel, one warp per (sample, group)
    NM_CHECK_ARG(S * (C / groups) <= (1 << 16), "nm_groupnorm_backward: C=%d (groups=%d) is only supported on small tensors", C, groups);
    gnb_small_kernel<<<dim3(groups, n), 32, 0, st>>>(xh, dz, gamma, beta, C, groups, S, eps, leaky, reinterpret_cast<__half*>(grad_in), chan);
    NM_CHECK_LAUNCH("gnb_small_kernel");
    if (dgamma || dbeta || dxsum) {
      gnb_param_grad_kernel<<<nm_cdiv(C, 4), 128, 0, st>>>(chan, n, C, out_scale, dgamma, dbeta, dxsum);
      NM_CHECK_LAUNCH("gnb_param_grad_kernel");
    }
    return NM_OK;
  }
  const dim3 rgrid(kGnbChunks, n);
  if (fwd_mean_rstd && fwd_xsum) {                     // statistics kept from the forward: no pass over x for them
    mean_rstd = const_cast<float*>(fwd_mean_rstd);
    xsum = const_cast<float*>(fwd_xsum);
  } else {
    gnb_reduce_kernel<true><<<rgrid, kGnbThreads, 0, st>>>(xh, dz, gamma, beta, mean_rstd, s, leaky, partial);
    NM_CHECK_LAUNCH("gnb_reduce_kernel<stats>");
    gnb_stats_finalize_kernel<<<nm_cdiv((long long)n * groups, 128), 128, 0, st>>>(partial, s, eps, mean_rstd, xsum);
    NM_CHECK_LAUNCH("gnb_stats_finalize_kernel");
  }
  gnb_reduce_kernel<false><<<rgrid, kGnbThreads, 0, st>>>(xh, dz, gamma, beta, mean_rstd, s, leaky, partial);
  NM_CHECK_LAUNCH("gnb_reduce_kernel<grad>");
  gnb_grad_finalize_kernel<<<nm_cdiv((long long)n * groups, 4), 128, 0, st>>>(partial, gamma, mean_rstd, xsum, s, m12, chan);
  NM_CHECK_LAUNCH("gnb_grad_finalize_kernel");
  if (dgamma || dbeta || dxsum) {
    gnb_param_grad_kernel<<<nm_cdiv(C, 4), 128, 0, st>>>(chan, n, C, out_scale, dgamma, dbeta, dxsum);
    NM_CHECK_LAUNCH("gnb_param_grad_kernel");
  }
  gnb_dx_kernel<<<gnb_dx_grid(s), kGnbThreads, 0, st>>>(xh, dz, gamma, beta, mean_rstd, m12, s, leaky, reinterpret_cast<__half*>(grad_in));
  NM_CHECK_LAUNCH("gnb_dx_kernel");
  return NM_OK;
}

extern "C" size_t nm_final_recon_backward_fused_workspace_bytes(int n, long long S, int C, int groups) {
  if (n <= 0 || C <= 0 || groups <= 0 || C % groups) return 0;
  return (gnb_ws_floats(n, C, groups) + (size_t)n * S /* dx14 */ + (size_t)n * kGnbChunks * (C + 1) /* tail partials */) * sizeof(float) + 64;
}

extern "C" int nm_final_recon_backward_fused(const void* x, const float* a, const float* b, const float* w, float bias,
                                             const float* bias_dev,
                                             float sharpness, const float* recon, const float* target, const float* grad_bce,
                                             float grad_scale, const float* gamma, const float* beta, const float* mean_rstd,
                                             const float* xsum, int groups, void* grad_x, float* dw, float* dbias, float* dgamma,
                                             float* dbeta, float* dxsum, void* workspace, int n, long long S, int C, void* stream) {
  NM_CHECK_ARG(x && a && b && w && recon && target && grad_bce && gamma && beta && mean_rstd && xsum && grad_x && dw && dbias &&
               workspace, "nm_final_recon_backward_fused: null pointer");
  NM_CHECK_ARG(C == 32, "nm_final_recon_backward_fused: C=%d unsupported (decoder tail is 32 channels)", C);
  GnbShape s;
  NM_CHECK_ARG(gnb_shape(n, S, C, groups, &s) && n <= 65535, "nm_final_recon_backward_fused: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(workspace);
  float* mr_unused = partial + (size_t)n * kGnbChunks * C * 2;
  float* m12 = mr_unused + (size_t)n * groups * 2;
  float* chan = m12 + (size_t)n * groups * 2;
  float* xs_unused = chan + (size_t)n * C * 3;
  float* dx14 = xs_unused + (size_t)n * C;
  float* tailp = dx14 + (size_t)n * S;
  const __half* xh = reinterpret_cast<const __half*>(x);
  const float inv = 1.0f / grad_scale;
  tail_bwd_pass1_kernel<32><<<dim3(kGnbChunks, n), 256, 0, st>>>(xh, a, b, w, bias, bias_dev, sharpness, recon, target, grad_bce, grad_scale, S, dx14,
                                                                partial, tailp);
  NM_CHECK_LAUNCH("tail_bwd_pass1_kernel");
  tail_reduce_kernel<<<1, 64, 0, st>>>(tailp, (long long)n * kGnbChunks, C, dw, dbias);
  NM_CHECK_LAUNCH("tail_reduce_kernel");
  gnb_grad_finalize_kernel<<<nm_cdiv((long long)n * groups, 4), 128, 0, st>>>(partial, gamma, mean_rstd, xsum, s, m12, chan);
  NM_CHECK_LAUNCH("gnb_grad_finalize_kernel");
  if (dgamma || dbeta || dxsum) {
    gnb_param_grad_kernel<<<nm_cdiv(C, 4), 128, 0, st>>>(chan, n, C, inv, dgamma, dbeta, dxsum);
    NM_CHECK_LAUNCH("gnb_param_grad_kernel");
  }
  gnb_dx_rank1_kernel<<<gnb_dx_grid(s), kGnbThreads, 0, st>>>(xh, dx14, w, gamma, beta, mean_rstd, m12, s, reinterpret_cast<__half*>(grad_x));
  NM_CHECK_LAUNCH("gnb_dx_rank1_kernel");
  return NM_OK;
}
