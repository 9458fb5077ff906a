// Pointwise-shaped convolutions of the feature nets, Cin = 32 / 64: the learned 2x down-sample Conv3d(k2, s2)
// (Pool3DBlock, reference modules/vox_modules.py:49-61) and the 1x1 skip convolution of Res3DBlock
// (modules/vox_modules.py:35-38).  These layers move 10-20 bytes per FLOP-hundred: they are HBM-bound, so the kernel
// is organised around the memory pipe, not the tensor pipe:
//   * every input byte is loaded exactly once, by 16-byte coalesced global loads straight into the registers that
//     become the mma.sync A fragments - the K order of the GEMM is permuted (here and in the packed weights) so that
//     the 8 consecutive channels a lane loads are exactly the K slots that lane owns in two k-steps: no shared-memory
//     staging, no shuffles;
//   * the producer's GroupNorm scale/shift (+LeakyReLU) is applied to those registers (fp32 math), optionally summed
//     with a second normalised tensor (the two branches of a Res3DBlock), so the activated / summed tensor never
//     exists in HBM;
//   * fragment columns are permuted so that a lane owns 8 consecutive output channels: 16-byte coalesced stores
//     straight from the accumulators;
//   * the per-channel sum / sum of squares the following GroupNorm needs are accumulated from the fp32 accumulators
//     and written as one deterministic partial per block.
// fp16 operands, fp32 accumulation (same numerics as the tcgen05 kernels).
#include "common.cuh"
#include "../../include/nm_b200.h"

namespace {

__device__ __forceinline__ void mma_m16n8k16(float (&c)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2,
                                             const uint32_t a3, const uint2 b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b.x), "r"(b.y));
}

// Output channel of fragment column g of n-block nb (see the header comment)
__host__ __device__ inline int pw_channel(int nb, int g) { return (nb >> 2) * 32 + (g >> 1) * 8 + (nb & 3) * 2 + (g & 1); }

// wfrag[(tap*CK + m)*2 + s][nb][lane] = B fragment of k-step s of the 32-channel group m of tap `tap`:
// K slot (2t+j+8*hi) <-> ci = m*32 + t*8 + s*4 + hi*2 + j
__global__ void pack_pw_kernel(const float* __restrict__ w /* (Cout, Cin, k, k, k) */, int Cin, int Cout, int taps,
                               uint2* __restrict__ wfrag) {
  const int NB = Cout / 8, CK = Cin / 32;
  const int total = taps * CK * 2 * NB * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int lane = i & 31, nb = (i >> 5) % NB, ks = (i >> 5) / NB, s = ks & 1, m = (ks >> 1) % CK, tap = (ks >> 1) / CK;
    const int g = lane >> 2, t = lane & 3, co = pw_channel(nb, g);
    auto wv = [&](int ci) -> float { return w[((long long)co * Cin + ci) * taps + tap]; };
    const int c0 = m * 32 + t * 8 + s * 4;
    __half2 b0 = __floats2half2_rn(wv(c0), wv(c0 + 1));
    __half2 b1 = __floats2half2_rn(wv(c0 + 2), wv(c0 + 3));
    wfrag[i] = make_uint2(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1));
  }
}

// 8 fp16 channels: y = lrelu?(x * a + b) [+ x2 * a2 + b2], fp32 math, back to fp16; coefficients from shared memory
__device__ __forceinline__ uint4 xform8(const uint4 v, const float* a, const float* b, const bool act) {
  const uint32_t in[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&in[i]));
    __half2 h = __floats2half2_rn(fmaf(f.x, a[2 * i], b[2 * i]), fmaf(f.y, a[2 * i + 1], b[2 * i + 1]));
    if (act) h = __hmax2(h, __hmul2(h, __float2half2_rn(0.01f)));
    o[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ uint4 xform8_dual(const uint4 v, const uint4 v2, const float* a, const float* b,
                                             const float* a2, const float* b2, const bool act) {
  const uint32_t in[4] = {v.x, v.y, v.z, v.w}, in2[4] = {v2.x, v2.y, v2.z, v2.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&in[i]));
    const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&in2[i]));
    float y0 = fmaf(f.x, a[2 * i], b[2 * i]), y1 = fmaf(f.y, a[2 * i + 1], b[2 * i + 1]);
    if (act) { y0 = nm_lrelu(y0); y1 = nm_lrelu(y1); }
    y0 += fmaf(f2.x, a2[2 * i], b2[2 * i]);
    y1 += fmaf(f2.y, a2[2 * i + 1], b2[2 * i + 1]);
    __half2 h = __floats2half2_rn(y0, y1);
    o[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

struct PwParams {
  const act_t* x;
  const act_t* x2;         // optional second input (same shape), added after its own scale/shift
  const uint2* wfrag;
  const float* bias;
  act_t* out;
  const float* in_scale;   // (n, Cin) or null
  const float* in_shift;
  const float* in_scale2;  // (n, Cin) or null (x2 added as is)
  const float* in_shift2;
  float* stats;            // [n][chunks][COUT][2] or null
  int in_act;
  int H, W;                // input dims (D implied)
  int OH, OW;              // output dims
  long long in_frame, out_frame;   // elements per frame
  int tiles_per_block;     // M-tiles (16 output voxels) per block
  uint32_t ow_magic, oh_magic;
};

// TAPS = 1: 1x1 conv; TAPS = 8: k2 s2.  One M-tile (16 output voxels) per warp iteration; a "step" is one
// (tap, 32-channel group): 2 (or 4 with a second input) 16-byte loads per lane, prefetched one step ahead.
template <int CIN, int COUT, int TAPS, bool DUAL>
__global__ void __launch_bounds__(256, (COUT >= 128 || CIN >= 64) ? 1 : 2) conv_pw_kernel(const PwParams p) {
  constexpr int NB = COUT / 8, CK = CIN / 32, STEPS = TAPS * CK;
  extern __shared__ __align__(16) uint2 s_w[];           // [STEPS * 2][NB][32]
  __shared__ float s_red[8][COUT][2];
  __shared__ float s_ab[4][CIN];                         // a1 | b1 | a2 | b2 of this frame
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int i = threadIdx.x; i < STEPS * 2 * NB * 32; i += 256) s_w[i] = __ldg(p.wfrag + i);
  const bool xf = p.in_scale != nullptr;
  constexpr bool dual = DUAL;
  const bool act = p.in_act != 0;
  for (int i = threadIdx.x; i < CIN; i += 256) {
    s_ab[0][i] = xf ? p.in_scale[n * CIN + i] : 1.f;
    s_ab[1][i] = xf ? p.in_shift[n * CIN + i] : 0.f;
    s_ab[2][i] = p.in_scale2 ? p.in_scale2[n * CIN + i] : 1.f;
    s_ab[3][i] = p.in_shift2 ? p.in_shift2[n * CIN + i] : 0.f;
  }
  float bia[NB][2];
#pragma unroll
  for (int nb = 0; nb < NB; nb++) {
    bia[nb][0] = __ldg(p.bias + pw_channel(nb, 2 * t));
    bia[nb][1] = __ldg(p.bias + pw_channel(nb, 2 * t + 1));
  }
  float ssum[NB][2], ssq[NB][2];
#pragma unroll
  for (int nb = 0; nb < NB; nb++) ssum[nb][0] = ssum[nb][1] = ssq[nb][0] = ssq[nb][1] = 0.f;
  __syncthreads();

  const long long fr_in = (long long)n * p.in_frame + t * 8;
  const act_t* xin = p.x + fr_in;
  const act_t* xin2 = dual ? p.x2 + fr_in : nullptr;
  act_t* xout = p.out + (long long)n * p.out_frame + t * 8;
  const int tile_base = chunk * p.tiles_per_block;
  const int tiles = p.tiles_per_block;

  // input offset (elements) of output voxel o, tap 0
  auto in_off = [&](int o) -> int {
    if (TAPS == 1) return o * CIN;
    const int q = (int)__umulhi((uint32_t)o, p.ow_magic);       // o / OW
    const int oz = o - q * p.OW;
    const int ox = (int)__umulhi((uint32_t)q, p.oh_magic);      // q / OH
    const int oy = q - ox * p.OH;
    return (((2 * ox) * p.H + 2 * oy) * p.W + 2 * oz) * CIN;
  };
  auto step_off = [&](int st) -> int {                          // element offset of step st = tap * CK + m
    const int tap = st / CK, m = st % CK;
    int o = m * 32;
    if (TAPS != 1) o += ((((tap >> 2) & 1) * p.H + ((tap >> 1) & 1)) * p.W + (tap & 1)) * CIN;
    return o;
  };

  uint4 cur[2], nxt[2], cur2[2], nxt2[2];
  int off[2];
  if (warp < tiles) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
      off[r] = in_off((tile_base + warp) * 16 + r * 8 + g);
      cur[r] = __ldg(reinterpret_cast<const uint4*>(xin + off[r]));
      if (dual) cur2[r] = __ldg(reinterpret_cast<const uint4*>(xin2 + off[r]));
    }
  }
#pragma unroll 1
  for (int tl = warp; tl < tiles; tl += 8) {
    const int tile0 = tile_base + tl;
    float c[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; nb++) {
      c[nb][0] = c[nb][2] = bia[nb][0];
      c[nb][1] = c[nb][3] = bia[nb][1];
    }
    int noff[2];
    const bool more = tl + 8 < tiles;
#pragma unroll
    for (int r = 0; r < 2; r++) noff[r] = more ? in_off((tile0 + 8) * 16 + r * 8 + g) : off[r];
#pragma unroll
    for (int st = 0; st < STEPS; st++) {
      // prefetch the next step's rows (next tap / channel group, or step 0 of this warp's next M-tile)
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if (st + 1 < STEPS) {
          nxt[r] = __ldg(reinterpret_cast<const uint4*>(xin + off[r] + step_off(st + 1)));
          if (dual) nxt2[r] = __ldg(reinterpret_cast<const uint4*>(xin2 + off[r] + step_off(st + 1)));
        } else if (more) {
          nxt[r] = __ldg(reinterpret_cast<const uint4*>(xin + noff[r]));
          if (dual) nxt2[r] = __ldg(reinterpret_cast<const uint4*>(xin2 + noff[r]));
        }
      }
      const int cb = (st % CK) * 32 + t * 8;                    // first channel of this lane's 8
      uint4 r0 = cur[0], r1 = cur[1];
      if (dual) {
        r0 = xform8_dual(cur[0], cur2[0], s_ab[0] + cb, s_ab[1] + cb, s_ab[2] + cb, s_ab[3] + cb, act);
        r1 = xform8_dual(cur[1], cur2[1], s_ab[0] + cb, s_ab[1] + cb, s_ab[2] + cb, s_ab[3] + cb, act);
      } else if (xf) {
        r0 = xform8(cur[0], s_ab[0] + cb, s_ab[1] + cb, act);
        r1 = xform8(cur[1], s_ab[0] + cb, s_ab[1] + cb, act);
      }
#pragma unroll
      for (int nb = 0; nb < NB; nb++) {
        mma_m16n8k16(c[nb], r0.x, r1.x, r0.y, r1.y, s_w[((st * 2) * NB + nb) * 32 + lane]);
        mma_m16n8k16(c[nb], r0.z, r1.z, r0.w, r1.w, s_w[((st * 2 + 1) * NB + nb) * 32 + lane]);
      }
#pragma unroll
      for (int r = 0; r < 2; r++) { cur[r] = nxt[r]; cur2[r] = nxt2[r]; }
    }
#pragma unroll
    for (int r = 0; r < 2; r++) off[r] = noff[r];
    // epilogue: statistics from the fp32 accumulators, 16-byte stores of 8 consecutive channels
#pragma unroll
    for (int nb = 0; nb < NB; nb++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const float v0 = c[nb][j], v1 = c[nb][2 + j];
        ssum[nb][j] += v0 + v1;
        ssq[nb][j] = fmaf(v0, v0, fmaf(v1, v1, ssq[nb][j]));
      }
    act_t* o0 = xout + (long long)(tile0 * 16 + g) * COUT;
#pragma unroll
    for (int q = 0; q < NB / 4; q++) {
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        __half2 h0 = __floats2half2_rn(c[4 * q + i][0], c[4 * q + i][1]);
        __half2 h1 = __floats2half2_rn(c[4 * q + i][2], c[4 * q + i][3]);
        pk[i] = *reinterpret_cast<uint32_t*>(&h0);
        pk[4 + i] = *reinterpret_cast<uint32_t*>(&h1);
      }
      *reinterpret_cast<uint4*>(o0 + q * 32) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(o0 + 8 * COUT + q * 32) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
  }
  if (p.stats == nullptr) return;
  // per-channel partials of this block: sum over the 8 fragment rows (shuffles), then over the 8 warps (fixed order)
#pragma unroll
  for (int nb = 0; nb < NB; nb++)
#pragma unroll
    for (int j = 0; j < 2; j++) {
      float a = ssum[nb][j], q = ssq[nb][j];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (g == 0) {
        const int ch = pw_channel(nb, 2 * t + j);
        s_red[warp][ch][0] = a;
        s_red[warp][ch][1] = q;
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < COUT * 2; i += 256) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) a += s_red[w][i >> 1][i & 1];
    p.stats[(((long long)n * gridDim.x + chunk) * COUT) * 2 + i] = a;
  }
}

// ConvTranspose3d(k2, s2), Cin = 32 (Upsample3DBlock, reference modules/vox_modules.py:63-75): every input voxel
// feeds the 8 output voxels (2x+dx, 2y+dy, 2z+dz) through 8 independent 1x1 convs.  One M-tile = 16 input voxels,
// loaded once (two 16-byte loads per lane, same K permutation as above); per tap 2 x NB MMAs and 16-byte stores of
// 8 consecutive channels.  GroupNorm statistics of the output from the accumulators.
// wfrag[(tap*2 + s)][nb][lane]; weight is the nn.ConvTranspose3d (Cin, Cout, 2, 2, 2) tensor.
__global__ void pack_pwT_kernel(const float* __restrict__ w, int Cout, uint2* __restrict__ wfrag) {
  const int NB = Cout / 8;
  const int total = 8 * 2 * NB * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int lane = i & 31, nb = (i >> 5) % NB, ks = (i >> 5) / NB, s = ks & 1, tap = ks >> 1;
    const int g = lane >> 2, t = lane & 3, co = pw_channel(nb, g);
    auto wv = [&](int ci) -> float { return w[((long long)ci * Cout + co) * 8 + tap]; };
    const int c0 = t * 8 + s * 4;
    __half2 b0 = __floats2half2_rn(wv(c0), wv(c0 + 1));
    __half2 b1 = __floats2half2_rn(wv(c0 + 2), wv(c0 + 3));
    wfrag[i] = make_uint2(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1));
  }
}

struct PwTParams {
  const act_t* x;
  const uint2* wfrag;
  const float* bias;
  act_t* out;
  float* stats;            // [n][chunks][COUT][2] or null
  int H, W;                // input dims (D implied)
  long long in_frame, out_frame;
  int tiles_per_block;
  uint32_t w_magic, h_magic;
};

template <int COUT>
__global__ void __launch_bounds__(256) convT_pw_kernel(const PwTParams p) {
  constexpr int NB = COUT / 8;
  extern __shared__ __align__(16) uint2 s_w[];           // [8 taps * 2][NB][32]
  __shared__ float s_red[8][COUT][2];
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int i = threadIdx.x; i < 16 * NB * 32; i += 256) s_w[i] = __ldg(p.wfrag + i);
  float bia[NB][2];
#pragma unroll
  for (int nb = 0; nb < NB; nb++) {
    bia[nb][0] = __ldg(p.bias + pw_channel(nb, 2 * t));
    bia[nb][1] = __ldg(p.bias + pw_channel(nb, 2 * t + 1));
  }
  float ssum[NB][2], ssq[NB][2];
#pragma unroll
  for (int nb = 0; nb < NB; nb++) ssum[nb][0] = ssum[nb][1] = ssq[nb][0] = ssq[nb][1] = 0.f;
  __syncthreads();
  const act_t* xin = p.x + (long long)n * p.in_frame + t * 8;
  act_t* xout = p.out + (long long)n * p.out_frame + t * 8;
  const int OH = 2 * p.H, OW = 2 * p.W;
#pragma unroll 1
  for (int tl = warp; tl < p.tiles_per_block; tl += 8) {
    const int v0 = (chunk * p.tiles_per_block + tl) * 16 + g;
    uint4 a[2];
    long long obase[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int v = v0 + r * 8;
      a[r] = __ldg(reinterpret_cast<const uint4*>(xin + (long long)v * 32));
      const int q = (int)__umulhi((uint32_t)v, p.w_magic);      // v / W
      const int z = v - q * p.W;
      const int x = (int)__umulhi((uint32_t)q, p.h_magic);      // q / H
      const int y = q - x * p.H;
      obase[r] = (((long long)(2 * x) * OH + 2 * y) * OW + 2 * z) * COUT;
    }
#pragma unroll
    for (int tap = 0; tap < 8; tap++) {
      float c[NB][4];
#pragma unroll
      for (int nb = 0; nb < NB; nb++) {
        c[nb][0] = c[nb][2] = bia[nb][0];
        c[nb][1] = c[nb][3] = bia[nb][1];
        mma_m16n8k16(c[nb], a[0].x, a[1].x, a[0].y, a[1].y, s_w[((tap * 2) * NB + nb) * 32 + lane]);
        mma_m16n8k16(c[nb], a[0].z, a[1].z, a[0].w, a[1].w, s_w[((tap * 2 + 1) * NB + nb) * 32 + lane]);
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const float u0 = c[nb][j], u1 = c[nb][2 + j];
          ssum[nb][j] += u0 + u1;
          ssq[nb][j] = fmaf(u0, u0, fmaf(u1, u1, ssq[nb][j]));
        }
      }
      const long long toff = ((long long)(((tap >> 2) & 1) * OH + ((tap >> 1) & 1)) * OW + (tap & 1)) * COUT;
#pragma unroll
      for (int q = 0; q < NB / 4; q++) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          __half2 h0 = __floats2half2_rn(c[4 * q + i][0], c[4 * q + i][1]);
          __half2 h1 = __floats2half2_rn(c[4 * q + i][2], c[4 * q + i][3]);
          pk[i] = *reinterpret_cast<uint32_t*>(&h0);
          pk[4 + i] = *reinterpret_cast<uint32_t*>(&h1);
        }
        *reinterpret_cast<uint4*>(xout + obase[0] + toff + q * 32) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(xout + obase[1] + toff + q * 32) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
  }
  if (p.stats == nullptr) return;
#pragma unroll
  for (int nb = 0; nb < NB; nb++)
#pragma unroll
    for (int j = 0; j < 2; j++) {
      float a_ = ssum[nb][j], q = ssq[nb][j];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        a_ += __shfl_xor_sync(0xffffffffu, a_, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (g == 0) {
        const int ch = pw_channel(nb, 2 * t + j);
        s_red[warp][ch][0] = a_;
        s_red[warp][ch][1] = q;
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < COUT * 2; i += 256) {
    float a_ = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) a_ += s_red[w][i >> 1][i & 1];
    p.stats[(((long long)n * gridDim.x + chunk) * COUT) * 2 + i] = a_;
  }
}

struct PwTPlan { bool ok; int tiles, tiles_per_block, chunks; };
PwTPlan plan_pwT(int n, int D, int H, int W, int Cin, int Cout) {
  PwTPlan pl;
  memset(&pl, 0, sizeof(pl));
  const long long M = (long long)D * H * W;
  if (Cin != 32 || (Cout != 32 && Cout != 64) || n <= 0 || n > 65535 || W % 8 != 0 || M % 16 != 0 ||
      M * 8 * Cout >= (1ll << 40))
    return pl;
  pl.tiles = (int)(M / 16);
  int tpb = 8;                                        // one M-tile per warp: many small blocks (frames are tiny)
  while (tpb > 1 && pl.tiles % tpb != 0) tpb >>= 1;
  pl.tiles_per_block = tpb;
  pl.chunks = pl.tiles / tpb;
  pl.ok = true;
  return pl;
}

struct PwPlan {
  bool ok;
  int taps, OD, OH, OW, tiles, tiles_per_block, chunks;
};

bool pw_shape_ok(int Cin, int Cout, int k, int stride) {
  const bool ks = (k == 1 && stride == 1) || (k == 2 && stride == 2);
  if (!ks) return false;
  if (Cin == 32) return Cout == 32 || Cout == 64;
  if (Cin == 64) return (k == 1 && (Cout == 64 || Cout == 128)) || (k == 2 && Cout == 64);
  return false;
}

PwPlan plan_pw(int n, int D, int H, int W, int Cin, int Cout, int k, int stride) {
  PwPlan pl;
  memset(&pl, 0, sizeof(pl));
  if (!pw_shape_ok(Cin, Cout, k, stride) || n <= 0 || n > 65535) return pl;
  if (stride == 2 && ((D | H | W) & 1)) return pl;
  pl.taps = k == 1 ? 1 : 8;
  pl.OD = D / stride; pl.OH = H / stride; pl.OW = W / stride;
  const long long M = (long long)pl.OD * pl.OH * pl.OW;
  if (pl.OW % 8 != 0 || M % 16 != 0 || (long long)D * H * W * Cin >= (1ll << 31)) return pl;
  pl.tiles = (int)(M / 16);
  // 8 warps x 8 M-tiles per block when the frame is large enough
  int tpb = 64;
  while (tpb > 1 && pl.tiles % tpb != 0) tpb >>= 1;
  pl.tiles_per_block = tpb;
  pl.chunks = pl.tiles / tpb;
  pl.ok = true;
  return pl;
}

template <int CIN, int COUT, int TAPS, bool DUAL = false>
int launch_pw(const PwParams& p, int chunks, int n, cudaStream_t st) {
  const size_t smem = (size_t)TAPS * (CIN / 16) * (COUT / 8) * 32 * sizeof(uint2);
  if (smem > 48 * 1024) NM_PER_DEVICE_ONCE({
    NM_CHECK_CUDA(cudaFuncSetAttribute(conv_pw_kernel<CIN, COUT, TAPS, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  conv_pw_kernel<CIN, COUT, TAPS, DUAL><<<dim3(chunks, n), 256, smem, st>>>(p);
  NM_CHECK_LAUNCH("conv3d_pw");
  return NM_OK;
}

}  // namespace

extern "C" int nm_conv3d_pw_supported(int n, int D, int H, int W, int Cin, int Cout, int k, int stride) {
  return plan_pw(n, D, H, W, Cin, Cout, k, stride).ok ? 1 : 0;
}

extern "C" int nm_conv3d_pw_dual_supported(int n, int D, int H, int W, int Cin, int Cout, int k, int stride) {
  return Cin == 64 && Cout == 64 && k == 2 && plan_pw(n, D, H, W, Cin, Cout, k, stride).ok ? 1 : 0;
}

extern "C" int nm_conv3d_pw_stats_chunks(int n, int D, int H, int W, int Cin, int Cout, int k, int stride) {
  const PwPlan pl = plan_pw(n, D, H, W, Cin, Cout, k, stride);
  return pl.ok ? pl.chunks : 0;
}

extern "C" size_t nm_conv3d_pw_packed_bytes(int Cin, int Cout, int k) {
  return (size_t)(k == 1 ? 1 : 8) * (Cin / 16) * (Cout / 8) * 32 * sizeof(uint2);
}

extern "C" int nm_pack_conv_pw_weights(const float* weight, int Cin, int Cout, int k, void* packed, void* stream) {
  NM_CHECK_ARG(weight && packed, "nm_pack_conv_pw_weights: null pointer");
  NM_CHECK_ARG(pw_shape_ok(Cin, Cout, k, k), "nm_pack_conv_pw_weights: Cin=%d Cout=%d k=%d unsupported", Cin, Cout, k);
  pack_pw_kernel<<<32, 256, 0, (cudaStream_t)stream>>>(weight, Cin, Cout, k == 1 ? 1 : 8, (uint2*)packed);
  NM_CHECK_LAUNCH("pack_conv_pw_weights");
  return NM_OK;
}

extern "C" int nm_conv3d_pw(const void* x, const void* packed_w, const float* bias, void* out, int n, int D, int H,
                            int W, int Cin, int Cout, int k, int stride, const float* in_scale, const float* in_shift,
                            int in_act, const void* x2, const float* in_scale2, const float* in_shift2,
                            float* stats_partial, void* stream) {
  NM_CHECK_ARG(x && packed_w && bias && out, "nm_conv3d_pw: null pointer");
  NM_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "nm_conv3d_pw: in_scale and in_shift go together");
  NM_CHECK_ARG((in_scale2 == nullptr) == (in_shift2 == nullptr) && (x2 || !in_scale2),
               "nm_conv3d_pw: in_scale2 / in_shift2 go together and need x2");
  NM_CHECK_ARG(!x2 || (Cin == 64 && Cout == 64 && k == 2), "nm_conv3d_pw: the second input is only implemented for the "
               "64->64 k2 s2 pool (see nm_conv3d_pw_dual_supported)");
  if (n == 0) return NM_OK;
  const PwPlan pl = plan_pw(n, D, H, W, Cin, Cout, k, stride);
  NM_CHECK_ARG(pl.ok, "nm_conv3d_pw: unsupported shape n=%d %dx%dx%d Cin=%d Cout=%d k=%d s=%d (see nm_conv3d_pw_supported)",
               n, D, H, W, Cin, Cout, k, stride);
  PwParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const act_t*)x; p.x2 = (const act_t*)x2; p.wfrag = (const uint2*)packed_w; p.bias = bias; p.out = (act_t*)out;
  p.in_scale = in_scale; p.in_shift = in_shift; p.in_scale2 = in_scale2; p.in_shift2 = in_shift2;
  p.in_act = in_act; p.stats = stats_partial;
  p.H = H; p.W = W; p.OH = pl.OH; p.OW = pl.OW;
  p.in_frame = (long long)D * H * W * Cin;
  p.out_frame = (long long)pl.OD * pl.OH * pl.OW * Cout;
  p.tiles_per_block = pl.tiles_per_block;
  p.ow_magic = 0xffffffffu / (uint32_t)pl.OW + 1u;
  p.oh_magic = 0xffffffffu / (uint32_t)pl.OH + 1u;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 32) {
    if (pl.taps == 8 && Cout == 32) return launch_pw<32, 32, 8>(p, pl.chunks, n, st);
    if (pl.taps == 8 && Cout == 64) return launch_pw<32, 64, 8>(p, pl.chunks, n, st);
    if (pl.taps == 1 && Cout == 32) return launch_pw<32, 32, 1>(p, pl.chunks, n, st);
    return launch_pw<32, 64, 1>(p, pl.chunks, n, st);
  }
  if (pl.taps == 8) return x2 ? launch_pw<64, 64, 8, true>(p, pl.chunks, n, st) : launch_pw<64, 64, 8>(p, pl.chunks, n, st);
  if (Cout == 64) return launch_pw<64, 64, 1>(p, pl.chunks, n, st);
  return launch_pw<64, 128, 1>(p, pl.chunks, n, st);
}

extern "C" int nm_conv_transpose3d_pw_supported(int n, int D, int H, int W, int Cin, int Cout) {
  return plan_pwT(n, D, H, W, Cin, Cout).ok ? 1 : 0;
}

extern "C" int nm_conv_transpose3d_pw_stats_chunks(int n, int D, int H, int W, int Cin, int Cout) {
  const PwTPlan pl = plan_pwT(n, D, H, W, Cin, Cout);
  return pl.ok ? pl.chunks : 0;
}

extern "C" size_t nm_conv_transpose3d_pw_packed_bytes(int Cin, int Cout) {
  return (size_t)8 * 2 * (Cout / 8) * 32 * sizeof(uint2);
}

extern "C" int nm_pack_conv_transpose3d_pw_weights(const float* weight, int Cin, int Cout, void* packed, void* stream) {
  NM_CHECK_ARG(weight && packed, "nm_pack_conv_transpose3d_pw_weights: null pointer");
  NM_CHECK_ARG(Cin == 32 && (Cout == 32 || Cout == 64), "nm_pack_conv_transpose3d_pw_weights: Cin=%d Cout=%d unsupported", Cin, Cout);
  pack_pwT_kernel<<<16, 256, 0, (cudaStream_t)stream>>>(weight, Cout, (uint2*)packed);
  NM_CHECK_LAUNCH("pack_conv_transpose3d_pw_weights");
  return NM_OK;
}

extern "C" int nm_conv_transpose3d_pw(const void* x, const void* packed_w, const float* bias, void* out, int n, int D,
                                      int H, int W, int Cin, int Cout, float* stats_partial, void* stream) {
  NM_CHECK_ARG(x && packed_w && bias && out, "nm_conv_transpose3d_pw: null pointer");
  if (n == 0) return NM_OK;
  const PwTPlan pl = plan_pwT(n, D, H, W, Cin, Cout);
  NM_CHECK_ARG(pl.ok, "nm_conv_transpose3d_pw: unsupported shape n=%d %dx%dx%d Cin=%d Cout=%d", n, D, H, W, Cin, Cout);
  PwTParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const act_t*)x; p.wfrag = (const uint2*)packed_w; p.bias = bias; p.out = (act_t*)out; p.stats = stats_partial;
  p.H = H; p.W = W;
  p.in_frame = (long long)D * H * W * Cin;
  p.out_frame = (long long)D * H * W * 8 * Cout;
  p.tiles_per_block = pl.tiles_per_block;
  p.w_magic = 0xffffffffu / (uint32_t)W + 1u;
  p.h_magic = 0xffffffffu / (uint32_t)H + 1u;
  const size_t smem = nm_conv_transpose3d_pw_packed_bytes(Cin, Cout);
  cudaStream_t st = (cudaStream_t)stream;
  if (Cout == 32) convT_pw_kernel<32><<<dim3(pl.chunks, n), 256, smem, st>>>(p);
  else convT_pw_kernel<64><<<dim3(pl.chunks, n), 256, smem, st>>>(p);
  NM_CHECK_LAUNCH("conv_transpose3d_pw");
  return NM_OK;
}
