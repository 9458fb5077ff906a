// CUDA-core convolutions:
//  * the first k5 layer of each feature net (Cin = 1 occupancy + 3 analytic CoordConv
//    channels; reference: Basic3DBlock(1+3, C, 5) at model/kypt_detector.py:266 fed by
//    add_coord_channels, utils/kypt_detector_utils.py:4-26).  The coordinate channels are
//    never materialised: their contribution is a per-boundary-class affine function of the
//    voxel position (tables built once per weight set); the occupancy channel is ~98 %
//    zeros, so only non-zero taps are accumulated.  Output is dense-equivalent.
//  * ConvTranspose3d(k2, s2) of the hour-glass (modules/vox_modules.py:68) - tiny layers.
//  * a generic direct convolution used as the on-device cross-check of the tcgen05 kernel.
#include "common.cuh"
#include "../../include/nm_b200.h"   // the definitions below must match the public declarations
#include <type_traits>

namespace {

// ------------------------------------------------------------------ first layer
// cls = (cx*5+cy)*5+cz with per-axis class 0,1 (p=0,1), 2 (interior), 3,4 (p=G-2,G-1).
// cls_tab[cls][co] = (sum_a S1_a, S0_x, S0_y, S0_z) with S0_a = sum of the valid-tap weights of coordinate
// channel a and S1_a = sum of (k_a - 2) * weight: the CoordConv term of a voxel of that class is
// bias + step*.x + .y*lin[x] + .z*lin[y] + .w*lin[z], step = 2/(G-1).
// wfrag[s][nb][lane]: the occupancy-channel weights as fp16 B fragments of mma.m16n8k16 (columns permuted, see
// first_conv_prep_kernel).  K index = s*16 + col,
// col < 8: window row r = 2s (r = kx*5+ky), kz = col; col >= 8: r = 2s+1, kz = col-8; kz >= 5 and r = 25 are zero.
constexpr int kFirstKSteps = 13;

__global__ void first_conv_prep_kernel(const float* __restrict__ w /* (Cout,4,5,5,5) */, int Cout,
                                       float4* __restrict__ cls_tab, uint2* __restrict__ wfrag) {
  const int cls = blockIdx.x;
  const int cl[3] = {cls / 25, (cls / 5) % 5, cls % 5};
  int lo[3], hi[3];
  for (int a = 0; a < 3; a++) {
    lo[a] = cl[a] == 0 ? 2 : (cl[a] == 1 ? 1 : 0);
    hi[a] = cl[a] == 4 ? 2 : (cl[a] == 3 ? 3 : 4);
  }
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    double s0[3], s1 = 0.0;
    for (int c = 0; c < 3; c++) {
      s0[c] = 0.0;
      for (int kx = lo[0]; kx <= hi[0]; kx++)
        for (int ky = lo[1]; ky <= hi[1]; ky++)
          for (int kz = lo[2]; kz <= hi[2]; kz++) {
            const double v = (double)w[(((long long)co * 4 + 1 + c) * 5 + kx) * 25 + ky * 5 + kz];
            const int kc = c == 0 ? kx : (c == 1 ? ky : kz);
            s0[c] += v;
            s1 += v * (double)(kc - 2);
          }
    }
    cls_tab[cls * Cout + co] = make_float4((float)s1, (float)s0[0], (float)s0[1], (float)s0[2]);
  }
  // B fragments: b0 = {B[2t][g], B[2t+1][g]}, b1 = {B[2t+8][g], B[2t+9][g]} with n = nb*8 + g
  const int NB = Cout / 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kFirstKSteps * NB * 32; i += gridDim.x * blockDim.x) {
    const int lane = i & 31, nb = (i >> 5) % NB, s = (i >> 5) / NB;
    // fragment column g of n-block nb is output channel (nb/4)*32 + (g>>1)*8 + (nb%4)*2 + (g&1): thread t of a quad
    // then owns the 8 consecutive channels t*8 .. t*8+7 of a voxel (one 16-byte store, no staging)
    const int g = lane >> 2, t = lane & 3, co = (nb >> 2) * 32 + (g >> 1) * 8 + (nb & 3) * 2 + (g & 1);
    auto wv = [&](int col) -> float {
      const int r = 2 * s + (col >> 3), kz = col & 7;
      if (r >= 25 || kz >= 5) return 0.f;
      return w[((long long)co * 4) * 125 + r * 5 + kz];
    };
    __half2 b0 = __floats2half2_rn(wv(2 * t), wv(2 * t + 1));
    __half2 b1 = __floats2half2_rn(wv(2 * t + 8), wv(2 * t + 9));
    wfrag[i] = make_uint2(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1));
  }
}

__device__ __forceinline__ void mma_m16n8k16(float (&c)[4], const uint32_t (&a)[4], const uint2 b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

__device__ __forceinline__ int axis_class(int p, int G) {
  return p == 0 ? 0 : (p == 1 ? 1 : (p == G - 2 ? 3 : (p == G - 1 ? 4 : 2)));
}

// One 256-thread block owns a 4 (x) x 8 (y) column of the volume and walks it along z in 8-voxel tiles.  The
// column's halo (8 x 12 rows of G + 8 cells, fp16, zero padded) is loaded once with coalesced 16-byte loads, together
// with one occupancy bit mask per row.  Warp w owns x = w>>1 and four y rows; an M-tile of the warp-level MMA is
// two y rows x 8 z voxels (fragment row g = z, rows 8..15 = the second y row).  The accumulators start as the
// analytic CoordConv term - one FMA per output from per-thread coefficient registers that already contain the x and
// y parts; they are reloaded where the thread's boundary class changes (first, second and last z tile).  The
// occupancy channel is added by tensor-core MMAs over K = (window row, kz), skipping every 16-wide K step whose two
// window rows are empty for the M-tile - the input is a sparse surface, so most M-tiles skip all 13.  fp16 operands
// (occupancy 0/1 is exact), fp32 accumulation.  Channels are processed 32 at a time; fragment columns are permuted
// so that a thread owns 8 consecutive channels of a voxel and stores them directly.
__constant__ int c_first_row_off[26] = {0, 1, 2, 3, 4, 12, 13, 14, 15, 16, 24, 25, 26, 27, 28, 36, 37, 38, 39, 40,
                                        48, 49, 50, 51, 52, 52};   // (r/5)*12 + r%5; r = 25 (zero weights) -> 24

template <int COUT>
__global__ void __launch_bounds__(256, 2)
first_conv_kernel(const float* __restrict__ occ, const float4* __restrict__ cls_tab, const uint2* __restrict__ wfrag,
                  const float* __restrict__ bias, const float* __restrict__ lin, int G, act_t* __restrict__ out,
                  float* __restrict__ stats /* [n][blocks per frame][COUT][2] or null */) {
  extern __shared__ __align__(16) uint32_t dyn_smem[];
  __shared__ float s_red[8][COUT][2];
  // halo_w[R][HW]: row R = hx*12 + hy, half index i = z + 2 (z = -2 .. G+5 and zero padding up to 2*HW)
  // hb[R][HBW]: bit i of the row: halo value != 0
  // win[zt][hx*13 + hyp]: 12-bit occupancy window of z tile zt, rows (hx, hyp) | (hx, hyp+1)  (hyp = 0..10)
  // tab[(xl*NYL + yl)*5 + cz][COUT]: coefficient rows of the boundary classes this block touches (xl, yl = class
  //   minus the class of the block's first x / y), .x = step * sum S1 + bias, channels (i, t)-transposed per 32
  const int HW = G / 2 + 8, HBW = (G + 4 + 31) / 32 + 1, NZT = G / 8, NYL = G == 8 ? 5 : 3;
  float4* tab = reinterpret_cast<float4*>(dyn_smem);
  uint32_t* halo_w = dyn_smem + 3 * NYL * 5 * COUT * 4;
  uint32_t* hb = halo_w + 96 * HW;
  uint32_t* win = hb + 96 * HBW;
  const int n = blockIdx.z;
  const int x0 = blockIdx.y * 4, y0 = blockIdx.x * 8;
  const float* src = occ + (long long)n * G * G * G;
  const float step = 2.0f / (float)(G - 1);
  const int Q = G / 4;                            // float4 per row
  const int total = 96 * Q;
  const uint32_t qmagic = 0xffffffffu / (uint32_t)Q + 1u;   // i / Q = umulhi(i, qmagic) for the small i used here

  for (int i = threadIdx.x; i < 96 * HBW; i += 256) hb[i] = 0;
  for (int i = threadIdx.x; i < 96 * 8; i += 256) {           // border words of every row: 0 and G/2+1 .. G/2+7
    const int R = i >> 3, j = i & 7;
    halo_w[R * HW + (j == 0 ? 0 : G / 2 + j)] = 0;
  }
  {
    const int xbase = axis_class(x0, G), ybase = axis_class(y0, G);
    const int nx = axis_class(x0 + 3, G) - xbase + 1, ny = axis_class(y0 + 7, G) - ybase + 1;
    for (int i = threadIdx.x; i < nx * ny * 5 * COUT; i += 256) {
      const int ch = i % COUT, row = i / COUT, cz = row % 5, yl = (row / 5) % ny, xl = row / (5 * ny);
      float4 k = __ldg(cls_tab + (((xbase + xl) * 5 + ybase + yl) * 5 + cz) * COUT + ch);
      k.x = fmaf(step, k.x, __ldg(bias + ch));
      tab[((xl * NYL + yl) * 5 + cz) * COUT + (ch & ~31) + (ch & 7) * 4 + ((ch >> 3) & 3)] = k;
    }
  }
  __syncthreads();                                // hb zeroed before the atomicOr's below
#pragma unroll 1
  for (int base = 0; base < total; base += 256 * 6) {
    float4 v[6];
#pragma unroll
    for (int u = 0; u < 6; u++) {
      const int i = base + u * 256 + threadIdx.x;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < total) {
        const int R = (int)__umulhi((uint32_t)i, qmagic), q = i - R * Q;
        const int hx = R / 12, hy = R - hx * 12;
        const int x = x0 + hx - 2, y = y0 + hy - 2;
        if ((unsigned)x < (unsigned)G && (unsigned)y < (unsigned)G)
          v[u] = __ldg(reinterpret_cast<const float4*>(src + ((long long)x * G + y) * G) + q);
      }
    }
#pragma unroll
    for (int u = 0; u < 6; u++) {
      const int i = base + u * 256 + threadIdx.x;
      if (i >= total) continue;
      const int R = (int)__umulhi((uint32_t)i, qmagic), q = i - R * Q;
      __half2 h0 = __floats2half2_rn(v[u].x, v[u].y), h1 = __floats2half2_rn(v[u].z, v[u].w);
      uint32_t* row = halo_w + R * HW + 2 * q + 1;             // half index 4q + 2
      row[0] = *reinterpret_cast<uint32_t*>(&h0);
      row[1] = *reinterpret_cast<uint32_t*>(&h1);
      const uint32_t nib = (v[u].x != 0.f ? 1u : 0u) | (v[u].y != 0.f ? 2u : 0u) | (v[u].z != 0.f ? 4u : 0u) |
                           (v[u].w != 0.f ? 8u : 0u);
      if (nib) {
        const int bit = 4 * q + 2;
        atomicOr(hb + R * HBW + (bit >> 5), nib << (bit & 31));
        if ((bit & 31) > 28) atomicOr(hb + R * HBW + (bit >> 5) + 1, nib >> (32 - (bit & 31)));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NZT * 8 * 11; i += 256) {
    const int zt = i / 88, rem = i - zt * 88, hx = rem / 11, hyp = rem - hx * 11;
    const uint32_t* hr = hb + (hx * 12 + hyp) * HBW + (zt >> 2);
    win[zt * 104 + hx * 13 + hyp] = __funnelshift_r(hr[0] | hr[HBW], hr[1] | hr[HBW + 1], (zt & 3) * 8) & 0xfffu;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int lx = warp >> 1, yw = (warp & 1) * 4;
  const int x = x0 + lx;
  const float linx = __ldg(lin + x);
  const int xl = axis_class(x, G) - axis_class(x0, G);
  // window rows of this lane for the emptiness test (lanes 0..24), M-tile 0
  const int lr = lane < 25 ? lane : 0;
  const uint32_t* win_lane = win + (lx + lr / 5) * 13 + yw + lr % 5;
  const uint32_t* halo_lane = halo_w + (lx * 12 + yw) * HW + (g >> 1) + t;   // + row offset * HW + z0 / 2
  const int sh = (g & 1) * 16;

  float ssum[8], ssq[8];                                    // GroupNorm statistics of channels chb + t*8 + i
#pragma unroll 1
  for (int pass = 0; pass < COUT / 16; pass++) {            // (32-channel group, M-tile)
    const int chb = (pass >> 1) * 32, mt = pass & 1;
    const int ya = yw + 2 * mt;
    if (mt == 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) ssum[i] = ssq[i] = 0.f;
    }
    const float liny0 = __ldg(lin + y0 + ya), liny1 = __ldg(lin + y0 + ya + 1);
    const int yl0 = axis_class(y0 + ya, G) - axis_class(y0, G), yl1 = axis_class(y0 + ya + 1, G) - axis_class(y0, G);
    const float4* tab0 = tab + ((xl * NYL + yl0) * 5) * COUT + chb + t;
    const float4* tab1 = tab + ((xl * NYL + yl1) * 5) * COUT + chb + t;
    float kb[2][8], kz[2][8];                               // interior z class: c = kb + kz * lin[z]
    act_t* outp = out + ((((long long)n * G + x) * G + (y0 + ya)) * G + g) * COUT + chb + t * 8;
    const uint2* wf_lane = wfrag + (chb >> 3) * 32 + lane;
    const uint32_t* win_p = win_lane + 2 * mt;
    const uint32_t* halo_p = halo_lane + 2 * mt * HW;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float4* tp = (h ? tab1 : tab0) + 2 * COUT;
      const float ly = h ? liny1 : liny0;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float4 k = tp[i * 4];
        kb[h][i] = fmaf(k.y, linx, fmaf(k.z, ly, k.x));
        kz[h][i] = k.w;
      }
    }
    // GroupNorm statistics: interior M-tiles without occupancy hits are pure kb + kz * lin[z]: their sum and sum of
    // squares follow from (count, sum lin, sum lin^2) - 3 FMAs per M-tile instead of 32, folded once per pass; M-tiles
    // with hits and the two boundary tiles accumulate their accumulators explicitly.
    float e_cnt = 0.f, e_s1 = 0.f, e_s2 = 0.f;
    // A thread's z boundary class differs from the interior only in the first and the last z tile (z = 0, 1, G-2, G-1):
    // those two tiles evaluate the CoordConv term straight from the class table (3 FMAs per output); the tiles in
    // between use the interior coefficients held in registers (the first version reloaded the coefficients and folded
    // the statistics on 3 of the 8 tiles: ~210 instructions on those iterations against ~60 on the others).
    auto tile = [&](const int zt, auto bnd_c) {
      constexpr bool BND = decltype(bnd_c)::value;
      const int z0 = zt * 8;
      const float linz = __ldg(lin + z0 + g);
      float c[4][4];
      if (BND) {
        const int cz = axis_class(z0 + g, G);
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const float4* tp = (h ? tab1 : tab0) + cz * COUT;
          const float ly = h ? liny1 : liny0;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const float4 k = tp[i * 4];
            c[i >> 1][2 * h + (i & 1)] = fmaf(k.w, linz, fmaf(k.y, linx, fmaf(k.z, ly, k.x)));
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          c[i >> 1][i & 1] = fmaf(kz[0][i], linz, kb[0][i]);
          c[i >> 1][2 + (i & 1)] = fmaf(kz[1][i], linz, kb[1][i]);
        }
      }
      // occupancy channel
      bool hit;
      {
        const uint32_t rowmask = __ballot_sync(0xffffffffu, lane < 25 && win_p[zt * 104] != 0);
        uint32_t km = (rowmask | (rowmask >> 1)) & 0x1555555u;      // bit 2s: K step s has a non-empty window row
        hit = km != 0;
        while (km) {
          const int s2 = __ffs(km) - 1;                             // 2 * s
          km &= km - 1;
          const uint32_t* p0 = halo_p + c_first_row_off[s2] * HW + (z0 >> 1);
          const uint32_t* p1 = halo_p + c_first_row_off[s2 + 1] * HW + (z0 >> 1);
          uint32_t a[4];                                            // halves z0 + g + 2t, + 1 (odd g straddles two words)
          a[0] = __funnelshift_r(p0[0], p0[1], sh);
          a[1] = __funnelshift_r(p0[HW], p0[HW + 1], sh);
          a[2] = __funnelshift_r(p1[0], p1[1], sh);
          a[3] = __funnelshift_r(p1[HW], p1[HW + 1], sh);
          const uint2* wf = wf_lane + (s2 >> 1) * (COUT / 8) * 32;
#pragma unroll
          for (int nb = 0; nb < 4; nb++) mma_m16n8k16(c[nb], a, __ldg(wf + nb * 32));
        }
      }
      if (stats != nullptr) {
        if (BND || hit) {
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const float v0 = c[i >> 1][i & 1], v1 = c[i >> 1][2 + (i & 1)];
            ssum[i] += v0 + v1;
            ssq[i] = fmaf(v0, v0, fmaf(v1, v1, ssq[i]));
          }
        } else {
          e_cnt += 1.f;
          e_s1 += linz;
          e_s2 = fmaf(linz, linz, e_s2);
        }
      }
      // thread (g, t) holds channels t*8 .. t*8+7 of voxel g in both y rows: two 512-byte coalesced warp stores
      uint32_t pk[8];
#pragma unroll
      for (int nb = 0; nb < 4; nb++) {
        __half2 h0 = __floats2half2_rn(c[nb][0], c[nb][1]);
        __half2 h1 = __floats2half2_rn(c[nb][2], c[nb][3]);
        pk[nb] = *reinterpret_cast<uint32_t*>(&h0);
        pk[4 + nb] = *reinterpret_cast<uint32_t*>(&h1);
      }
      act_t* dst = outp + (long long)z0 * COUT;
      *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(dst + (long long)G * COUT) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    };
    tile(0, std::true_type{});
#pragma unroll 1
    for (int zt = 1; zt < NZT - 1; zt++) tile(zt, std::false_type{});
    if (NZT > 1) tile(NZT - 1, std::true_type{});
    if (e_cnt != 0.f) {
#pragma unroll
      for (int h = 0; h < 2; h++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float b_ = kb[h][i], z_ = kz[h][i];
          ssum[i] += fmaf(e_cnt, b_, z_ * e_s1);
          ssq[i] += fmaf(e_cnt * b_, b_, fmaf(2.f * b_ * z_, e_s1, z_ * z_ * e_s2));
        }
    }
    if (stats != nullptr && mt == 1) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        float a = ssum[i], q = ssq[i];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (g == 0) {
          s_red[warp][chb + t * 8 + i][0] = a;
          s_red[warp][chb + t * 8 + i][1] = q;
        }
      }
    }
  }
  if (stats == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.x; i < COUT * 2; i += 256) {       // fixed-order fold of the 8 warps: deterministic
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) a += s_red[w][i >> 1][i & 1];
    stats[(((long long)n * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * COUT * 2 + i] = a;
  }
}

// ------------------------------------------------------------------ generic direct conv (cross-check)
// x: (n, D, H, W, Cin) fp16; w: (Cout, Cin, k, k, k) fp32 (the nn.Conv3d layout); out (n, OD, OH, OW, Cout) fp16
__global__ void __launch_bounds__(128)
conv_direct_kernel(const act_t* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   act_t* __restrict__ out, int D, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                   int OD, int OH, int OW, long long total) {
  const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
  if (i >= total) return;
  const int co = (int)(i % Cout);
  long long r = i / Cout;
  const int ow = (int)(r % OW); r /= OW;
  const int oh = (int)(r % OH); r /= OH;
  const int od = (int)(r % OD);
  const long long n = r / OD;
  float acc = bias ? bias[co] : 0.f;
  for (int kd = 0; kd < k; kd++) {
    const int id = od * stride + kd - pad;
    if ((unsigned)id >= (unsigned)D) continue;
    for (int kh = 0; kh < k; kh++) {
      const int ih = oh * stride + kh - pad;
      if ((unsigned)ih >= (unsigned)H) continue;
      for (int kw = 0; kw < k; kw++) {
        const int iw = ow * stride + kw - pad;
        if ((unsigned)iw >= (unsigned)W) continue;
        const act_t* px = x + (((n * D + id) * H + ih) * (long long)W + iw) * Cin;
        const float* pw = w + ((long long)co * Cin * k * k * k) + (kd * k + kh) * k + kw;
        for (int ci = 0; ci < Cin; ci++) acc = fmaf(__half2float(px[ci]), pw[(long long)ci * k * k * k], acc);
      }
    }
  }
  out[i] = __float2half_rn(acc);
}

// ------------------------------------------------------------------ ConvTranspose3d k2 s2
// x: (n, D, H, W, Cin) fp16; w: tap-major (8, Cin, Cout) fp32; out (n, 2D, 2H, 2W, Cout)
// one thread per (output voxel, 8 output channels)
__global__ void __launch_bounds__(128)
convT_k2s2_kernel(const act_t* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  act_t* __restrict__ out, int D, int H, int W, int Cin, int Cout, long long total8) {
  const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
  if (i >= total8) return;
  const int c8n = Cout >> 3;
  const int c8 = (int)(i % c8n);
  long long r = i / c8n;
  const int ow = (int)(r % (2 * W)); r /= 2 * W;
  const int oh = (int)(r % (2 * H)); r /= 2 * H;
  const int od = (int)(r % (2 * D));
  const long long n = r / (2 * D);
  const int tap = ((od & 1) * 2 + (oh & 1)) * 2 + (ow & 1);
  const act_t* px = x + (((n * D + (od >> 1)) * H + (oh >> 1)) * (long long)W + (ow >> 1)) * Cin;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = bias[c8 * 8 + k];
  // w is tap-major [8][Cin][Cout] (host-side permute of the (Cin, Cout, 2, 2, 2) parameter): the 8 output
  // channels of this thread are two float4 loads per input channel
  const float4* pw = reinterpret_cast<const float4*>(w + ((long long)tap * Cin) * Cout + c8 * 8);
  for (int ci8 = 0; ci8 < Cin; ci8 += 8) {
    float xv[8];
    nm_unpack8(*reinterpret_cast<const half8*>(px + ci8), xv);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float4 w0 = __ldg(pw + (long long)(ci8 + j) * (Cout / 4));
      const float4 w1 = __ldg(pw + (long long)(ci8 + j) * (Cout / 4) + 1);
      acc[0] = fmaf(xv[j], w0.x, acc[0]); acc[1] = fmaf(xv[j], w0.y, acc[1]);
      acc[2] = fmaf(xv[j], w0.z, acc[2]); acc[3] = fmaf(xv[j], w0.w, acc[3]);
      acc[4] = fmaf(xv[j], w1.x, acc[4]); acc[5] = fmaf(xv[j], w1.y, acc[5]);
      acc[6] = fmaf(xv[j], w1.z, acc[6]); acc[7] = fmaf(xv[j], w1.w, acc[7]);
    }
  }
  reinterpret_cast<half8*>(out)[i] = nm_pack8(acc);
}

}  // namespace

extern "C" size_t nm_first_conv_tables_bytes(int Cout) {
  return (size_t)125 * Cout * sizeof(float4) + (size_t)kFirstKSteps * (Cout / 8) * 32 * sizeof(uint2);
}

extern "C" int nm_first_conv_prepare(const float* weight, int Cout, void* tables, void* stream) {
  NM_CHECK_ARG(weight && tables, "nm_first_conv_prepare: null pointer");
  NM_CHECK_ARG(Cout == 32 || Cout == 64, "nm_first_conv_prepare: Cout=%d unsupported", Cout);
  float4* t = (float4*)tables;
  first_conv_prep_kernel<<<125, 64, 0, (cudaStream_t)stream>>>(weight, Cout, t, (uint2*)(t + (size_t)125 * Cout));
  NM_CHECK_LAUNCH("first_conv_prepare");
  return NM_OK;
}

extern "C" int nm_first_conv_stats_chunks(int G) { return (G / 8) * (G / 4); }

extern "C" int nm_first_conv_k5(const float* occ, const void* tables, const float* bias, const float* linspace,
                                int n, int G, int Cout, void* out, float* stats_partial, void* stream) {
  NM_CHECK_ARG(occ && tables && bias && linspace && out, "nm_first_conv_k5: null pointer");
  NM_CHECK_ARG(G % 8 == 0 && G >= 8, "nm_first_conv_k5: grid %d must be a multiple of 8", G);
  if (n == 0) return NM_OK;
  const float4* t = (const float4*)tables;
  const uint2* wf = (const uint2*)(t + (size_t)125 * Cout);
  NM_CHECK_ARG(n <= 65535 && G <= 128, "nm_first_conv_k5: n=%d (max 65535) or grid %d (max 128) too large", n, G);
  dim3 grid(G / 8, G / 4, n);
  const size_t smem = ((size_t)96 * (G / 2 + 8 + (G + 4 + 31) / 32 + 1) + (size_t)(G / 8) * 104) * sizeof(uint32_t) +
                      (size_t)3 * (G == 8 ? 5 : 3) * 5 * Cout * sizeof(float4);
  NM_PER_DEVICE_ONCE({
    NM_CHECK_CUDA(cudaFuncSetAttribute(first_conv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    NM_CHECK_CUDA(cudaFuncSetAttribute(first_conv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  });
  cudaStream_t st = (cudaStream_t)stream;
  if (Cout == 32) {
    first_conv_kernel<32><<<grid, 256, smem, st>>>(occ, t, wf, bias, linspace, G, (act_t*)out, stats_partial);
  } else if (Cout == 64) {
    first_conv_kernel<64><<<grid, 256, smem, st>>>(occ, t, wf, bias, linspace, G, (act_t*)out, stats_partial);
  } else {
    NM_CHECK_ARG(false, "nm_first_conv_k5: Cout=%d unsupported", Cout);
  }
  NM_CHECK_LAUNCH("first_conv_k5");
  return NM_OK;
}

extern "C" int nm_conv3d_direct(const void* x, const float* weight, const float* bias, void* out, int n, int D,
                                int H, int W, int Cin, int Cout, int k, int stride, int pad, void* stream) {
  NM_CHECK_ARG(x && weight && out, "nm_conv3d_direct: null pointer");
  const int OD = (D + 2 * pad - k) / stride + 1, OH = (H + 2 * pad - k) / stride + 1,
            OW = (W + 2 * pad - k) / stride + 1;
  const long long total = (long long)n * OD * OH * OW * Cout;
  if (total == 0) return NM_OK;
  conv_direct_kernel<<<nm_cdiv(total, 128), 128, 0, (cudaStream_t)stream>>>(
      (const act_t*)x, weight, bias, (act_t*)out, D, H, W, Cin, Cout, k, stride, pad, OD, OH, OW, total);
  NM_CHECK_LAUNCH("conv3d_direct");
  return NM_OK;
}

extern "C" int nm_conv_transpose3d_k2s2(const void* x, const float* weight, const float* bias, void* out, int n,
                                        int D, int H, int W, int Cin, int Cout, void* stream) {
  NM_CHECK_ARG(x && weight && bias && out, "nm_conv_transpose3d_k2s2: null pointer");
  NM_CHECK_ARG(Cout % 8 == 0 && Cin % 8 == 0, "nm_conv_transpose3d_k2s2: Cin=%d Cout=%d not multiples of 8", Cin, Cout);
  const long long total8 = (long long)n * 8 * D * H * W * (Cout / 8);
  if (total8 == 0) return NM_OK;
  convT_k2s2_kernel<<<nm_cdiv(total8, 128), 128, 0, (cudaStream_t)stream>>>((const act_t*)x, weight, bias,
                                                                           (act_t*)out, D, H, W, Cin, Cout, total8);
  NM_CHECK_LAUNCH("conv_transpose3d_k2s2");
  return NM_OK;
}
