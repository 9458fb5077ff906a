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

namespace {

// ------------------------------------------------------------------ first layer
// tables[cls][c][co][2]: cls = (cx*5+cy)*5+cz with per-axis class 0,1 (p=0,1), 2 (interior),
// 3,4 (p=G-2,G-1); [0] = sum of valid-tap weights, [1] = sum of (k_c-2)*weight.
__global__ void first_conv_prep_kernel(const float* __restrict__ w /* (Cout,4,5,5,5) */, int Cout,
                                       float* __restrict__ tables, float* __restrict__ wocc /* [125][Cout] */) {
  const int cls = blockIdx.x;
  const int cl[3] = {cls / 25, (cls / 5) % 5, cls % 5};
  // valid tap range per class
  int lo[3], hi[3];
  for (int a = 0; a < 3; a++) {
    lo[a] = cl[a] == 0 ? 2 : (cl[a] == 1 ? 1 : 0);
    hi[a] = cl[a] == 4 ? 2 : (cl[a] == 3 ? 3 : 4);
  }
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    for (int c = 0; c < 3; c++) {
      double s0 = 0.0, s1 = 0.0;
      for (int kx = lo[0]; kx <= hi[0]; kx++)
        for (int ky = lo[1]; ky <= hi[1]; ky++)
          for (int kz = lo[2]; kz <= hi[2]; kz++) {
            const double v = (double)w[(((long long)co * 4 + 1 + c) * 5 + kx) * 25 + ky * 5 + kz];
            const int kc = c == 0 ? kx : (c == 1 ? ky : kz);
            s0 += v;
            s1 += v * (double)(kc - 2);
          }
      tables[(((long long)cls * 3 + c) * Cout + co) * 2] = (float)s0;
      tables[(((long long)cls * 3 + c) * Cout + co) * 2 + 1] = (float)s1;
    }
    if (cls == 0)
      for (int tap = 0; tap < 125; tap++) wocc[tap * Cout + co] = w[((long long)co * 4) * 125 + tap];
  }
}

template <int COUT>
__global__ void __launch_bounds__(256)
first_conv_kernel(const float* __restrict__ occ, const float* __restrict__ wocc, const float* __restrict__ tables,
                  const float* __restrict__ bias, const float* __restrict__ lin, int G, act_t* __restrict__ out) {
  // tile: 4 (x) x 8 (y) x 8 (z) outputs, halo 8 x 12 x 12
  __shared__ float halo[8][12][12];
  __shared__ uint32_t hbits[8][12];        // bit z of [hx][hy]: halo[hx][hy][z] != 0
  __shared__ float4 s_lin[COUT];           // interior blocks: (K0 + bias, Sx, Sy, Sz) per channel
  extern __shared__ float s_w[];  // [125][COUT] weights, later reused as the [256][COUT] fp16 store staging (128*COUT floats)
  const int n = blockIdx.y;
  const int tz = G / 8, ty = G / 8;
  int b = blockIdx.x;
  const int bz = b % tz; b /= tz;
  const int by = b % ty; b /= ty;
  const int bx = b;
  const int x0 = bx * 4, y0 = by * 8, z0 = bz * 8;
  const float* src = occ + (long long)n * G * G * G;
  bool any = false;
  // halo: 96 (x, y) rows of 12 z-values; one thread per row, six 8-byte loads (z0 - 2 is 8-byte aligned),
  // the row's occupancy bit mask is built in the same pass
  if (threadIdx.x < 96) {
    const int hx = threadIdx.x / 12, hy = threadIdx.x % 12;
    const int x = x0 + hx - 2, y = y0 + hy - 2;
    const bool row_ok = (unsigned)x < (unsigned)G && (unsigned)y < (unsigned)G;
    const float2* rp = reinterpret_cast<const float2*>(src + ((long long)x * G + y) * G + (z0 - 2));
    uint32_t word = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) {
      const int z = z0 - 2 + 2 * j;                 // both elements of a pair are in or out together (G, z0 even)
      float2 v = make_float2(0.f, 0.f);
      if (row_ok && (unsigned)z < (unsigned)G) v = __ldg(rp + j);
      halo[hx][hy][2 * j] = v.x;
      halo[hx][hy][2 * j + 1] = v.y;
      word |= (v.x != 0.f ? 1u : 0u) << (2 * j);
      word |= (v.y != 0.f ? 1u : 0u) << (2 * j + 1);
    }
    hbits[hx][hy] = word;
    any = word != 0;
  }
  const int block_any = __syncthreads_or(any);
  if (block_any)
    for (int i = threadIdx.x; i < 125 * COUT / 4; i += 256)
      reinterpret_cast<float4*>(s_w)[i] = reinterpret_cast<const float4*>(wocc)[i];
  // A block whose voxels are all >= 2 cells away from every face sees the full 5^3 window everywhere: the
  // CoordConv term is then one affine function of (x, y, z) per channel (boundary class (2,2,2)).
  const bool interior = x0 >= 2 && x0 + 4 <= G - 2 && y0 >= 2 && y0 + 8 <= G - 2 && z0 >= 2 && z0 + 8 <= G - 2;
  if (interior && threadIdx.x >= 128 && threadIdx.x < 128 + COUT) {
    const int c = threadIdx.x - 128;
    const float step = 2.0f / (float)(G - 1);
    const float2* t = reinterpret_cast<const float2*>(tables) + (long long)(62 * 3) * COUT;   // class (2,2,2) = 62
    const float2 sx = __ldg(t + c), sy = __ldg(t + COUT + c), sz = __ldg(t + 2 * COUT + c);
    s_lin[c] = make_float4(bias[c] + step * (sx.y + sy.y + sz.y), sx.x, sy.x, sz.x);
  }
  __syncthreads();

  const int lz = threadIdx.x & 7, ly = (threadIdx.x >> 3) & 7, lx = threadIdx.x >> 6;
  const int x = x0 + lx, y = y0 + ly, z = z0 + lz;
  float acc[COUT];
  if (interior) {
    const float lx_ = lin[x], ly_ = lin[y], lz_ = lin[z];
#pragma unroll
    for (int c = 0; c < COUT; c++) {
      const float4 k = s_lin[c];
      acc[c] = fmaf(lx_, k.y, fmaf(ly_, k.z, fmaf(lz_, k.w, k.x)));
    }
  } else {
    // CoordConv channels: per-class affine form
#pragma unroll
    for (int c = 0; c < COUT; c++) acc[c] = bias[c];
    const int p[3] = {x, y, z};
    int cls = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int c = p[a] == 0 ? 0 : (p[a] == 1 ? 1 : (p[a] == G - 2 ? 3 : (p[a] == G - 1 ? 4 : 2)));
      cls = cls * 5 + c;
    }
    const float step = 2.0f / (float)(G - 1);
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float base = lin[p[a]];
      const float2* t = reinterpret_cast<const float2*>(tables) + ((long long)cls * 3 + a) * COUT;
#pragma unroll
      for (int c = 0; c < COUT; c++) {
        const float2 sv = __ldg(t + c);
        acc[c] = fmaf(base, sv.x, fmaf(step, sv.y, acc[c]));
      }
    }
  }
  // occupancy channel: only non-zero taps.  One 5-bit row mask per (kx, ky) from the bit-halo replaces five
  // per-tap loads + tests; a tap's FMA block still runs once per warp when any lane has a hit.
  if (block_any) {
    for (int kx = 0; kx < 5; kx++)
#pragma unroll
      for (int ky = 0; ky < 5; ky++) {
        const uint32_t m = (hbits[lx + kx][ly + ky] >> lz) & 31u;
        if (__any_sync(0xffffffffu, m != 0)) {
#pragma unroll
          for (int kz = 0; kz < 5; kz++) {
            if ((m >> kz) & 1u) {
              const float v = halo[lx + kx][ly + ky][lz + kz];
              const float4* wr = reinterpret_cast<const float4*>(s_w + ((kx * 5 + ky) * 5 + kz) * COUT);
#pragma unroll
              for (int c4 = 0; c4 < COUT / 4; c4++) {
                const float4 w4 = wr[c4];
                acc[c4 * 4 + 0] = fmaf(v, w4.x, acc[c4 * 4 + 0]);
                acc[c4 * 4 + 1] = fmaf(v, w4.y, acc[c4 * 4 + 1]);
                acc[c4 * 4 + 2] = fmaf(v, w4.z, acc[c4 * 4 + 2]);
                acc[c4 * 4 + 3] = fmaf(v, w4.w, acc[c4 * 4 + 3]);
              }
            }
          }
        }
      }
  }
  // Stage the block's 256 voxel rows in shared memory (reusing the weight buffer), then write them out with
  // fully coalesced 16-byte stores: per-thread row stores touch a different cache line in every lane.
  __syncthreads();
  half8* stage = reinterpret_cast<half8*>(s_w);            // [256 voxels][COUT/8] 16-byte chunks, XOR-swizzled
  constexpr int CH = COUT / 8;
#pragma unroll
  for (int c8 = 0; c8 < CH; c8++) stage[threadIdx.x * CH + (c8 ^ (threadIdx.x & (CH - 1)))] = nm_pack8(acc + c8 * 8);
  __syncthreads();
  // the tile is 32 (x, y) rows of 8 consecutive z voxels = 8 * COUT * 2 contiguous bytes each
  constexpr int ROW_CHUNKS = 8 * CH;
  for (int i = threadIdx.x; i < 256 * CH; i += 256) {
    const int r = i / ROW_CHUNKS, pos = i % ROW_CHUNKS;     // r = lx * 8 + ly
    const int v = r * 8 + pos / CH, c8 = pos % CH;          // voxel index inside the block, channel chunk
    half8* dst = reinterpret_cast<half8*>(
        out + (((long long)n * G + x0 + (r >> 3)) * G * G + (long long)(y0 + (r & 7)) * G + z0) * COUT);
    dst[pos] = stage[v * CH + (c8 ^ (v & (CH - 1)))];
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
  return ((size_t)125 * 3 * Cout * 2 + (size_t)125 * Cout) * sizeof(float);
}

extern "C" int nm_first_conv_prepare(const float* weight, int Cout, void* tables, void* stream) {
  NM_CHECK_ARG(weight && tables, "nm_first_conv_prepare: null pointer");
  NM_CHECK_ARG(Cout == 32 || Cout == 64, "nm_first_conv_prepare: Cout=%d unsupported", Cout);
  float* t = (float*)tables;
  first_conv_prep_kernel<<<125, 64, 0, (cudaStream_t)stream>>>(weight, Cout, t, t + (size_t)125 * 3 * Cout * 2);
  NM_CHECK_LAUNCH("first_conv_prepare");
  return NM_OK;
}

extern "C" int nm_first_conv_k5(const float* occ, const void* tables, const float* bias, const float* linspace,
                                int n, int G, int Cout, void* out, void* stream) {
  NM_CHECK_ARG(occ && tables && bias && linspace && out, "nm_first_conv_k5: null pointer");
  NM_CHECK_ARG(G % 8 == 0 && G >= 8, "nm_first_conv_k5: grid %d must be a multiple of 8", G);
  if (n == 0) return NM_OK;
  const float* t = (const float*)tables;
  const float* wocc = t + (size_t)125 * 3 * Cout * 2;
  dim3 grid((G / 4) * (G / 8) * (G / 8), n);
  cudaStream_t st = (cudaStream_t)stream;
  if (Cout == 32) {
    first_conv_kernel<32><<<grid, 256, 128 * 32 * sizeof(float), st>>>(occ, wocc, t, bias, linspace, G, (act_t*)out);
  } else if (Cout == 64) {
    first_conv_kernel<64><<<grid, 256, 128 * 64 * sizeof(float), st>>>(occ, wocc, t, bias, linspace, G, (act_t*)out);
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
