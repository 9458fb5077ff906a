// Point-cloud voxelization (reference: utils/dataset_utils.py:21-31) and the
// clip-global normalisation in front of it (utils/dataset_utils.py:9-19).
//
// HBM-bound scatter: read N*3 coordinates per frame, write G^3 fp32 per frame.
// The reference store is an idempotent `grid[idx] = 1.0`, so duplicates need no
// atomics; a warp-level __match_any_sync on the linear cell index lets one lane
// per distinct cell issue the store (config #5: up to ~200 points per cell).
#include "common.cuh"
#include "../../include/nm_b200.h"   // the definitions below must match the public declarations

namespace {

// One thread per point.  Coordinates are staged through shared memory so that the
// global loads are fully coalesced (a warp reads 96 consecutive scalars).
template <typename T>
__global__ void __launch_bounds__(256)
voxelize_kernel(const T* __restrict__ pts, int n_pts, int G, float* __restrict__ grid,
                int* __restrict__ err_flag) {
  __shared__ T stage[256 * 3];
  const int frame = blockIdx.y;
  const long long base = (long long)frame * n_pts;
  const int first = blockIdx.x * 256;
  const int n_here = min(256, n_pts - first);
  const T* src = pts + (base + first) * 3;
  for (int i = threadIdx.x; i < n_here * 3; i += 256) stage[i] = src[i];
  __syncthreads();
  const bool live = (int)threadIdx.x < n_here;
  int lin = -1;
  if (live) {
    // float64 quotient by (step + 1e-5), truncation toward zero: dataset_utils.py:28
    const double denom = 2.0 / (double)G + 1e-5;
    const int ix = (int)__ddiv_rn((double)stage[threadIdx.x * 3 + 0] + 1.0, denom);
    const int iy = (int)__ddiv_rn((double)stage[threadIdx.x * 3 + 1] + 1.0, denom);
    const int iz = (int)__ddiv_rn((double)stage[threadIdx.x * 3 + 2] + 1.0, denom);
    if ((unsigned)ix < (unsigned)G && (unsigned)iy < (unsigned)G && (unsigned)iz < (unsigned)G) {
      lin = (ix * G + iy) * G + iz;
    } else if (err_flag) {
      // numpy would wrap a negative index / raise on >= G; the drop-in reports it
      atomicOr(err_flag, 1);
    }
  }
  // warp-aggregated store: one lane per distinct cell
  const unsigned peers = __match_any_sync(0xffffffffu, lin);
  const int leader = __ffs(peers) - 1;
  if (lin >= 0 && (int)(threadIdx.x & 31) == leader)
    grid[(long long)frame * G * G * G + lin] = 1.0f;
}

// ---- clip-global min/max (episodic_normalization, dataset_utils.py:11-12) ----
// Stage 1: per-block partial min/max of each coordinate; stage 2 folded into the
// normalise kernel's prologue (it re-reduces the <= 1024 partials of its clip).
__global__ void __launch_bounds__(256)
clip_minmax_kernel(const float* __restrict__ pts, long long n_per_clip, float* __restrict__ partial,
                   int blocks_per_clip) {
  const int clip = blockIdx.y;
  const float* p = pts + (long long)clip * n_per_clip * 3;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  // scalar index i maps to coordinate i % 3; stride is a multiple of 3 so each
  // thread always sees the same coordinate phase pattern
  const long long total = n_per_clip * 3;
  for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 3; i < total;
       i += (long long)blocks_per_clip * 256 * 3) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const float v = p[i + c];
      lo[c] = fminf(lo[c], v);
      hi[c] = fmaxf(hi[c], v);
    }
  }
  __shared__ float s_lo[8][3], s_hi[8][3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      s_lo[threadIdx.x >> 5][c] = lo[c];
      s_hi[threadIdx.x >> 5][c] = hi[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float a = s_lo[0][threadIdx.x], b = s_hi[0][threadIdx.x];
    for (int w = 1; w < 8; w++) {
      a = fminf(a, s_lo[w][threadIdx.x]);
      b = fmaxf(b, s_hi[w][threadIdx.x]);
    }
    float* out = partial + ((long long)clip * blocks_per_clip + blockIdx.x) * 6;
    out[threadIdx.x] = a;
    out[3 + threadIdx.x] = b;
  }
}

// Normalise + voxelize in one pass over raw fp32 points of one clip.
// fp32 op order of the reference: ((x - bmin) * scale / (blen + 1e-5)) * 2 - 1, then the
// float64 promotion `+ [x_trans, 0, z_trans]` (dataset_utils.py:15), then voxelize's
// float64 quotient.  No FMA contraction: every step is an explicit _rn intrinsic.
__global__ void __launch_bounds__(256)
normalize_voxelize_kernel(const float* __restrict__ pts, int T, int n_pts, int G, float scale,
                          double x_trans, double z_trans, const float* __restrict__ partial,
                          int blocks_per_clip, float* __restrict__ grid, float* __restrict__ bounds_out,
                          int* __restrict__ err_flag) {
  __shared__ float s_b[6];
  __shared__ float stage[256 * 3];
  const int clip = blockIdx.z, t = blockIdx.y;
  if (threadIdx.x < 6) {
    const bool is_hi = threadIdx.x >= 3;
    float v = is_hi ? -INFINITY : INFINITY;
    for (int b = 0; b < blocks_per_clip; b++) {
      const float q = partial[((long long)clip * blocks_per_clip + b) * 6 + threadIdx.x];
      v = is_hi ? fmaxf(v, q) : fminf(v, q);
    }
    s_b[threadIdx.x] = v;
    if (bounds_out && blockIdx.x == 0 && t == 0) bounds_out[clip * 6 + threadIdx.x] = v;
  }
  const long long frame = (long long)clip * T + t;
  const int first = blockIdx.x * 256;
  const int n_here = min(256, n_pts - first);
  const float* src = pts + (frame * n_pts + first) * 3;
  for (int i = threadIdx.x; i < n_here * 3; i += 256) stage[i] = src[i];
  __syncthreads();
  int lin = -1;
  if ((int)threadIdx.x < n_here) {
    const float ext = fmaxf(fmaxf(__fsub_rn(s_b[3], s_b[0]), __fsub_rn(s_b[4], s_b[1])),
                            __fsub_rn(s_b[5], s_b[2]));
    const float den = __fadd_rn(ext, 1e-5f);       // numpy>=2: stays float32 for float32 input
    const double denom = 2.0 / (double)G + 1e-5;
    int idx[3];
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float v = __fsub_rn(stage[threadIdx.x * 3 + c], s_b[c]);
      v = __fmul_rn(v, scale);
      v = __fdiv_rn(v, den);
      v = __fsub_rn(__fmul_rn(v, 2.0f), 1.0f);
      const double d = (double)v + (c == 0 ? x_trans : (c == 2 ? z_trans : 0.0));
      idx[c] = (int)__ddiv_rn(d + 1.0, denom);
      ok = ok && ((unsigned)idx[c] < (unsigned)G);
    }
    if (ok) lin = (idx[0] * G + idx[1]) * G + idx[2];
    else if (err_flag) atomicOr(err_flag, 1);
  }
  const unsigned peers = __match_any_sync(0xffffffffu, lin);
  const int leader = __ffs(peers) - 1;
  if (lin >= 0 && (int)(threadIdx.x & 31) == leader) grid[frame * G * G * G + lin] = 1.0f;
}

}  // namespace

extern "C" int nm_voxelize(const void* points, int points_are_f64, int n_frames, int n_points, int grid_size,
                           float* grid_out, int* err_flag, void* stream) {
  NM_CHECK_ARG(n_frames >= 0 && n_points >= 0 && grid_size > 0 && grid_size <= 1024,
               "nm_voxelize: bad sizes F=%d N=%d G=%d", n_frames, n_points, grid_size);
  NM_CHECK_ARG(grid_out && (points || n_points == 0 || n_frames == 0), "nm_voxelize: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cells = (size_t)grid_size * grid_size * grid_size;
  if (n_frames == 0) return NM_OK;
  NM_CHECK_CUDA(cudaMemsetAsync(grid_out, 0, cells * n_frames * sizeof(float), st));
  if (n_points == 0) return NM_OK;
  dim3 grid(nm_cdiv(n_points, 256), n_frames);
  if (points_are_f64)
    voxelize_kernel<double><<<grid, 256, 0, st>>>((const double*)points, n_points, grid_size, grid_out, err_flag);
  else
    voxelize_kernel<float><<<grid, 256, 0, st>>>((const float*)points, n_points, grid_size, grid_out, err_flag);
  NM_CHECK_LAUNCH("nm_voxelize");
  return NM_OK;
}

extern "C" size_t nm_normalize_voxelize_workspace_bytes(int n_clips) {
  return (size_t)n_clips * 64 * 6 * sizeof(float);
}

extern "C" int nm_normalize_voxelize(const float* raw_points, int n_clips, int T, int n_points, int grid_size,
                                     float scale, double x_trans, double z_trans, float* grid_out,
                                     float* bounds_out, void* workspace, int* err_flag, void* stream) {
  NM_CHECK_ARG(raw_points && grid_out && workspace, "nm_normalize_voxelize: null pointer");
  NM_CHECK_ARG(n_clips > 0 && T > 0 && n_points > 0 && grid_size > 0, "nm_normalize_voxelize: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int bpc = 64;
  const size_t cells = (size_t)grid_size * grid_size * grid_size;
  NM_CHECK_CUDA(cudaMemsetAsync(grid_out, 0, cells * n_clips * T * sizeof(float), st));
  clip_minmax_kernel<<<dim3(bpc, n_clips), 256, 0, st>>>(raw_points, (long long)T * n_points,
                                                         (float*)workspace, bpc);
  NM_CHECK_LAUNCH("nm_normalize_voxelize(minmax)");
  normalize_voxelize_kernel<<<dim3(nm_cdiv(n_points, 256), T, n_clips), 256, 0, st>>>(
      raw_points, T, n_points, grid_size, scale, x_trans, z_trans, (const float*)workspace, bpc, grid_out,
      bounds_out, err_flag);
  NM_CHECK_LAUNCH("nm_normalize_voxelize");
  return NM_OK;
}
