// Consumers of the path's outputs: the voxel chamfer metric and the semantic score of
// utils/eval_utils.py, and the retarget post-processing of vis_retarget.py (skin weights, forward kinematics of
// the retargeted skeleton, linear blend skinning).  All HBM/latency-bound integer / fp32 work on CUDA cores.
#include "common.cuh"
#include "../../include/nm_b200.h"

// ----------------------------------------------------------------------------------------------------------
// voxel chamfer (utils/eval_utils.py:29-56)
//   pass 1: per frame and per volume, ordered compaction of the occupied voxels into packed coordinates
//           (reads each grid once: 2 x G^3 x 4 B per frame; recon is binarised in place like the reference, :37-38)
//   pass 2: brute-force nearest neighbour in INDEX space: squared distances are small integers, so the per-frame
//           sums are exact 64-bit integers whatever the summation order (deterministic atomics)
//   pass 3: chamfer = (2/(G-1))^2 * (sum_gt / n_gt + sum_recon / n_recon)
// ----------------------------------------------------------------------------------------------------------
static constexpr int kCompactThreads = 1024;

__global__ void __launch_bounds__(kCompactThreads)
vox_compact_kernel(const float* __restrict__ gt, float* __restrict__ recon, int G, long long S, int binarize,
                   uint32_t* __restrict__ lists, int* __restrict__ counts) {
  const int frame = blockIdx.x, which = blockIdx.y;
  float* vol = which == 0 ? const_cast<float*>(gt) + (long long)frame * S : recon + (long long)frame * S;
  uint32_t* list = lists + ((long long)frame * 2 + which) * S;
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (long long start = 0; start < S; start += (long long)kCompactThreads * 4) {
    const long long e = start + (long long)threadIdx.x * 4;
    uint32_t flags = 0;
    if (e < S) {                                                    // S = G^3 with G % 4 == 0 (checked by the host)
      float4 v = *reinterpret_cast<const float4*>(vol + e);
      if (which == 0) {
        flags = (v.x != 0.f) | ((v.y != 0.f) << 1) | ((v.z != 0.f) << 2) | ((v.w != 0.f) << 3);
      } else {
        flags = (v.x >= 0.5f) | ((v.y >= 0.5f) << 1) | ((v.z >= 0.5f) << 2) | ((v.w >= 0.5f) << 3);
        if (binarize) {
          // recon[recon >= 0.5] = 1; recon[recon < 0.5] = 0 (NaN stays NaN and counts as occupied, like torch.where)
          float4 o;
          o.x = v.x >= 0.5f ? 1.f : (v.x < 0.5f ? 0.f : v.x);
          o.y = v.y >= 0.5f ? 1.f : (v.y < 0.5f ? 0.f : v.y);
          o.z = v.z >= 0.5f ? 1.f : (v.z < 0.5f ? 0.f : v.z);
          o.w = v.w >= 0.5f ? 1.f : (v.w < 0.5f ? 0.f : v.w);
          *reinterpret_cast<float4*>(vol + e) = o;
        }
        flags |= (v.x != v.x) | ((v.y != v.y) << 1) | ((v.z != v.z) << 2) | ((v.w != v.w) << 3);
      }
    }
    const int cnt = __popc(flags);
    int incl = cnt;                                                 // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
    {
      const int wt = warp_tot[lane];                                // 32 warps: every warp scans the warp totals
      int wi = wt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      wbase = __shfl_sync(0xffffffffu, wi - wt, warp);
      total = __shfl_sync(0xffffffffu, wi, 31);
    }
    int pos = base_s + wbase + incl - cnt;
    if (flags) {
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (flags & (1u << q)) {
          const long long idx = e + q;
          const int k = (int)(idx % G), j = (int)((idx / G) % G), i = (int)(idx / ((long long)G * G));
          list[pos++] = ((uint32_t)i << 20) | ((uint32_t)j << 10) | (uint32_t)k;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) base_s += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[frame * 2 + which] = base_s;
}

static constexpr int kNearThreads = 256;
static constexpr int kNearTile = 2048;

__global__ void __launch_bounds__(kNearThreads)
vox_nearest_kernel(const uint32_t* __restrict__ lists, const int* __restrict__ counts, long long S,
                   unsigned long long* __restrict__ sums) {
  const int frame = blockIdx.z, dir = blockIdx.y;                   // dir 0: gt -> recon, dir 1: recon -> gt
  const int nq = counts[frame * 2 + dir], nt = counts[frame * 2 + (dir ^ 1)];
  if (nq == 0 || nt == 0) return;
  const uint32_t* q = lists + ((long long)frame * 2 + dir) * S;
  const uint32_t* t = lists + ((long long)frame * 2 + (dir ^ 1)) * S;
  __shared__ short4 tile[kNearTile];
  __shared__ unsigned long long block_sum;
  if (threadIdx.x == 0) block_sum = 0ull;
  unsigned long long acc = 0ull;
  for (int q0 = blockIdx.x * kNearThreads; q0 < nq; q0 += gridDim.x * kNearThreads) {
    const int qi = q0 + threadIdx.x;
    const bool live = qi < nq;
    const uint32_t pq = live ? q[qi] : 0u;
    const int qx = pq >> 20, qy = (pq >> 10) & 1023, qz = pq & 1023;
    int best = 0x7fffffff;
    for (int t0 = 0; t0 < nt; t0 += kNearTile) {
      const int m = min(kNearTile, nt - t0);
      __syncthreads();
      for (int i = threadIdx.x; i < m; i += kNearThreads) {
        const uint32_t p = t[t0 + i];
        tile[i] = make_short4((short)(p >> 20), (short)((p >> 10) & 1023), (short)(p & 1023), 0);
      }
      __syncthreads();
      if (live) {
#pragma unroll 8
        for (int i = 0; i < m; i++) {
          const short4 p = tile[i];
          const int dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
          best = min(best, dx * dx + dy * dy + dz * dz);
        }
      }
    }
    if (live) acc += (unsigned long long)best;
  }
  // block reduction (integers: exact and order-independent)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&block_sum, acc);
  __syncthreads();
  if (threadIdx.x == 0 && block_sum) atomicAdd(&sums[frame * 2 + dir], block_sum);
}

__global__ void vox_chamfer_finalize_kernel(const int* __restrict__ counts, const unsigned long long* __restrict__ sums,
                                            int n, int G, float* __restrict__ out, int* __restrict__ occupied_out,
                                            int* __restrict__ err_flag) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const int c0 = counts[2 * f], c1 = counts[2 * f + 1];
  if (occupied_out) { occupied_out[2 * f] = c0; occupied_out[2 * f + 1] = c1; }
  if (c0 == 0 || c1 == 0) {                                         // torch: min over an empty dimension raises
    out[f] = __int_as_float(0x7fc00000);
    if (err_flag) atomicOr(err_flag, 1);
    return;
  }
  const double s = 2.0 / (double)(G - 1);                           // coords = idx / ((G-1)/2) - 1  (:43-44)
  out[f] = (float)(s * s * ((double)sums[2 * f] / c0 + (double)sums[2 * f + 1] / c1));
}

static size_t chamfer_metric_ws(int n, int G) {
  const size_t S = (size_t)G * G * G;
  return (size_t)n * 2 * sizeof(unsigned long long) + (size_t)n * 2 * sizeof(int) + (size_t)n * 2 * S * sizeof(uint32_t) + 64;
}

extern "C" size_t nm_voxel_chamfer_workspace_bytes(int n, int G) { return chamfer_metric_ws(n, G); }

extern "C" int nm_voxel_chamfer(const float* gt, float* recon, int n, int G, int binarize_recon, float* out,
                                int* occupied_out, int* err_flag, void* workspace, void* stream) {
  NM_CHECK_ARG(gt && recon && out && workspace, "nm_voxel_chamfer: null pointer");
  NM_CHECK_ARG(n > 0 && G >= 4 && G <= 1024 && G % 4 == 0, "nm_voxel_chamfer: need n > 0, 4 <= G <= 1024, G %% 4 == 0 (got n=%d G=%d)", n, G);
  NM_CHECK_ARG(n <= 65535, "nm_voxel_chamfer: at most 65535 frames per call (got %d)", n);
  cudaStream_t st = (cudaStream_t)stream;
  const long long S = (long long)G * G * G;
  unsigned long long* sums = reinterpret_cast<unsigned long long*>(workspace);
  int* counts = reinterpret_cast<int*>(sums + (size_t)n * 2);
  uintptr_t lp = (reinterpret_cast<uintptr_t>(counts + (size_t)n * 2) + 15) & ~(uintptr_t)15;
  uint32_t* lists = reinterpret_cast<uint32_t*>(lp);
  NM_CHECK_CUDA(cudaMemsetAsync(sums, 0, (size_t)n * 2 * sizeof(unsigned long long), st));
  vox_compact_kernel<<<dim3(n, 2), kCompactThreads, 0, st>>>(gt, recon, G, S, binarize_recon, lists, counts);
  NM_CHECK_LAUNCH("vox_compact_kernel");
  vox_nearest_kernel<<<dim3(8, 2, n), kNearThreads, 0, st>>>(lists, counts, S, sums);
  NM_CHECK_LAUNCH("vox_nearest_kernel");
  vox_chamfer_finalize_kernel<<<nm_cdiv(n, 128), 128, 0, st>>>(counts, sums, n, G, out, occupied_out, err_flag);
  NM_CHECK_LAUNCH("vox_chamfer_finalize_kernel");
  return NM_OK;
}

// ----------------------------------------------------------------------------------------------------------
// semantic score (utils/eval_utils.py:60-90): nearest detected keypoint of every ground-truth joint
// ----------------------------------------------------------------------------------------------------------
__global__ void semantic_mask_kernel(float* __restrict__ kypt, int rows, float threshold) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float4* p = reinterpret_cast<float4*>(kypt) + r;
  if ((*p).w < threshold) *p = make_float4(1e4f, 1e4f, 1e4f, 1.f);   // :68-69
}

__global__ void semantic_nearest_kernel(const float* __restrict__ kypt, const float* __restrict__ gt, int F, int K, int Kgt,
                                        long long* __restrict__ idx_out, int* __restrict__ hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F * Kgt) return;
  const int f = i / Kgt, kg = i % Kgt;
  const float gx = gt[3 * i], gy = gt[3 * i + 1], gz = gt[3 * i + 2];
  const float4* det = reinterpret_cast<const float4*>(kypt) + (long long)f * K;
  float best = 0.f;
  int arg = 0;
  for (int k = 0; k < K; k++) {
    const float4 d = det[k];
    const float dx = gx - d.x, dy = gy - d.y, dz = gz - d.z;
    const float dist = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    if (k == 0 || dist < best) { best = dist; arg = k; }              // first minimum, like torch.min(...).indices
  }
  idx_out[i] = arg;
  atomicAdd(&hist[kg * K + arg], 1);                                   // one_hot[closest].sum(0) (:83-84)
}

extern "C" int nm_semantic_nearest(float* keypoints, const float* gt_keypoints, int F, int K, int Kgt, float threshold,
                                   long long* idx_out, int* hist, void* stream) {
  NM_CHECK_ARG(keypoints && gt_keypoints && idx_out && hist, "nm_semantic_nearest: null pointer");
  NM_CHECK_ARG(F > 0 && K > 0 && Kgt > 0, "nm_semantic_nearest: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  NM_CHECK_CUDA(cudaMemsetAsync(hist, 0, (size_t)K * Kgt * sizeof(int), st));
  semantic_mask_kernel<<<nm_cdiv((long long)F * K, 256), 256, 0, st>>>(keypoints, F * K, threshold);
  NM_CHECK_LAUNCH("semantic_mask_kernel");
  semantic_nearest_kernel<<<nm_cdiv((long long)F * Kgt, 128), 128, 0, st>>>(keypoints, gt_keypoints, F, K, Kgt, idx_out, hist);
  NM_CHECK_LAUNCH("semantic_nearest_kernel");
  return NM_OK;
}

// ----------------------------------------------------------------------------------------------------------
// retarget post-processing (vis_retarget.py)
// ----------------------------------------------------------------------------------------------------------
static constexpr int kMaxJoints = 64;

__device__ __forceinline__ float dist3(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

// extract_skin_weights (vis_retarget.py:21-62): one thread per point instead of the reference's Python loop
__global__ void __launch_bounds__(256)
skin_weights_kernel(const float* __restrict__ points, int N, const float* __restrict__ keypoints, const int* __restrict__ parents,
                    int K, int root, float hardness, float threshold, float* __restrict__ skin, int* __restrict__ nearest_out,
                    int* __restrict__ err_flag) {
  __shared__ float4 kp[kMaxJoints];
  __shared__ float bone[kMaxJoints][3];
  __shared__ int par[kMaxJoints];
  __shared__ int invalid[kMaxJoints];
  if (threadIdx.x < K) {
    kp[threadIdx.x] = reinterpret_cast<const float4*>(keypoints)[threadIdx.x];
    par[threadIdx.x] = parents[threadIdx.x];
    invalid[threadIdx.x] = kp[threadIdx.x].w < threshold;              // :33
  }
  __syncthreads();
  if (threadIdx.x < K) {
    const int k = threadIdx.x;
    int p = par[k];
    if (p == k) {                                                        // :38-39
      bone[k][0] = kp[k].x; bone[k][1] = kp[k].y; bone[k][2] = kp[k].z;
    } else {
      int hops = 0;
      while (invalid[p] && hops <= K) { p = par[p]; hops++; }            // :41-42 (the reference spins forever on a cycle)
      if (hops > K && err_flag) atomicOr(err_flag, 1);
      bone[k][0] = (kp[k].x + kp[p].x) / 2.f;                            // :44
      bone[k][1] = (kp[k].y + kp[p].y) / 2.f;
      bone[k][2] = (kp[k].z + kp[p].z) / 2.f;
    }
  }
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float px = points[3 * n], py = points[3 * n + 1], pz = points[3 * n + 2];
  float best = 0.f;
  int child = 0;
  for (int k = 0; k < K; k++) {
    float d = dist3(px, py, pz, bone[k][0], bone[k][1], bone[k][2]);   // :46
    if (invalid[k] || k == root) d = 1e4f;                              // :48-49
    if (k == 0 || d < best) { best = d; child = k; }                    // :52 argmin, first minimum
  }
  const int parent = par[child];                                        // :56 (the ORIGINAL parent table)
  const float cd = expf(dist3(px, py, pz, kp[child].x, kp[child].y, kp[child].z) * hardness);
  const float pd = expf(dist3(px, py, pz, kp[parent].x, kp[parent].y, kp[parent].z) * hardness);
  float* row = skin + (long long)n * K;
  for (int k = 0; k < K; k++) row[k] = 0.f;
  row[parent] = cd / (cd + pd);                                         // :59
  row[child] = pd / (cd + pd);                                          // :60 (wins when parent == child)
  if (nearest_out) nearest_out[n] = child;
}

extern "C" int nm_skin_weights(const float* points, int N, const float* keypoints, const int* parents, int K, int root,
                               float hardness, float threshold, float* skin_out, int* nearest_out, int* err_flag,
                               void* stream) {
  NM_CHECK_ARG(points && keypoints && parents && skin_out, "nm_skin_weights: null pointer");
  NM_CHECK_ARG(N > 0 && K > 0 && K <= kMaxJoints && root >= 0 && root < K, "nm_skin_weights: need N > 0, 0 < K <= %d, 0 <= root < K", kMaxJoints);
  skin_weights_kernel<<<nm_cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(points, N, keypoints, parents, K, root, hardness,
                                                                        threshold, skin_out, nearest_out, err_flag);
  NM_CHECK_LAUNCH("skin_weights_kernel");
  return NM_OK;
}

// forward kinematics of the retargeted skeleton (vis_retarget.py:279-287): pos[root] = root_pos[t];
// pos[idx] = R[t, idx] @ offset[idx] + pos[parents[idx]] in priority order; clip(-1, 1) afterwards (:300)
__global__ void retarget_fk_kernel(const float* __restrict__ R, const float* __restrict__ offset, const float* __restrict__ root_pos,
                                   const int* __restrict__ order, const int* __restrict__ parents, int T, int K, int clip,
                                   float* __restrict__ pos) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float* P = pos + (long long)t * K * 3;
  const float* Rt = R + (long long)t * K * 9;
  const int root = order[0];
  for (int k = 0; k < K; k++) { P[3 * k] = 0.f; P[3 * k + 1] = 0.f; P[3 * k + 2] = 0.f; }
  P[3 * root] = root_pos[3 * t]; P[3 * root + 1] = root_pos[3 * t + 1]; P[3 * root + 2] = root_pos[3 * t + 2];
  for (int i = 1; i < K; i++) {
    const int idx = order[i], p = parents[idx];
    const float* M = Rt + idx * 9;
    const float ox = offset[3 * idx], oy = offset[3 * idx + 1], oz = offset[3 * idx + 2];
    for (int r = 0; r < 3; r++)
      P[3 * idx + r] = __fmaf_rn(M[3 * r + 2], oz, __fmaf_rn(M[3 * r + 1], oy, M[3 * r] * ox)) + P[3 * p + r];
  }
  if (clip)
    for (int i = 0; i < 3 * K; i++) P[i] = fminf(fmaxf(P[i], -1.f), 1.f);
}

extern "C" int nm_retarget_fk(const float* R, const float* offset, const float* root_pos, const int* order, const int* parents,
                              int T, int K, int clip, float* pos_out, void* stream) {
  NM_CHECK_ARG(R && offset && root_pos && order && parents && pos_out, "nm_retarget_fk: null pointer");
  NM_CHECK_ARG(T > 0 && K > 0, "nm_retarget_fk: empty input");
  retarget_fk_kernel<<<nm_cdiv(T, 64), 64, 0, (cudaStream_t)stream>>>(R, offset, root_pos, order, parents, T, K, clip, pos_out);
  NM_CHECK_LAUNCH("retarget_fk_kernel");
  return NM_OK;
}

// linear blend skinning (vis_retarget.py:263-270, 315-322):
//   local[n,k] = R_inv[k] @ (p[n] - joint[k])   (R_inv == NULL: identity)
//   out[t,n]   = sum_k skin[n,k] * (T[t,k,:3,:3] @ local[n,k] + T[t,k,:3,3])
__global__ void __launch_bounds__(256)
lbs_kernel(const float* __restrict__ points, int N, const float* __restrict__ joints, const float* __restrict__ R_inv,
           const float* __restrict__ T3x4, const float* __restrict__ skin, int K, float* __restrict__ out) {
  __shared__ float sT[kMaxJoints][12];
  __shared__ float sR[kMaxJoints][9];
  __shared__ float sJ[kMaxJoints][3];
  const int t = blockIdx.y;
  for (int i = threadIdx.x; i < K * 12; i += blockDim.x) sT[i / 12][i % 12] = T3x4[(long long)t * K * 12 + i];
  for (int i = threadIdx.x; i < K * 9; i += blockDim.x) sR[i / 9][i % 9] = R_inv ? R_inv[i] : ((i % 9) % 4 == 0 ? 1.f : 0.f);
  for (int i = threadIdx.x; i < K * 3; i += blockDim.x) sJ[i / 3][i % 3] = joints[i];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float px = points[3 * n], py = points[3 * n + 1], pz = points[3 * n + 2];
  float ax = 0.f, ay = 0.f, az = 0.f;
  const float* w = skin + (long long)n * K;
  for (int k = 0; k < K; k++) {
    const float wk = w[k];
    if (wk == 0.f) continue;
    const float dx = px - sJ[k][0], dy = py - sJ[k][1], dz = pz - sJ[k][2];
    const float lx = sR[k][0] * dx + sR[k][1] * dy + sR[k][2] * dz;
    const float ly = sR[k][3] * dx + sR[k][4] * dy + sR[k][5] * dz;
    const float lz = sR[k][6] * dx + sR[k][7] * dy + sR[k][8] * dz;
    ax += wk * (sT[k][0] * lx + sT[k][1] * ly + sT[k][2] * lz + sT[k][3]);
    ay += wk * (sT[k][4] * lx + sT[k][5] * ly + sT[k][6] * lz + sT[k][7]);
    az += wk * (sT[k][8] * lx + sT[k][9] * ly + sT[k][10] * lz + sT[k][11]);
  }
  float* o = out + ((long long)t * N + n) * 3;
  o[0] = ax; o[1] = ay; o[2] = az;
}

extern "C" int nm_linear_blend_skinning(const float* points, int N, const float* joints, const float* R_inv, const float* T3x4,
                                        const float* skin, int T, int K, float* out, void* stream) {
  NM_CHECK_ARG(points && joints && T3x4 && skin && out, "nm_linear_blend_skinning: null pointer");
  NM_CHECK_ARG(N > 0 && T > 0 && T <= 65535 && K > 0 && K <= kMaxJoints, "nm_linear_blend_skinning: need N, T > 0, T <= 65535, 0 < K <= %d", kMaxJoints);
  lbs_kernel<<<dim3(nm_cdiv(N, 256), T), 256, 0, (cudaStream_t)stream>>>(points, N, joints, R_inv, T3x4, skin, K, out);
  NM_CHECK_LAUNCH("lbs_kernel");
  return NM_OK;
}
