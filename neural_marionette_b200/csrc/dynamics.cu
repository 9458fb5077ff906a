// One fused kernel per HSVRNN time step (reference: model/hsvrnn_bvh.py:89-135 / 171-225,
// extract_kypt_from_latent_and_state :255-286, utils/geo_utils.py:3-78, nn.GRUCell :57-58).
//
// The reference spends a step in ~4.7k ATen launches and ~750 host syncs; the arithmetic is
// ~6 MFLOP/sample over 6.1 MB of fp32 weights, i.e. latency / L2-bound.  Here a CTA owns NB
// batch elements and walks the whole step in shared memory:
//   prior MLP (optional) -> posterior MLP -> z_s = mu + sigma*eps_s -> decoders for all S
//   samples -> 6-D rotations + forward kinematics down the skeleton -> nearest-sample pick
//   -> GRU cell.
// All matrices are stored transposed ([in][out], prepared once by the host) so that a thread
// owns an output column: weight reads are coalesced across the CTA, activations are
// shared-memory broadcasts, no shuffles.  fp32 throughout (the nearest-sample pick is a
// discrete decision; keep it as close to the reference arithmetic as possible).
#include "common.cuh"
#include "../../include/nm_b200.h"
#include <stdlib.h>

namespace {

constexpr int kThreads = 256;
constexpr int H = 512, Z = 128, HID = 128;
constexpr int KP_MAX = 24;
constexpr int VC = 8;  // vectors per register chunk

__device__ __forceinline__ float softplus_d(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_d(float x) { return 1.0f / (1.0f + expf(-x)); }

// out[v][c] = act(bias[c] + sum_i WT[i][col0 + c] * x[v][i]) for v < nvec, c < ncols.
// x rows beyond nvec (up to the next multiple of VC) must be readable (zero padded).
// act: 0 none, 1 leaky relu, 2 tanh.  `add` (optional, [v / add_div][c]) is added before act.
__device__ void cols_matvec(const float* __restrict__ WT, int ldw, int col0, int ncols, int I,
                            const float* x, int xstride, int nvec, const float* __restrict__ bias,
                            const float* add, int add_stride, int add_div, int act, float* out, int ostride) {
  for (int c = threadIdx.x; c < ncols; c += kThreads) {
    const float* w = WT + col0 + c;
    for (int v0 = 0; v0 < nvec; v0 += VC) {
      float acc[VC];
#pragma unroll
      for (int u = 0; u < VC; u++) acc[u] = 0.f;
#pragma unroll 4
      for (int i = 0; i < I; i++) {
        const float wi = __ldg(w + (long long)i * ldw);
#pragma unroll
        for (int u = 0; u < VC; u++) acc[u] = fmaf(wi, x[(v0 + u) * xstride + i], acc[u]);
      }
#pragma unroll
      for (int u = 0; u < VC; u++) {
        const int v = v0 + u;
        if (v < nvec) {
          float r = acc[u] + (bias ? bias[c] : 0.f);
          if (add) r += add[(v / add_div) * add_stride + c];
          if (act == 1) r = nm_lrelu(r);
          else if (act == 2) r = tanhf(r);
          out[v * ostride + c] = r;
        }
      }
    }
  }
}

// 6-D -> rotation (geo_utils.py:56-78): x = a/|a|, z = (x × b)/|x × b|, y = z × x; columns (x, y, z)
__device__ __forceinline__ void rot6d(const float* p, float* R /* row-major 3x3 */) {
  const float na = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) + 1e-10f;
  const float x0 = p[0] / na, x1 = p[1] / na, x2 = p[2] / na;
  float z0 = x1 * p[5] - x2 * p[4], z1 = x2 * p[3] - x0 * p[5], z2 = x0 * p[4] - x1 * p[3];
  const float nz = sqrtf(z0 * z0 + z1 * z1 + z2 * z2) + 1e-10f;
  z0 /= nz; z1 /= nz; z2 /= nz;
  const float y0 = z1 * x2 - z2 * x1, y1 = z2 * x0 - z0 * x2, y2 = z0 * x1 - z1 * x0;
  R[0] = x0; R[1] = y0; R[2] = z0;
  R[3] = x1; R[4] = y1; R[5] = z1;
  R[6] = x2; R[7] = y2; R[8] = z2;
}

// Forward kinematics for one (batch element, sample): hsvrnn_bvh.py:264-281.
// root_raw: 3 + K tanh outputs; rot: K*6; offset: K*3; Rg scratch K*9; flat out K*4.
__device__ void fk_pose(const float* root_raw, const float* rot, const float* offset, const int* order,
                        const int* parents, int K, float* Rg, float* flat) {
  const int root = order[0];
  rot6d(rot + root * 6, Rg + root * 9);
  flat[root * 4 + 0] = root_raw[0];
  flat[root * 4 + 1] = root_raw[1];
  flat[root * 4 + 2] = root_raw[2];
  for (int j = 1; j < K; j++) {
    const int i = order[j], pa = parents[i];
    float Rl[9];
    rot6d(rot + i * 6, Rl);
    const float* Rp = Rg + pa * 9;
    float* Ri = Rg + i * 9;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) Ri[r * 3 + c] = Rp[r * 3] * Rl[c] + Rp[r * 3 + 1] * Rl[3 + c] + Rp[r * 3 + 2] * Rl[6 + c];
    const float* o = offset + i * 3;
#pragma unroll
    for (int r = 0; r < 3; r++) flat[i * 4 + r] = Ri[r * 3] * o[0] + Ri[r * 3 + 1] * o[1] + Ri[r * 3 + 2] * o[2] + flat[pa * 4 + r];
  }
  for (int i = 0; i < K; i++) flat[i * 4 + 3] = (root_raw[3 + i] + 1.0f) * 0.5f;
}

struct StepArgs {
  nm_hsvrnn_weights w;
  const float* h_in;      // (B, 512)
  const float* kp;        // (B, 4K) detected keypoints (posterior mode) or null
  const float* eps;       // (S, B, 128)
  const float* offset;    // (B, K, 3)
  const int* order;       // (K)
  const int* parents;     // (K)
  float* h_out;           // (B, 512)
  float* kp_out;          // (B, 4K)
  float* z_out;           // (B, 128) or null
  float* R_out;           // (B, K, 9) or null
  float* post_out;        // (B, 256) posterior mean | std, or null
  float* prior_out;       // (B, 256) prior mean | std, or null
  int B, K, S, posterior;
};

template <int NB>
__global__ void __launch_bounds__(kThreads)
hsvrnn_step_kernel(const StepArgs a) {
  extern __shared__ float sm[];
  const int K = a.K, S = a.S, K4 = 4 * a.K;
  const int b0 = blockIdx.x * NB;
  const int nb = min(NB, a.B - b0);
  const int nvs = nb * S;                       // live (element, sample) vectors
  const int NVS = ((NB * S + VC - 1) / VC) * VC;  // padded
  constexpr int NBP = ((NB + VC - 1) / VC) * VC;
  // ---- shared-memory carve-up (floats)
  float* s_in = sm;                         // [NBP][H + 96]  h | kp   (posterior MLP input)
  const int IN = H + K4;
  float* s_hid = s_in + NBP * (H + 96);     // [NBP][HID]
  float* s_dist = s_hid + NBP * HID;        // [NBP][2Z]  mean | std (posterior, or prior in prior mode)
  float* s_z = s_dist + NBP * 2 * Z;        // [NVS][Z]
  float* s_ah = s_z + NVS * Z;              // [NBP][2 HID]  W0[:, :H] h + b0 for (root | joint) decoders
  float* s_h2 = s_ah + NBP * 2 * HID;       // [NVS][2 HID]
  float* s_dec = s_h2 + NVS * 2 * HID;      // [NVS][176]  root(3+K) at 0, joint(6K) at 32
  float* s_flat = s_dec + NVS * 176;        // [NVS][96]
  float* s_rg = s_flat + NVS * 96;          // [NVS][K*9]
  float* s_x = s_rg + NVS * KP_MAX * 9;     // [NBP][224]  GRU input
  float* s_off = s_x + NBP * 224;           // [NB][K*3]
  int* s_tree = reinterpret_cast<int*>(s_off + NB * KP_MAX * 3);  // order[K] | parents[K] | best[NB]
  int* s_best = s_tree + 2 * KP_MAX;

  // ---- stage 0: load, zero the padding rows
  for (int i = threadIdx.x; i < NBP * (H + 96); i += kThreads) {
    const int v = i / (H + 96), j = i % (H + 96);
    float val = 0.f;
    if (v < nb) {
      if (j < H) val = a.h_in[(long long)(b0 + v) * H + j];
      else if (a.posterior && j - H < K4) val = a.kp[(long long)(b0 + v) * K4 + (j - H)];
    }
    s_in[i] = val;
  }
  for (int i = threadIdx.x; i < NVS * Z; i += kThreads) s_z[i] = 0.f;
  for (int i = threadIdx.x; i < NBP * HID; i += kThreads) s_hid[i] = 0.f;
  for (int i = threadIdx.x; i < NBP * 224; i += kThreads) s_x[i] = 0.f;
  for (int i = threadIdx.x; i < nb * K * 3; i += kThreads) s_off[(i / (K * 3)) * KP_MAX * 3 + i % (K * 3)] = a.offset[(long long)b0 * K * 3 + i];
  for (int i = threadIdx.x; i < K; i += kThreads) {
    s_tree[i] = a.order[i];
    s_tree[KP_MAX + i] = a.parents[i];
  }
  __syncthreads();

  // ---- prior MLP (hsvrnn_bvh.py:92-96 / 210-214): 512 -> 128 -> 256
  if (a.prior_out || !a.posterior) {
    cols_matvec(a.w.prior0_wt, HID, 0, HID, H, s_in, H + 96, nb, a.w.prior0_b, nullptr, 0, 1, 1, s_hid, HID);
    __syncthreads();
    cols_matvec(a.w.prior2_wt, 2 * Z, 0, 2 * Z, HID, s_hid, HID, nb, a.w.prior2_b, nullptr, 0, 1, 0, s_dist, 2 * Z);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * Z; i += kThreads) {
      const int v = i / Z, j = i % Z;
      const float sd = softplus_d(s_dist[v * 2 * Z + Z + j]) + 1e-4f;
      s_dist[v * 2 * Z + Z + j] = sd;
    }
    __syncthreads();
    if (a.prior_out)
      for (int i = threadIdx.x; i < nb * 2 * Z; i += kThreads) a.prior_out[(long long)b0 * 2 * Z + i] = s_dist[i];
    __syncthreads();
  }
  // ---- posterior MLP (hsvrnn_bvh.py:99-104): (512 + 4K) -> 128 -> 256
  if (a.posterior) {
    cols_matvec(a.w.post0_wt, HID, 0, HID, IN, s_in, H + 96, nb, a.w.post0_b, nullptr, 0, 1, 1, s_hid, HID);
    __syncthreads();
    cols_matvec(a.w.post2_wt, 2 * Z, 0, 2 * Z, HID, s_hid, HID, nb, a.w.post2_b, nullptr, 0, 1, 0, s_dist, 2 * Z);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * Z; i += kThreads) {
      const int v = i / Z, j = i % Z;
      s_dist[v * 2 * Z + Z + j] = softplus_d(s_dist[v * 2 * Z + Z + j]) + 1e-4f;
    }
    __syncthreads();
    if (a.post_out)
      for (int i = threadIdx.x; i < nb * 2 * Z; i += kThreads) a.post_out[(long long)b0 * 2 * Z + i] = s_dist[i];
  }
  // ---- reparameterised samples z = mean + std * eps (Normal.rsample, :107 / :216); vector index v*S + s
  for (int i = threadIdx.x; i < nvs * Z; i += kThreads) {
    const int vs = i / Z, j = i % Z, v = vs / S, s = vs % S;
    const float e = a.eps[((long long)s * a.B + b0 + v) * Z + j];
    s_z[vs * Z + j] = s_dist[v * 2 * Z + j] + s_dist[v * 2 * Z + Z + j] * e;
  }
  // ---- decoders, first layer split: W0 [h; z] = W0[:, :H] h (shared by the S samples) + W0[:, H:] z
  cols_matvec(a.w.root0_wt, HID, 0, HID, H, s_in, H + 96, nb, a.w.root0_b, nullptr, 0, 1, 0, s_ah, 2 * HID);
  cols_matvec(a.w.joint0_wt, HID, 0, HID, H, s_in, H + 96, nb, a.w.joint0_b, nullptr, 0, 1, 0, s_ah + HID, 2 * HID);
  __syncthreads();
  cols_matvec(a.w.root0_wt + (long long)H * HID, HID, 0, HID, Z, s_z, Z, nvs, nullptr, s_ah, 2 * HID, S, 1, s_h2, 2 * HID);
  cols_matvec(a.w.joint0_wt + (long long)H * HID, HID, 0, HID, Z, s_z, Z, nvs, nullptr, s_ah + HID, 2 * HID, S, 1,
              s_h2 + HID, 2 * HID);
  __syncthreads();
  // second layers: root/intensity (3+K, tanh), joint (6K)
  cols_matvec(a.w.root2_wt, 3 + K, 0, 3 + K, HID, s_h2, 2 * HID, nvs, a.w.root2_b, nullptr, 0, 1, 2, s_dec, 176);
  cols_matvec(a.w.joint2_wt, 6 * K, 0, 6 * K, HID, s_h2 + HID, 2 * HID, nvs, a.w.joint2_b, nullptr, 0, 1, 0, s_dec + 32, 176);
  __syncthreads();
  // ---- rotations + forward kinematics, one thread per (element, sample)
  if ((int)threadIdx.x < nvs) {
    const int vs = threadIdx.x, v = vs / S;
    fk_pose(s_dec + vs * 176, s_dec + vs * 176 + 32, s_off + v * KP_MAX * 3, s_tree, s_tree + KP_MAX, K,
            s_rg + vs * KP_MAX * 9, s_flat + vs * 96);
  }
  __syncthreads();
  // ---- nearest sample to the detected keypoints (:116-123); prior mode: the single sample
  if ((int)threadIdx.x < nb) {
    const int v = threadIdx.x;
    int best = 0;
    if (a.posterior) {
      float bd = INFINITY;
      for (int s = 0; s < S; s++) {
        float d = 0.f;
        for (int j = 0; j < K4; j++) {
          const float t = s_in[v * (H + 96) + H + j] - s_flat[(v * S + s) * 96 + j];
          d += t * t;
        }
        if (d < bd) { bd = d; best = s; }
      }
    }
    s_best[v] = best;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb * (K4 + Z); i += kThreads) {
    const int v = i / (K4 + Z), j = i % (K4 + Z);
    const int vs = v * S + s_best[v];
    const float val = j < K4 ? s_flat[vs * 96 + j] : s_z[vs * Z + (j - K4)];
    s_x[v * 224 + j] = val;
    if (j < K4) a.kp_out[(long long)(b0 + v) * K4 + j] = val;
    else if (a.z_out) a.z_out[(long long)(b0 + v) * Z + (j - K4)] = val;
  }
  if (a.R_out)
    for (int i = threadIdx.x; i < nb * K * 9; i += kThreads) {
      const int v = i / (K * 9), j = i % (K * 9);
      a.R_out[(long long)(b0 + v) * K * 9 + j] = s_rg[(v * S + s_best[v]) * KP_MAX * 9 + j];
    }
  __syncthreads();
  // ---- GRU cell (gate order r, z, n); thread per hidden unit
  const int XI = K4 + Z;
  for (int j = threadIdx.x; j < H; j += kThreads) {
    float gi[3][NB], gh[3][NB];
#pragma unroll
    for (int g = 0; g < 3; g++)
#pragma unroll
      for (int v = 0; v < NB; v++) { gi[g][v] = 0.f; gh[g][v] = 0.f; }
    for (int i = 0; i < XI; i++) {
      const float w0 = __ldg(a.w.gru_ih_wt + (long long)i * 3 * H + j);
      const float w1 = __ldg(a.w.gru_ih_wt + (long long)i * 3 * H + H + j);
      const float w2 = __ldg(a.w.gru_ih_wt + (long long)i * 3 * H + 2 * H + j);
#pragma unroll
      for (int v = 0; v < NB; v++) {
        const float xv = s_x[v * 224 + i];
        gi[0][v] = fmaf(w0, xv, gi[0][v]);
        gi[1][v] = fmaf(w1, xv, gi[1][v]);
        gi[2][v] = fmaf(w2, xv, gi[2][v]);
      }
    }
    for (int i = 0; i < H; i++) {
      const float w0 = __ldg(a.w.gru_hh_wt + (long long)i * 3 * H + j);
      const float w1 = __ldg(a.w.gru_hh_wt + (long long)i * 3 * H + H + j);
      const float w2 = __ldg(a.w.gru_hh_wt + (long long)i * 3 * H + 2 * H + j);
#pragma unroll
      for (int v = 0; v < NB; v++) {
        const float hv = s_in[v * (H + 96) + i];
        gh[0][v] = fmaf(w0, hv, gh[0][v]);
        gh[1][v] = fmaf(w1, hv, gh[1][v]);
        gh[2][v] = fmaf(w2, hv, gh[2][v]);
      }
    }
#pragma unroll
    for (int v = 0; v < NB; v++) {
      if (v < nb) {
        const float r = sigmoid_d(gi[0][v] + a.w.gru_ih_b[j] + gh[0][v] + a.w.gru_hh_b[j]);
        const float zg = sigmoid_d(gi[1][v] + a.w.gru_ih_b[H + j] + gh[1][v] + a.w.gru_hh_b[H + j]);
        const float ng = tanhf(gi[2][v] + a.w.gru_ih_b[2 * H + j] + r * (gh[2][v] + a.w.gru_hh_b[2 * H + j]));
        a.h_out[(long long)(b0 + v) * H + j] = (1.0f - zg) * ng + zg * s_in[v * (H + 96) + j];
      }
    }
  }
}

template <int NB>
size_t step_smem_bytes(int S) {
  const int NVS = ((NB * S + VC - 1) / VC) * VC;
  const int NBP = ((NB + VC - 1) / VC) * VC;
  size_t f = (size_t)NBP * (H + 96) + NBP * HID + NBP * 2 * Z + (size_t)NVS * Z + NBP * 2 * HID + (size_t)NVS * 2 * HID +
             (size_t)NVS * 176 + (size_t)NVS * 96 + (size_t)NVS * KP_MAX * 9 + NBP * 224 + NB * KP_MAX * 3;
  return f * sizeof(float) + (2 * KP_MAX + NB + 8) * sizeof(int);
}

// ------------------------------------------------------------------ small batches: one thread-block cluster per element
// With one CTA per batch element a single SM walks the 6.1 MB of weights alone (296 us per step at B <= 148: 21 GB/s).
// Here a cluster of kCl = 8 CTAs shares one element: the output columns of the big mat-vecs (the four first layers that
// read h, the GRU's hidden-state half - 3.4 MB - and its input half) are split over the cluster ranks, every thread owns
// a quad of columns and a slice of the input (16-byte weight loads, 8 in flight per thread), the 512 hidden values and
// the 512 distribution parameters are broadcast into every CTA's shared memory through DSMEM, and two cluster barriers
// order the three phases.  The small tail (z-part of the decoders, second layers from a shared-memory copy that is
// prefetched with cp.async during the first phases, rotations, forward kinematics, nearest-sample pick) runs redundantly
// in every CTA; each rank then finishes its 64 hidden units of the GRU.
constexpr int kCl = 8;
constexpr int kGruSlice = H / kCl;             // 64 hidden units per rank

__device__ __forceinline__ uint32_t cl_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store v at the same shared-memory offset in every CTA of the cluster
__device__ __forceinline__ void cl_bcast(float* local_ptr, float v) {
  const uint32_t la = (uint32_t)__cvta_generic_to_shared(local_ptr);
#pragma unroll
  for (int r = 0; r < kCl; r++) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(r));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
  }
}
__device__ __forceinline__ void cl_cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

// acc[v][0..3] += sum_{i in [i0, i1)} WT[i][col .. col+3] * x[v][i]; 16-byte weight loads, U rows in flight per thread
// (the mat-vecs are L2-latency bound: 16 rows = 64 KB in flight per CTA for the single-vector phases)
template <int NV>
__device__ __forceinline__ void quad_dot(const float* __restrict__ WT, int ldw, int col, int i0, int i1, const float* x,
                                         int xstride, float (&acc)[NV][4]) {
  constexpr int U = NV == 1 ? 16 : 8;
  const float* w = WT + (long long)i0 * ldw + col;
  int i = i0;
  for (; i + U <= i1; i += U) {
    float4 q[U];
#pragma unroll
    for (int u = 0; u < U; u++) q[u] = __ldg(reinterpret_cast<const float4*>(w + (long long)u * ldw));
    w += (long long)U * ldw;
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int v = 0; v < NV; v++) {
        const float xv = x[v * xstride + i + u];
        acc[v][0] = fmaf(q[u].x, xv, acc[v][0]);
        acc[v][1] = fmaf(q[u].y, xv, acc[v][1]);
        acc[v][2] = fmaf(q[u].z, xv, acc[v][2]);
        acc[v][3] = fmaf(q[u].w, xv, acc[v][3]);
      }
  }
  for (; i < i1; i++) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(w));
    w += ldw;
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const float xv = x[v * xstride + i];
      acc[v][0] = fmaf(q.x, xv, acc[v][0]);
      acc[v][1] = fmaf(q.y, xv, acc[v][1]);
      acc[v][2] = fmaf(q.z, xv, acc[v][2]);
      acc[v][3] = fmaf(q.w, xv, acc[v][3]);
    }
  }
}

struct ClSmem {   // offsets in floats
  int in, hid, ah, gh, dist, part, z, h2, dec, flat, rl, rg, x, off, w2, tree, total;
};
__host__ __device__ inline ClSmem cl_layout(int SV) {
  ClSmem m;
  int o = 0;
  m.in = o;   o += H + 96;
  m.hid = o;  o += 2 * HID;                 // prior0 | post0 outputs
  m.ah = o;   o += 2 * HID;                 // root0 | joint0 applied to h (+ bias)
  m.gh = o;   o += 3 * kGruSlice;           // this rank's W_hh h (3 gates x 64 units)
  m.dist = o; o += 4 * Z;                   // prior mean | std, posterior mean | std
  m.part = o; o += 4 * SV * 256 > 1024 ? 4 * SV * 256 : 1024;
  m.z = o;    o += SV * Z;
  m.h2 = o;   o += SV * 2 * HID;
  m.dec = o;  o += SV * 176;
  m.flat = o; o += SV * 96;
  m.rl = o;   o += SV * KP_MAX * 9;
  m.rg = o;   o += SV * KP_MAX * 9;
  m.x = o;    o += 224;
  m.off = o;  o += KP_MAX * 3;
  m.w2 = o;   o += HID * (3 + KP_MAX) + HID * 6 * KP_MAX;
  m.tree = o; o += 2 * KP_MAX + 8;
  m.total = o;
  return m;
}

template <int SV>
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kThreads, 1)
hsvrnn_step_cluster_kernel(const StepArgs a) {
  extern __shared__ __align__(16) float sm[];
  const ClSmem L = cl_layout(SV);
  float* s_in = sm + L.in;
  float* s_hid = sm + L.hid;
  float* s_ah = sm + L.ah;
  float* s_gh = sm + L.gh;
  float* s_dist = sm + L.dist;
  float* s_part = sm + L.part;
  float* s_z = sm + L.z;
  float* s_h2 = sm + L.h2;
  float* s_dec = sm + L.dec;
  float* s_flat = sm + L.flat;
  float* s_rl = sm + L.rl;
  float* s_rg = sm + L.rg;
  float* s_x = sm + L.x;
  float* s_off = sm + L.off;
  float* s_w2r = sm + L.w2;                              // root2^T [HID][3 + K]
  int* s_tree = reinterpret_cast<int*>(sm + L.tree);
  const int K = a.K, S = a.S, K4 = 4 * a.K, R2 = 3 + a.K, J2 = 6 * a.K;
  float* s_w2j = s_w2r + HID * R2;                       // joint2^T [HID][6 K]
  const int b = blockIdx.x / kCl;
  const int rank = (int)cl_rank();
  const int tid = threadIdx.x;
  const bool need_prior = a.prior_out != nullptr || !a.posterior;
  const int IN = H + K4;

  // ---- stage 0: inputs; the second-layer weights start streaming into shared memory (16-byte aligned segments)
  {
    const int nr = HID * R2, nj = HID * J2;
    const int head = (int)(((16 - ((uintptr_t)a.w.root2_wt & 15)) & 15) >> 2);   // floats until 16-byte alignment (0 in practice)
    for (int i = tid * 4; i + 4 <= nr - head; i += kThreads * 4) cl_cp_async16(s_w2r + head + i, a.w.root2_wt + head + i);
    for (int i = tid; i < head; i += kThreads) s_w2r[i] = a.w.root2_wt[i];
    for (int i = head + ((nr - head) & ~3) + tid; i < nr; i += kThreads) s_w2r[i] = a.w.root2_wt[i];
    // joint2: HID * 6K floats; s_w2j may start off 16-byte alignment when HID * R2 is odd -> plain loads
    if ((((uintptr_t)s_w2j | (uintptr_t)a.w.joint2_wt) & 15) == 0) {
      for (int i = tid * 4; i + 4 <= nj; i += kThreads * 4) cl_cp_async16(s_w2j + i, a.w.joint2_wt + i);
      for (int i = (nj & ~3) + tid; i < nj; i += kThreads) s_w2j[i] = a.w.joint2_wt[i];
    } else {
      for (int i = tid; i < nj; i += kThreads) s_w2j[i] = __ldg(a.w.joint2_wt + i);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int j = tid; j < H + 96; j += kThreads) {
    float val = 0.f;
    if (j < H) val = a.h_in[(long long)b * H + j];
    else if (a.posterior && j - H < K4) val = a.kp[(long long)b * K4 + (j - H)];
    s_in[j] = val;
  }
  for (int i = tid; i < K * 3; i += kThreads) s_off[i] = a.offset[(long long)b * K * 3 + i];
  for (int i = tid; i < K; i += kThreads) {
    s_tree[i] = a.order[i];
    s_tree[KP_MAX + i] = a.parents[i];
  }
  __syncthreads();

  // ---- phase A: everything that reads h (and kp).  256 columns per rank = 64 quads x 4 input slices.
  //   quads 0-3 prior0, 4-7 post0, 8-11 root0[:H], 12-15 joint0[:H] (16 columns each), 16-63 GRU hidden half (3 x 64)
  {
    const int qd = tid & 63, ks = tid >> 6;
    float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    if (qd < 16) {
      const int job = qd >> 2, col = rank * 16 + (qd & 3) * 4;
      const float* WT = job == 0 ? a.w.prior0_wt : (job == 1 ? a.w.post0_wt : (job == 2 ? a.w.root0_wt : a.w.joint0_wt));
      const int I = job == 1 ? IN : H;
      const bool live = job == 0 ? need_prior : (job == 1 ? a.posterior != 0 : true);
      if (live) {
        const int per = (I + 3) / 4;
        quad_dot<1>(WT, HID, col, min(I, ks * per), min(I, (ks + 1) * per), s_in, 0, acc);
      }
    } else {
      const int g = (qd - 16) >> 4, col = g * H + rank * kGruSlice + ((qd - 16) & 15) * 4;
      quad_dot<1>(a.w.gru_hh_wt, 3 * H, col, ks * (H / 4), (ks + 1) * (H / 4), s_in, 0, acc);
    }
#pragma unroll
    for (int e = 0; e < 4; e++) s_part[ks * 256 + qd * 4 + e] = acc[0][e];
  }
  __syncthreads();
  {
    const int c = tid;                                      // local column 0..255
    const float v = s_part[c] + s_part[256 + c] + s_part[512 + c] + s_part[768 + c];
    if (c < 64) {
      const int job = c >> 4, col = rank * 16 + (c & 15);
      if (job == 0) { if (need_prior) cl_bcast(s_hid + col, nm_lrelu(v + a.w.prior0_b[col])); }
      else if (job == 1) { if (a.posterior) cl_bcast(s_hid + HID + col, nm_lrelu(v + a.w.post0_b[col])); }
      else if (job == 2) cl_bcast(s_ah + col, v + a.w.root0_b[col]);
      else cl_bcast(s_ah + HID + col, v + a.w.joint0_b[col]);
    } else {
      s_gh[c - 64] = v;                                    // [gate][unit]
    }
  }
  cl_sync();

  // ---- phase B: second layers of the prior / posterior MLPs: 512 columns, 64 per rank = 16 quads x 16 input slices
  {
    const int qd = tid & 15, ks = tid >> 4;
    const int job = qd >> 3, col = rank * 32 + (qd & 7) * 4;   // job 0: prior2, 1: post2; 32 columns of each per rank
    float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    const bool live = job == 0 ? need_prior : a.posterior != 0;
    if (live) quad_dot<1>(job == 0 ? a.w.prior2_wt : a.w.post2_wt, 2 * Z, col, ks * 8, ks * 8 + 8, s_hid + job * HID, 0, acc);
#pragma unroll
    for (int e = 0; e < 4; e++) s_part[ks * 64 + qd * 4 + e] = acc[0][e];
  }
  __syncthreads();
  if (tid < 64) {
    float v = 0.f;
    for (int ks = 0; ks < 16; ks++) v += s_part[ks * 64 + tid];
    const int job = tid >> 5, col = rank * 32 + (tid & 31);
    const bool live = job == 0 ? need_prior : a.posterior != 0;
    if (live) {
      v += job == 0 ? a.w.prior2_b[col] : a.w.post2_b[col];
      if (col >= Z) v = softplus_d(v) + 1e-4f;
      cl_bcast(s_dist + job * 2 * Z + col, v);
      float* gout = job == 0 ? a.prior_out : a.post_out;
      if (gout) gout[(long long)b * 2 * Z + col] = v;
    }
  }
  cl_sync();

  // ---- from here on every rank works on its own copy.  z = mean + std * eps
  const float* dsel = s_dist + (a.posterior ? 2 * Z : 0);
  for (int i = tid; i < SV * Z; i += kThreads) {
    const int sidx = i / Z, j = i % Z;
    s_z[i] = sidx < S ? dsel[j] + dsel[Z + j] * a.eps[((long long)sidx * a.B + b) * Z + j] : 0.f;
  }
  __syncthreads();
  // decoders, first layer, z part: 256 columns (root | joint) = 64 quads x 4 input slices, all samples at once
  {
    const int qd = tid & 63, ks = tid >> 6;
    float acc[SV][4];
#pragma unroll
    for (int v = 0; v < SV; v++) acc[v][0] = acc[v][1] = acc[v][2] = acc[v][3] = 0.f;
    const bool joint = qd >= 32;
    const int col = (qd & 31) * 4;
    quad_dot<SV>((joint ? a.w.joint0_wt : a.w.root0_wt) + (long long)H * HID, HID, col, ks * (Z / 4), (ks + 1) * (Z / 4), s_z, Z, acc);
#pragma unroll
    for (int v = 0; v < SV; v++)
#pragma unroll
      for (int e = 0; e < 4; e++) s_part[(ks * SV + v) * 256 + qd * 4 + e] = acc[v][e];
  }
  __syncthreads();
  for (int i = tid; i < S * 256; i += kThreads) {
    const int v = i >> 8, c = i & 255;
    const float t = s_part[(0 * SV + v) * 256 + c] + s_part[(1 * SV + v) * 256 + c] + s_part[(2 * SV + v) * 256 + c] +
                    s_part[(3 * SV + v) * 256 + c];
    s_h2[v * 2 * HID + c] = nm_lrelu(t + s_ah[c]);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  // second layers from shared memory: root / intensity (3 + K, tanh) at s_dec[0..], joint 6-D rotations at s_dec[32..]
  for (int i = tid; i < S * (R2 + J2); i += kThreads) {
    const int v = i / (R2 + J2), c = i % (R2 + J2);
    float t;
    if (c < R2) {
      t = a.w.root2_b[c];
      const float* hv = s_h2 + v * 2 * HID;
      for (int k = 0; k < HID; k++) t = fmaf(s_w2r[k * R2 + c], hv[k], t);
      s_dec[v * 176 + c] = tanhf(t);
    } else {
      const int cj = c - R2;
      t = a.w.joint2_b[cj];
      const float* hv = s_h2 + v * 2 * HID + HID;
      for (int k = 0; k < HID; k++) t = fmaf(s_w2j[k * J2 + cj], hv[k], t);
      s_dec[v * 176 + 32 + cj] = t;
    }
  }
  __syncthreads();
  // rotations in parallel over (sample, joint), then the kinematic chain per sample
  for (int i = tid; i < S * K; i += kThreads) rot6d(s_dec + (i / K) * 176 + 32 + (i % K) * 6, s_rl + (i / K) * KP_MAX * 9 + (i % K) * 9);
  __syncthreads();
  if (tid < S) {
    const int v = tid;
    const float* raw = s_dec + v * 176;
    float* Rg = s_rg + v * KP_MAX * 9;
    float* flat = s_flat + v * 96;
    const float* Rl0 = s_rl + v * KP_MAX * 9;
    const int root = s_tree[0];
    for (int e = 0; e < 9; e++) Rg[root * 9 + e] = Rl0[root * 9 + e];
    flat[root * 4] = raw[0]; flat[root * 4 + 1] = raw[1]; flat[root * 4 + 2] = raw[2];
    for (int j = 1; j < K; j++) {
      const int i = s_tree[j], pa = s_tree[KP_MAX + i];
      const float* Rl = Rl0 + i * 9;
      const float* Rp = Rg + pa * 9;
      float* Ri = Rg + i * 9;
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) Ri[r * 3 + c] = Rp[r * 3] * Rl[c] + Rp[r * 3 + 1] * Rl[3 + c] + Rp[r * 3 + 2] * Rl[6 + c];
      const float* o = s_off + i * 3;
#pragma unroll
      for (int r = 0; r < 3; r++) flat[i * 4 + r] = Ri[r * 3] * o[0] + Ri[r * 3 + 1] * o[1] + Ri[r * 3 + 2] * o[2] + flat[pa * 4 + r];
    }
    for (int i = 0; i < K; i++) flat[i * 4 + 3] = (raw[3 + i] + 1.0f) * 0.5f;
  }
  __syncthreads();
  if (tid == 0) {
    int best = 0;
    if (a.posterior) {
      float bd = INFINITY;
      for (int sidx = 0; sidx < S; sidx++) {
        float d = 0.f;
        for (int j = 0; j < K4; j++) {
          const float t = s_in[H + j] - s_flat[sidx * 96 + j];
          d += t * t;
        }
        if (d < bd) { bd = d; best = sidx; }
      }
    }
    s_tree[2 * KP_MAX] = best;
  }
  __syncthreads();
  const int best = s_tree[2 * KP_MAX];
  for (int j = tid; j < 224; j += kThreads) {
    float val = 0.f;
    if (j < K4) val = s_flat[best * 96 + j];
    else if (j < K4 + Z) val = s_z[best * Z + (j - K4)];
    s_x[j] = val;
    if (rank == 0) {
      if (j < K4) a.kp_out[(long long)b * K4 + j] = val;
      else if (j < K4 + Z && a.z_out) a.z_out[(long long)b * Z + (j - K4)] = val;
    }
  }
  if (rank == 0 && a.R_out)
    for (int j = tid; j < K * 9; j += kThreads) a.R_out[(long long)b * K * 9 + j] = s_rg[best * KP_MAX * 9 + j];
  __syncthreads();
  // ---- GRU: this rank's 64 hidden units; input half W_ih x: 192 columns = 48 quads x 4 input slices
  {
    const int XI = K4 + Z;
    const int qd = tid & 63, ks = tid >> 6;
    float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    if (qd < 48) {
      const int g = qd >> 4, col = g * H + rank * kGruSlice + (qd & 15) * 4;
      const int per = (XI + 3) / 4;
      quad_dot<1>(a.w.gru_ih_wt, 3 * H, col, min(XI, ks * per), min(XI, (ks + 1) * per), s_x, 0, acc);
    }
#pragma unroll
    for (int e = 0; e < 4; e++) s_part[ks * 256 + qd * 4 + e] = acc[0][e];
  }
  __syncthreads();
  if (tid < kGruSlice) {
    const int j = rank * kGruSlice + tid;
    float gi[3];
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const int c = g * kGruSlice + tid;
      gi[g] = s_part[c] + s_part[256 + c] + s_part[512 + c] + s_part[768 + c];
    }
    const float r = sigmoid_d(gi[0] + a.w.gru_ih_b[j] + s_gh[tid] + a.w.gru_hh_b[j]);
    const float zg = sigmoid_d(gi[1] + a.w.gru_ih_b[H + j] + s_gh[kGruSlice + tid] + a.w.gru_hh_b[H + j]);
    const float ng = tanhf(gi[2] + a.w.gru_ih_b[2 * H + j] + r * (s_gh[2 * kGruSlice + tid] + a.w.gru_hh_b[2 * H + j]));
    a.h_out[(long long)b * H + j] = (1.0f - zg) * ng + zg * s_in[j];
  }
}

// Pose decoding on its own (extract_kypt_from_latent_and_state, hsvrnn_bvh.py:255-286) for callers that
// drive the sub-modules by hand (vis_generation.py:108-113).
__global__ void __launch_bounds__(kThreads)
decode_pose_kernel(nm_hsvrnn_weights w, const float* __restrict__ dec_in /* (B, H+Z) */,
                   const float* __restrict__ offset, const int* __restrict__ order, const int* __restrict__ parents,
                   int B, int K, float* __restrict__ flat_out, float* __restrict__ R_out) {
  __shared__ float s_in[VC * (H + Z)];
  __shared__ float s_hid[VC * 2 * HID];
  __shared__ float s_dec[176];
  __shared__ float s_rg[KP_MAX * 9];
  __shared__ float s_flat[96];
  __shared__ int s_tree[2 * KP_MAX];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < VC * (H + Z); i += kThreads) s_in[i] = i < H + Z ? dec_in[(long long)b * (H + Z) + i] : 0.f;
  for (int i = threadIdx.x; i < K; i += kThreads) { s_tree[i] = order[i]; s_tree[KP_MAX + i] = parents[i]; }
  __syncthreads();
  cols_matvec(w.root0_wt, HID, 0, HID, H + Z, s_in, H + Z, 1, w.root0_b, nullptr, 0, 1, 1, s_hid, 2 * HID);
  cols_matvec(w.joint0_wt, HID, 0, HID, H + Z, s_in, H + Z, 1, w.joint0_b, nullptr, 0, 1, 1, s_hid + HID, 2 * HID);
  __syncthreads();
  cols_matvec(w.root2_wt, 3 + K, 0, 3 + K, HID, s_hid, 2 * HID, 1, w.root2_b, nullptr, 0, 1, 2, s_dec, 176);
  cols_matvec(w.joint2_wt, 6 * K, 0, 6 * K, HID, s_hid + HID, 2 * HID, 1, w.joint2_b, nullptr, 0, 1, 0, s_dec + 32, 176);
  __syncthreads();
  if (threadIdx.x == 0)
    fk_pose(s_dec, s_dec + 32, offset + (long long)b * K * 3, s_tree, s_tree + KP_MAX, K, s_rg, s_flat);
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * K; i += kThreads) flat_out[(long long)b * 4 * K + i] = s_flat[i];
  if (R_out)
    for (int i = threadIdx.x; i < 9 * K; i += kThreads) R_out[(long long)b * 9 * K + i] = s_rg[i];
}

// Bone offsets (get_offset, hsvrnn_bvh.py:236-253): lower median over T of |kp_k - kp_parent(k)| times the
// unit offset_param direction.  One thread per (b, k); T <= 64.
__global__ void bone_offset_kernel(const float* __restrict__ kp /* (B,T,K,4) */, const int* __restrict__ parents,
                                   const float* __restrict__ offset_param /* (K,3) */, int B, int T, int K,
                                   float* __restrict__ out /* (B,K,3) */) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * K) return;
  const int b = idx / K, k = idx % K, pa = parents[k];
  float d[64];
  for (int t = 0; t < T; t++) {
    const float* p = kp + (((long long)b * T + t) * K + k) * 4;
    const float* q = kp + (((long long)b * T + t) * K + pa) * 4;
    const float dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    d[t] = sqrtf(dx * dx + dy * dy + dz * dz);
  }
  // insertion sort, lower median = element (T-1)/2 (torch.median semantics)
  for (int i = 1; i < T; i++) {
    const float v = d[i];
    int j = i - 1;
    while (j >= 0 && d[j] > v) { d[j + 1] = d[j]; j--; }
    d[j + 1] = v;
  }
  const float med = d[(T - 1) / 2];
  const float* o = offset_param + k * 3;
  const float nrm = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]) + 1e-10f;
  for (int c = 0; c < 3; c++) out[(long long)idx * 3 + c] = (o[c] / nrm) * med;
}

}  // namespace

extern "C" int nm_hsvrnn_step(const nm_hsvrnn_weights* w, const float* h_in, const float* kp, const float* eps,
                              const float* offset, const int* order, const int* parents, int B, int K, int S,
                              int posterior, float* h_out, float* kp_out, float* z_out, float* R_out,
                              float* post_out, float* prior_out, void* stream) {
  NM_CHECK_ARG(w && h_in && eps && offset && order && parents && h_out && kp_out, "nm_hsvrnn_step: null pointer");
  NM_CHECK_ARG(!posterior || kp, "nm_hsvrnn_step: posterior step needs detected keypoints");
  NM_CHECK_ARG(K > 0 && K <= KP_MAX && S >= 1 && S <= 16, "nm_hsvrnn_step: K=%d S=%d unsupported", K, S);
  NM_CHECK_ARG(posterior || S == 1, "nm_hsvrnn_step: prior step draws one sample");
  if (B == 0) return NM_OK;
  StepArgs a;
  a.w = *w; a.h_in = h_in; a.kp = kp; a.eps = eps; a.offset = offset; a.order = order; a.parents = parents;
  a.h_out = h_out; a.kp_out = kp_out; a.z_out = z_out; a.R_out = R_out; a.post_out = post_out; a.prior_out = prior_out;
  a.B = B; a.K = K; a.S = S; a.posterior = posterior;
  cudaStream_t st = (cudaStream_t)stream;
  // small and medium batches: a cluster of 8 CTAs per element (weights split over the ranks, DSMEM broadcasts)
  static const bool use_cluster = []() { const char* e = getenv("NM_HSVRNN_CLUSTER"); return !(e && atoi(e) == 0); }();
  // measured (tools/time_rollout.py, us per step): B = 1: 30 vs 216, B = 16: 59 vs 228, B = 64: 144 vs 235, B = 256: 510 vs 240
  if (use_cluster && S <= 10 && B <= 96) {
    const int SV = S == 1 ? 1 : 10;
    const size_t smem = (size_t)cl_layout(SV).total * sizeof(float);
    if (SV == 1) {
      NM_CHECK_CUDA(cudaFuncSetAttribute(hsvrnn_step_cluster_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      hsvrnn_step_cluster_kernel<1><<<B * kCl, kThreads, smem, st>>>(a);
    } else {
      NM_CHECK_CUDA(cudaFuncSetAttribute(hsvrnn_step_cluster_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      hsvrnn_step_cluster_kernel<10><<<B * kCl, kThreads, smem, st>>>(a);
    }
    NM_CHECK_LAUNCH("hsvrnn_step(cluster)");
    return NM_OK;
  }
  if (B >= 4 * nm_num_sms() && step_smem_bytes<4>(S) <= 227 * 1024) {
    const size_t smem = step_smem_bytes<4>(S);
    NM_CHECK_CUDA(cudaFuncSetAttribute(hsvrnn_step_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hsvrnn_step_kernel<4><<<nm_cdiv(B, 4), kThreads, smem, st>>>(a);
  } else {
    const size_t smem = step_smem_bytes<1>(S);
    NM_CHECK_CUDA(cudaFuncSetAttribute(hsvrnn_step_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hsvrnn_step_kernel<1><<<B, kThreads, smem, st>>>(a);
  }
  NM_CHECK_LAUNCH("hsvrnn_step");
  return NM_OK;
}

extern "C" int nm_hsvrnn_decode_pose(const nm_hsvrnn_weights* w, const float* dec_in, const float* offset,
                                     const int* order, const int* parents, int B, int K, float* flat_out,
                                     float* R_out, void* stream) {
  NM_CHECK_ARG(w && dec_in && offset && order && parents && flat_out, "nm_hsvrnn_decode_pose: null pointer");
  NM_CHECK_ARG(K > 0 && K <= KP_MAX, "nm_hsvrnn_decode_pose: K=%d unsupported", K);
  if (B == 0) return NM_OK;
  decode_pose_kernel<<<B, kThreads, 0, (cudaStream_t)stream>>>(*w, dec_in, offset, order, parents, B, K, flat_out, R_out);
  NM_CHECK_LAUNCH("hsvrnn_decode_pose");
  return NM_OK;
}

extern "C" int nm_hsvrnn_bone_offsets(const float* keypoints, const int* parents, const float* offset_param, int B,
                                      int T, int K, float* out, void* stream) {
  NM_CHECK_ARG(keypoints && parents && offset_param && out, "nm_hsvrnn_bone_offsets: null pointer");
  NM_CHECK_ARG(T >= 1 && T <= 64, "nm_hsvrnn_bone_offsets: T=%d unsupported (1..64)", T);
  if (B == 0) return NM_OK;
  bone_offset_kernel<<<nm_cdiv(B * K, 128), 128, 0, (cudaStream_t)stream>>>(keypoints, parents, offset_param, B, T, K, out);
  NM_CHECK_LAUNCH("hsvrnn_bone_offsets");
  return NM_OK;
}
