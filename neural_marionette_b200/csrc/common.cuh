// Shared helpers for the nm_b200 kernels (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

// Activations at rest: fp16, channels-last (N, D, H, W, C).  fp16 (10-bit mantissa)
// rather than bf16 keeps the heat-map error an order of magnitude inside the
// 1e-2 budget at identical tensor-core throughput (kind::f16 covers both).
typedef __half act_t;

#define NM_OK 0
#define NM_ERR_ARG 1
#define NM_ERR_CUDA 2
#define NM_ERR_DRIVER 3

void nm_set_error(const char* fmt, ...);

#define NM_CHECK_ARG(cond, ...)                   \
  do {                                            \
    if (!(cond)) {                                \
      nm_set_error(__VA_ARGS__);                  \
      return NM_ERR_ARG;                          \
    }                                             \
  } while (0)

#define NM_CHECK_LAUNCH(name)                                                   \
  do {                                                                          \
    cudaError_t e_ = cudaGetLastError();                                        \
    if (e_ != cudaSuccess) {                                                    \
      nm_set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));      \
      return NM_ERR_CUDA;                                                       \
    }                                                                           \
  } while (0)

#define NM_CHECK_CUDA(expr)                                                     \
  do {                                                                          \
    cudaError_t e_ = (expr);                                                    \
    if (e_ != cudaSuccess) {                                                    \
      nm_set_error("%s: %s", #expr, cudaGetErrorString(e_));                    \
      return NM_ERR_CUDA;                                                       \
    }                                                                           \
  } while (0)

static inline int nm_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
int nm_num_sms();

__device__ __forceinline__ float nm_lrelu(float x) { return x > 0.f ? x : 0.01f * x; }

__device__ __forceinline__ float nm_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 8 halfs <-> 8 floats
struct __align__(16) half8 { __half2 h[4]; };
__device__ __forceinline__ void nm_unpack8(const half8& v, float* f) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float2 t = __half22float2(v.h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ half8 nm_pack8(const float* f) {
  half8 v;
#pragma unroll
  for (int i = 0; i < 4; i++) v.h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// Runs `body` on the first call per (call site, device) - kernel attributes such as the dynamic shared-memory limit are
// per device.  Two threads racing on the first call both run the (idempotent) body before either sets the bit.
#define NM_PER_DEVICE_ONCE(...)                                                       \
  do {                                                                                \
    static std::atomic<unsigned long long> nm_once_mask{0};                           \
    int nm_once_dev = 0;                                                              \
    cudaGetDevice(&nm_once_dev);                                                      \
    const unsigned long long nm_once_bit = 1ull << (nm_once_dev & 63);                \
    if (!(nm_once_mask.load(std::memory_order_acquire) & nm_once_bit)) {              \
      __VA_ARGS__;                                                                    \
      nm_once_mask.fetch_or(nm_once_bit, std::memory_order_release);                  \
    }                                                                                 \
  } while (0)
