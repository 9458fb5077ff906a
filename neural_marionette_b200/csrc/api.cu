// Error reporting and device queries for the nm_b200 C ABI.
#include "common.cuh"
#include "../../include/nm_b200.h"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void nm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int nm_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!sms[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

extern "C" const char* nm_last_error(void) { return g_err; }
extern "C" int nm_version(void) { return 100; }

extern "C" int nm_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { nm_set_error("no CUDA device"); return NM_ERR_CUDA; }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) { nm_set_error("nm_b200 kernels are built for sm_100a only (device is sm_%d*)", major); return NM_ERR_ARG; }
  return NM_OK;
}
