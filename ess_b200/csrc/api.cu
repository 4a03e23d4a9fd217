// Library info, error reporting and device capability check for libess_b200.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
}

void essb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int essb_version(void) { return ESSB_VERSION; }
extern "C" const char* essb_build_arch(void) { return "sm_100a"; }
extern "C" const char* essb_last_error(void) { return g_err; }

extern "C" int essb_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    essb_set_error("essb_device_check: no CUDA device: %s", cudaGetErrorString(e));
    return ESSB_ERR_ARCH;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    essb_set_error("essb_device_check: device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
    return ESSB_ERR_ARCH;
  }
  return ESSB_OK;
}
