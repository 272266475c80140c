// libtimer1.so core: error reporting and build/runtime introspection.
#include "t1_common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[1024] = "";

void t1_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* t1_last_error(void) { return g_err; }

extern "C" int t1_version(void) { return 100; }

// Returns 0 and fills {sm_count, cc_major, cc_minor, l2_bytes} for the current device.
extern "C" int t1_device_info(int* out4) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    t1_set_error("t1_device_info: %s", cudaGetErrorString(e));
    return T1_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&out4[0], cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&out4[1], cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&out4[2], cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(&out4[3], cudaDevAttrL2CacheSize, dev);
  return T1_OK;
}

// Number of kernel launches issued through this library since the last reset (bench.py "gpu_launches").
static unsigned long long g_launches = 0;
extern "C" void t1_count_launch(void) { ++g_launches; }
extern "C" unsigned long long t1_launch_count(int reset) {
  unsigned long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}
