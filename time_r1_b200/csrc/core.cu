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

// Programmatic-dependent-launch switch (see t1_common.cuh): the host sets it around the decode step it captures.
static int g_pdl = 0;
extern "C" void t1_set_pdl(int on) { g_pdl = on ? 1 : 0; }
int t1_pdl_enabled(void) { return g_pdl; }

// ---- tensor-map construction through the driver entry point (no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int t1_make_tmap_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows,
                      int box_cols) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    t1_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return T1_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    t1_set_error("cuTensorMapEncodeTiled failed: CUresult %d (ptr %p rows %lld cols %lld ld %lld)", (int)r, ptr, rows,
                 cols, ld);
    return T1_ERR_CUDA;
  }
  return T1_OK;
}
