// Error plumbing and device queries of the C ABI (include/cdscore.h).
#include <stdarg.h>
#include "common.cuh"
#include "../../include/cdscore.h"

static thread_local char g_err[512] = "";

void cds_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int cds_abi_version(void) { return 2; }
extern "C" const char* cds_last_error(void) { return g_err; }

extern "C" int cds_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    cds_set_error("cds_device_info: %s", cudaGetErrorString(e));
    return CDS_ERR_CUDA;
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return CDS_OK;
}
