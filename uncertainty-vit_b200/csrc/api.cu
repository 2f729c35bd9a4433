// Error plumbing and device queries of libb200vit.
#include <cstdarg>
#include <cstdio>

#include "../../include/b200vit.h"
#include "common.cuh"

static thread_local char g_err[512] = "";

void b200vit_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int b200vit_num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached[dev] = n;
  }
  return cached[dev];
}

extern "C" const char* b200vit_last_error(void) { return g_err; }
extern "C" int b200vit_abi_version(void) { return B200VIT_ABI_VERSION; }
extern "C" int b200vit_device_sm_count(void) { return b200vit_num_sms(); }
