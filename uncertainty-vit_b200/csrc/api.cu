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

static int g_sm_limit = 0;   // 0 = all SMs

int b200vit_num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached[dev] = n;
  }
  // every persistent kernel of the library sizes its grid from this number: a limit leaves the remaining SMs to a concurrent kernel
  // (the NCCL all-reduce overlapped with the backward pass)
  return (g_sm_limit > 0 && g_sm_limit < cached[dev]) ? g_sm_limit : cached[dev];
}

extern "C" const char* b200vit_last_error(void) { return g_err; }
extern "C" int b200vit_abi_version(void) { return B200VIT_ABI_VERSION; }
extern "C" int b200vit_device_sm_count(void) { return b200vit_num_sms(); }
extern "C" int b200vit_set_sm_limit(int32_t n) {
  const int prev = g_sm_limit;
  g_sm_limit = n > 0 ? (n & ~1) : 0;       // CTA pairs: keep it even
  return prev;
}
