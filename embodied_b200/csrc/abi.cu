// Error channel, version and launch counter of the C ABI.
#include <stdio.h>

#include <atomic>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

namespace emb {

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int fail_cuda(const char* who) {
  cudaError_t e = cudaGetLastError();
  snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", who, (int)e,
           cudaGetErrorString(e));
  return -100 - (int)e;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace emb

extern "C" {

const char* emb_last_error(void) { return g_err; }
int emb_abi_version(void) { return EMB_ABI_VERSION; }
uint64_t emb_launch_count(void) { return g_launches.load(); }

int emb_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return emb::fail_cuda("emb_device_sm_count");
  return sms;
}

}  // extern "C"
