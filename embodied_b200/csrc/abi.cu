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
void emb_launch_count_add(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int emb_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return emb::fail_cuda("emb_device_sm_count");
  return sms;
}

/* CUDA-event stopwatch that also works inside a stream capture: when `stream`
 * is capturing, the record becomes an event-record NODE of the graph
 * (cudaEventRecordExternal), re-recorded at every replay. */
int emb_event_create(void** ev) {
  cudaEvent_t e;
  if (!ev || cudaEventCreate(&e) != cudaSuccess) return emb::fail_cuda("emb_event_create");
  *ev = (void*)e;
  return 0;
}

int emb_event_record(void* ev, void* stream) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing((cudaStream_t)stream, &st) != cudaSuccess)
    return emb::fail_cuda("emb_event_record");
  const unsigned flags = st == cudaStreamCaptureStatusActive ? cudaEventRecordExternal
                                                             : cudaEventRecordDefault;
  if (cudaEventRecordWithFlags((cudaEvent_t)ev, (cudaStream_t)stream, flags) != cudaSuccess)
    return emb::fail_cuda("emb_event_record");
  return 0;
}

int emb_event_elapsed_ms(void* a, void* b, float* ms) {
  if (cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess)
    return emb::fail_cuda("emb_event_elapsed_ms");
  return 0;
}

int emb_event_destroy(void* ev) {
  if (cudaEventDestroy((cudaEvent_t)ev) != cudaSuccess) return emb::fail_cuda("emb_event_destroy");
  return 0;
}

}  // extern "C"
