// emb_probe_read: measures how fast a persistent grid can STREAM a buffer out of
// HBM with the scan kernels' weight path (one producer thread per CTA issuing
// cp.async.bulk into a shared-memory ring, consumers only releasing the stages)
// or with plain 16-byte loads.  A diagnostic for the roofline of rssm_*_tma.cu:
// the scan is read-only traffic, the measured "copy" peak is half reads, half writes.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"
#include "rssm_tma.cuh"

namespace {

using namespace rssm_tma;

__global__ void __launch_bounds__(kAllThreads, 1)
probe_tma_kernel(const unsigned char* src, size_t bytes_per_cta, int nstages, int stage_bytes,
                 unsigned* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  unsigned char* data = smem + 256;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kCWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned char* blk = src + (size_t)blockIdx.x * bytes_per_cta;
  int stage = 0;
  uint32_t phase = 0;
  if (tid >= kCThreads) {
    if (tid == kCThreads) {
      for (size_t off = 0; off < bytes_per_cta; off += stage_bytes) {
        mbar_wait(&empty[stage], phase ^ 1u);
        mbar_expect_tx(&full[stage], stage_bytes);
        bulk_g2s(data + (size_t)stage * stage_bytes, blk + off, stage_bytes, &full[stage]);
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }
  unsigned acc = 0;
  for (size_t off = 0; off < bytes_per_cta; off += stage_bytes) {
    mbar_wait(&full[stage], phase);
    acc += reinterpret_cast<const unsigned*>(data + (size_t)stage * stage_bytes)[tid];
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[stage]);
    if (++stage == nstages) { stage = 0; phase ^= 1u; }
  }
  if (acc == 0x12345678u) *sink = acc;
}

__global__ void __launch_bounds__(256)
probe_ldg_kernel(const uint4* src, size_t vecs_per_cta, unsigned* sink) {
  const uint4* p = src + (size_t)blockIdx.x * vecs_per_cta;
  unsigned acc = 0;
  for (size_t i = threadIdx.x; i < vecs_per_cta; i += 256 * 8) {
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] = i + j * 256 < vecs_per_cta ? __ldcs(p + i + j * 256) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

}  // namespace

extern "C" int emb_probe_read(const void* src, uint64_t bytes, int32_t ncta, int32_t mode,
                              int32_t nstages, int32_t stage_bytes, void* sink, void* stream) {
  const char* who = "emb_probe_read";
  if (ncta < 1 || bytes == 0) return emb::fail(-1, "%s: ncta=%d bytes=%llu", who, ncta, (unsigned long long)bytes);
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 0) {
    if (nstages < 1 || nstages > 16 || stage_bytes % 16 || stage_bytes < 16)
      return emb::fail(-1, "%s: nstages=%d stage_bytes=%d", who, nstages, stage_bytes);
    const size_t per = bytes / ncta / stage_bytes * stage_bytes;
    const size_t smem = 256 + (size_t)nstages * stage_bytes;
    if (smem > 227 * 1024) return emb::fail(-1, "%s: ring of %zu bytes", who, smem);
    if (cudaFuncSetAttribute(probe_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return emb::fail_cuda(who);
    probe_tma_kernel<<<ncta, kAllThreads, smem, s>>>((const unsigned char*)src, per, nstages, stage_bytes,
                                                     (unsigned*)sink);
  } else {
    const size_t vecs = bytes / 16 / ncta;
    probe_ldg_kernel<<<ncta, 256, 0, s>>>((const uint4*)src, vecs, (unsigned*)sink);
  }
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}
