// Sampling noise of the learner (noise is an explicit input of every sampled op).
// emb_gumbel_fill writes Gumbel(0, 1) noise for the categorical draws of one update
// (outs.OneHot / Categorical sample = arg-max of log-probabilities + Gumbel noise) in ONE
// write-only pass: Philox4x32-10 counters, four values per call, g = -log(E), E = -log(U)
// clamped to [1e-7, 46] (so g stays inside [-3.83, 16.2]).  HBM-bound: 4 bytes written per
// value, nothing read (the torch formulation exponential_ -> clamp_ -> log_ -> neg_ makes
// four passes, 28 bytes per value).
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {

__device__ __forceinline__ float gumbel_of(float u) {       // u in (0, 1]
  const float e = fminf(fmaxf(-logf(u), 1e-7f), 46.0f);
  return -logf(e);
}

__global__ void __launch_bounds__(256)
gumbel_kernel(float* __restrict__ out, int64_t n, unsigned long long seed) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)tid, 0ull, &st);    // one subsequence per thread
  for (int64_t i = tid * 4; i < n; i += nthreads * 4) {
    const float4 u = curand_uniform4(&st);
    const float4 g = make_float4(gumbel_of(u.x), gumbel_of(u.y), gumbel_of(u.z), gumbel_of(u.w));
    if (i + 3 < n) {
      __stcs(reinterpret_cast<float4*>(out + i), g);
    } else {
      const float v[4] = {g.x, g.y, g.z, g.w};
      for (int j = 0; j < 4 && i + j < n; ++j) out[i + j] = v[j];
    }
  }
}

}  // namespace

extern "C" int emb_gumbel_fill(float* out, int64_t n, uint64_t seed, void* stream) {
  const char* who = "emb_gumbel_fill";
  if (n < 0) return emb::fail(-1, "%s: n=%lld", who, (long long)n);
  if (n == 0) return 0;
  if (!out) return emb::fail(-1, "%s: null pointer", who);
  if (((uintptr_t)out & 15) != 0) return emb::fail(-1, "%s: out must be 16-byte aligned", who);
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n / 4 + 255) / 256;
  const int blocks = (int)(want < 1 ? 1 : (want > (int64_t)sms * 8 ? (int64_t)sms * 8 : want));
  gumbel_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, n, (unsigned long long)seed);
  if (cudaGetLastError() != cudaSuccess) return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}
