// emb_opt_agc_rms_momentum: the reference's optimiser chain on the flat
// parameter buffer in two HBM passes (dreamerv3/agent.py:342-379;
// embodied/jax/opt.py:109-164):
//
//   clip_by_agc(clip, pmin)   per tensor: g *= 1 / max(1, |g| / (clip * max(pmin, |w|)))
//   scale_by_rms(b2, eps)     nu = b2 nu + (1-b2) g^2 ;  g /= sqrt(nu * c2) + eps     c2 = 1/(1-b2^t)
//   scale_by_momentum(b1)     mu = b1 mu + (1-b1) g   ;  g  = mu * c1                 c1 = 1/(1-b1^t)
//   w -= lr * g
//
// Pass 1 reduces |g|^2 and |w|^2 per tensor (one CTA per 4096-element chunk, the
// chunk -> tensor table is built once by the host); pass 2 streams g, nu, mu, w
// once and writes nu, mu, w.  28 bytes per parameter + 8 for the norms: HBM-bound.
// Step-dependent scalars (lr, c1, c2) are read from DEVICE memory so the launch
// can sit inside a CUDA graph.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
opt_norms_kernel(const float* __restrict__ g, const float* __restrict__ w,
                 const emb_opt_chunk* __restrict__ chunks, float* __restrict__ partials) {
  const emb_opt_chunk c = chunks[blockIdx.x];
  const float4* g4 = reinterpret_cast<const float4*>(g + c.begin);
  const float4* w4 = reinterpret_cast<const float4*>(w + c.begin);
  float sg = 0.f, sw = 0.f;
  const int n4 = c.count >> 2;
  for (int i = threadIdx.x; i < n4; i += kThreads) {
    const float4 a = g4[i], b = w4[i];
    sg += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    sw += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
  }
  for (int i = (n4 << 2) + threadIdx.x; i < c.count; i += kThreads) {
    const float a = g[c.begin + i], b = w[c.begin + i];
    sg += a * a; sw += b * b;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
  }
  __shared__ float red[2][kThreads / 32];
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sg; red[1][threadIdx.x >> 5] = sw; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) { a += red[0][i]; b += red[1][i]; }
    // no atomics: the per-tensor sums are formed in a FIXED order by opt_norms_finish_kernel, so
    // every data-parallel rank computes bit-identical clip factors from identical gradients
    partials[2 * (size_t)blockIdx.x] = a;
    partials[2 * (size_t)blockIdx.x + 1] = b;
  }
}

// One warp per tensor: sum its chunks' partial (|g|^2, |w|^2) in a fixed order.
__global__ void opt_norms_finish_kernel(const float* __restrict__ partials, const int32_t* __restrict__ first,
                                        int32_t chunk0, int32_t tensor0, float* __restrict__ norms) {
  const int t = tensor0 + blockIdx.x;
  const int lo = first[t] - chunk0, hi = first[t + 1] - chunk0;
  float a = 0.f, b = 0.f;
  for (int i = lo + (int)threadIdx.x; i < hi; i += 32) {
    a += partials[2 * (size_t)i];
    b += partials[2 * (size_t)i + 1];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (threadIdx.x == 0) {
    norms[2 * t] = a;
    norms[2 * t + 1] = b;
  }
}

__global__ void __launch_bounds__(kThreads)
opt_update_kernel(const float* __restrict__ g, float* __restrict__ w, float* __restrict__ nu,
                  float* __restrict__ mu, __nv_bfloat16* __restrict__ low,
                  const emb_opt_chunk* __restrict__ chunks, const float* __restrict__ norms,
                  const float* __restrict__ hyper) {
  const emb_opt_chunk c = chunks[blockIdx.x];
  const float lr = hyper[0], c1 = hyper[1], c2 = hyper[2], b1 = hyper[3], b2 = hyper[4];
  const float eps = hyper[5], clip = hyper[6], pmin = hyper[7];
  const float gn = sqrtf(norms[2 * c.tensor]), pn = sqrtf(norms[2 * c.tensor + 1]);
  const float upper = clip * fmaxf(pmin, pn);
  const float scale = clip > 0.f ? 1.0f / fmaxf(1.0f, gn / upper) : 1.0f;      // opt.py:115-121
  auto one = [&](float gi, float& wi, float& nui, float& mui) {
    const float u0 = gi * scale;
    nui = b2 * nui + (1.0f - b2) * (u0 * u0);                                  // opt.py:137-138
    const float u1 = u0 / (sqrtf(nui * c2) + eps);                             // opt.py:139-141
    mui = (1.0f - b1) * u1 + b1 * mui;                                         // optax.update_moment
    wi -= lr * (mui * c1);
  };
  const int n4 = c.count >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g + c.begin);
  float4* w4 = reinterpret_cast<float4*>(w + c.begin);
  float4* nu4 = reinterpret_cast<float4*>(nu + c.begin);
  float4* mu4 = reinterpret_cast<float4*>(mu + c.begin);
  for (int i = threadIdx.x; i < n4; i += kThreads) {
    const float4 gi = g4[i];
    float4 wi = w4[i], ni = nu4[i], mi = mu4[i];
    one(gi.x, wi.x, ni.x, mi.x); one(gi.y, wi.y, ni.y, mi.y);
    one(gi.z, wi.z, ni.z, mi.z); one(gi.w, wi.w, ni.w, mi.w);
    w4[i] = wi; nu4[i] = ni; mu4[i] = mi;
    if (low) {                             // the compute-dtype copy the next forward pass reads
      const __nv_bfloat162 lo = __floats2bfloat162_rn(wi.x, wi.y), hi = __floats2bfloat162_rn(wi.z, wi.w);
      uint2 packed;
      packed.x = *reinterpret_cast<const uint32_t*>(&lo);
      packed.y = *reinterpret_cast<const uint32_t*>(&hi);
      reinterpret_cast<uint2*>(low + c.begin)[i] = packed;
    }
  }
  for (int i = (n4 << 2) + threadIdx.x; i < c.count; i += kThreads) {
    const int64_t at = c.begin + i;
    float wi = w[at], ni = nu[at], mi = mu[at];
    one(g[at], wi, ni, mi);
    w[at] = wi; nu[at] = ni; mu[at] = mi;
    if (low) low[at] = __float2bfloat16_rn(wi);
  }
}

}  // namespace

extern "C" int emb_opt_agc_rms_momentum_cast(const float* grad, float* param, float* nu, float* mu,
                                             void* param_bf16, const emb_opt_chunk* chunks,
                                             int32_t nchunks, float* norms, int32_t ntensors,
                                             const float* hyper, float* partials,
                                             const int32_t* tensor_first, void* stream) {
  const char* who = "emb_opt_agc_rms_momentum";
  if (nchunks < 0 || ntensors < 0) return emb::fail(-1, "%s: negative sizes", who);
  if (nchunks == 0) return 0;
  if (!grad || !param || !nu || !mu || !chunks || !norms || !hyper || !partials || !tensor_first)
    return emb::fail(-1, "%s: NULL argument", who);
  if (((uintptr_t)grad | (uintptr_t)param | (uintptr_t)nu | (uintptr_t)mu) & 15)
    return emb::fail(-1, "%s: buffers must be 16-byte aligned", who);
  if ((uintptr_t)param_bf16 & 7) return emb::fail(-1, "%s: param_bf16 must be 8-byte aligned", who);
  cudaStream_t s = (cudaStream_t)stream;
  opt_norms_kernel<<<nchunks, kThreads, 0, s>>>(grad, param, chunks, partials);
  emb::count_launch();
  opt_norms_finish_kernel<<<ntensors, 32, 0, s>>>(partials, tensor_first, 0, 0, norms);
  emb::count_launch();
  opt_update_kernel<<<nchunks, kThreads, 0, s>>>(grad, param, nu, mu,
                                                 reinterpret_cast<__nv_bfloat16*>(param_bf16), chunks,
                                                 norms, hyper);
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}

extern "C" int emb_opt_agc_rms_momentum(const float* grad, float* param, float* nu, float* mu,
                                        const emb_opt_chunk* chunks, int32_t nchunks,
                                        float* norms, int32_t ntensors, const float* hyper,
                                        float* partials, const int32_t* tensor_first, void* stream) {
  return emb_opt_agc_rms_momentum_cast(grad, param, nu, mu, nullptr, chunks, nchunks, norms,
                                       ntensors, hyper, partials, tensor_first, stream);
}

// ---- emb_allreduce_bucket_update -------------------------------------------------
// One gradient BUCKET (a contiguous run of whole tensors of the flat buffers) end to end on one
// stream: NCCL all-reduce (average over the data-parallel ranks: embodied/jax/opt.py:52-54
// `pmean`) of the bucket's gradients in place, then the optimiser chain on exactly those
// tensors.  The host launches it from the backward pass as soon as the bucket's last gradient
// has been accumulated, on a side stream, so both the exchange and the update of the early
// buckets (heads, decoder) overlap the rest of the backward pass.
namespace {

typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
nccl_allreduce_fn g_allreduce = nullptr;
nccl_errstr_fn g_nccl_errstr = nullptr;

bool load_nccl() {
  if (g_allreduce) return true;
  // the copy PyTorch already loaded (same communicator objects), else whatever the loader finds
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
  g_allreduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
  g_nccl_errstr = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
  return g_allreduce != nullptr;
}

}  // namespace

extern "C" int emb_allreduce_bucket_update(void* nccl_comm, float* grad, float* param, float* nu,
                                           float* mu, void* param_bf16, int64_t elem_begin,
                                           int64_t elem_count, const emb_opt_chunk* chunks,
                                           int32_t nchunks, float* norms, int32_t tensor_begin,
                                           int32_t tensor_count, const float* hyper, float* partials,
                                           const int32_t* tensor_first, int32_t chunk_begin,
                                           void* stream) {
  const char* who = "emb_allreduce_bucket_update";
  if (nchunks <= 0 || elem_count <= 0 || tensor_count <= 0) return 0;
  if (!grad || !param || !nu || !mu || !chunks || !norms || !hyper || !partials || !tensor_first ||
      tensor_begin < 0 || elem_begin < 0 || chunk_begin < 0)
    return emb::fail(-1, "%s: NULL / negative argument", who);
  cudaStream_t s = (cudaStream_t)stream;
  if (nccl_comm) {
    if (!load_nccl()) return emb::fail(-4, "%s: libnccl.so.2 (ncclAllReduce) not found", who);
    // ncclFloat32 = 7, ncclAvg = 4 (nccl.h)
    const int rc = g_allreduce(grad + elem_begin, grad + elem_begin, (size_t)elem_count, 7, 4, nccl_comm, s);
    if (rc != 0)
      return emb::fail(-5, "%s: ncclAllReduce failed: %s", who, g_nccl_errstr ? g_nccl_errstr(rc) : "?");
  }
  // partials of this bucket live at [chunk_begin, chunk_begin + nchunks) of the global scratch
  float* part = partials + 2 * (size_t)chunk_begin;
  opt_norms_kernel<<<nchunks, kThreads, 0, s>>>(grad, param, chunks, part);
  emb::count_launch();
  opt_norms_finish_kernel<<<tensor_count, 32, 0, s>>>(part, tensor_first, chunk_begin, tensor_begin, norms);
  emb::count_launch();
  opt_update_kernel<<<nchunks, kThreads, 0, s>>>(grad, param, nu, mu,
                                                 reinterpret_cast<__nv_bfloat16*>(param_bf16), chunks,
                                                 norms, hyper);
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}
