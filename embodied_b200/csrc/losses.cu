// emb_twohot_loss_fwd / _bwd, emb_twohot_pred, emb_loss_reduce: the head losses of
// dreamerv3/agent.py:172-182,237-240,382-479 (embodied/jax/outs.py:273-330 TwoHot) as
// one pass each instead of ~20 element-wise launches per call.
//
//   loss[r] = -(sum_k twohot(t1[r])_k logp_k) - w2 * (sum_k twohot(t2[r])_k logp_k)
//   twohot(t): the two bins around t, weighted by the distance to the OTHER one
//              (outs.py:314-327; targets outside the bin range land on the end bin)
//   pred[r] = sum_k softmax(logits[r])_k bins_k, summed SYMMETRICALLY around the middle bin
//             (outs.py:285-309) so that symmetric bins + uniform logits give exactly 0
//   grad    = g[r] * ((1 + w2) softmax - twohot(t1) - w2 twohot(t2))
//
// One warp per row, the nb <= 1024 bins strided over the lanes, every reduction a
// shuffle butterfly; fp32 throughout (outs.py:276).  HBM traffic = the logits once
// (forward) / once + the gradient (backward).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int wsum_i(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct TwoHot { int below, above; float wb, wa; };

// outs.py:314-324 for one target; `le` = #(bins <= t), `gt` = #(bins > t) (warp-reduced counts)
__device__ __forceinline__ TwoHot twohot_of(const float* bins, int nb, float t, int le, int gt) {
  TwoHot h;
  h.below = min(max(le - 1, 0), nb - 1);
  h.above = min(max(nb - gt, 0), nb - 1);
  const bool equal = h.below == h.above;
  const float db = equal ? 1.0f : fabsf(bins[h.below] - t);
  const float da = equal ? 1.0f : fabsf(bins[h.above] - t);
  const float total = db + da;
  h.wb = da / total;
  h.wa = db / total;
  return h;
}

__device__ __forceinline__ TwoHot twohot_warp(const float* bins, int nb, float t, int lane) {
  int le = 0, gt = 0;
  for (int k = lane; k < nb; k += 32) {
    const float b = bins[k];
    le += b <= t;
    gt += b > t;
  }
  return twohot_of(bins, nb, t, wsum_i(le), wsum_i(gt));
}

__global__ void twohot_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ t1,
                                  const float* __restrict__ t2, float w2, const float* __restrict__ bins,
                                  float* __restrict__ loss, float* __restrict__ lse_out, int64_t rows, int nb) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* x = logits + r * nb;
  float m = -INFINITY;
  for (int k = lane; k < nb; k += 32) m = fmaxf(m, x[k]);
  m = wmax(m);
  float s = 0.f;
  for (int k = lane; k < nb; k += 32) s += expf(x[k] - m);
  const float lse = m + logf(wsum(s));
  const TwoHot a = twohot_warp(bins, nb, t1[r], lane);
  float out = -(a.wb * (x[a.below] - lse) + a.wa * (x[a.above] - lse));
  if (t2 != nullptr) {
    const TwoHot b = twohot_warp(bins, nb, t2[r], lane);
    out += w2 * -(b.wb * (x[b.below] - lse) + b.wa * (x[b.above] - lse));
  }
  if (lane == 0) {
    loss[r] = out;
    lse_out[r] = lse;
  }
}

__global__ void twohot_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ t1,
                                  const float* __restrict__ t2, float w2, const float* __restrict__ bins,
                                  const float* __restrict__ lse_in, const float* __restrict__ gloss,
                                  float* __restrict__ glogits, int64_t rows, int nb) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* x = logits + r * nb;
  float* gx = glogits + r * nb;
  const float g = gloss[r], lse = lse_in[r];
  const TwoHot a = twohot_warp(bins, nb, t1[r], lane);
  TwoHot b = a;
  float mass = 1.0f;
  if (t2 != nullptr) {
    b = twohot_warp(bins, nb, t2[r], lane);
    mass += w2;
  }
  for (int k = lane; k < nb; k += 32) {
    float th = (k == a.below ? a.wb : 0.f) + (k == a.above ? a.wa : 0.f);
    if (t2 != nullptr) th += w2 * ((k == b.below ? b.wb : 0.f) + (k == b.above ? b.wa : 0.f));
    gx[k] = g * (mass * expf(x[k] - lse) - th);
  }
}

__global__ void twohot_pred_kernel(const float* __restrict__ logits, const float* __restrict__ bins,
                                   float* __restrict__ pred, int64_t rows, int nb) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* x = logits + r * nb;
  float m = -INFINITY;
  for (int k = lane; k < nb; k += 32) m = fmaxf(m, x[k]);
  m = wmax(m);
  float s = 0.f;
  for (int k = lane; k < nb; k += 32) s += expf(x[k] - m);
  const float inv = 1.0f / wsum(s);
  // pairs (mid - 1 - i, mid' + i) first, then the butterfly: symmetric inputs cancel exactly
  const int half = nb / 2, hi0 = nb - half;       // odd nb: the middle bin is index half
  float acc = 0.f;
  for (int i = lane; i < half; i += 32) {
    const int lo = half - 1 - i, hi = hi0 + i;
    // no FMA contraction: p*b + p*(-b) must cancel exactly (outs.py:286-290)
    const float pl = __fmul_rn(__fmul_rn(expf(x[lo] - m), inv), bins[lo]);
    const float ph = __fmul_rn(__fmul_rn(expf(x[hi] - m), inv), bins[hi]);
    acc = __fadd_rn(acc, __fadd_rn(pl, ph));
  }
  acc = wsum(acc);
  if (lane == 0) {
    if (nb & 1) acc += expf(x[half] - m) * inv * bins[half];
    pred[r] = acc;
  }
}

// total = sum_i scale_i * mean(x_i); means[i] = mean(x_i).  One CTA per tensor, then CTA 0 of a
// second launch-free step: the last CTA to finish adds the scaled means up (threadfence counter).
constexpr int kMaxTerms = 16;
struct ReduceArgs {
  const float* ptr[kMaxTerms];
  long long count[kMaxTerms];
  float scale[kMaxTerms];
  int n;
};

__global__ void loss_reduce_kernel(const ReduceArgs a, float* __restrict__ means, float* __restrict__ total,
                                   unsigned* __restrict__ ticket) {
  __shared__ float red[32];
  __shared__ bool last;
  const int i = blockIdx.x;
  const float* x = a.ptr[i];
  float s = 0.f;
  for (long long k = threadIdx.x; k < a.count[i]; k += blockDim.x) s += x[k];
  s = wsum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    means[i] = t / (float)a.count[i];
    __threadfence();
    last = atomicAdd(ticket, 1u) == (unsigned)(a.n - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float t = 0.f;
    for (int j = 0; j < a.n; ++j) t += a.scale[j] * reinterpret_cast<volatile float*>(means)[j];   // fixed order
    *total = t;
    *ticket = 0;
  }
}

}  // namespace

extern "C" int emb_twohot_loss_fwd(const float* logits, const float* target, const float* target2,
                                   float weight2, const float* bins, float* loss, float* lse,
                                   int64_t rows, int32_t nbins, void* stream) {
  const char* who = "emb_twohot_loss_fwd";
  if (rows <= 0) return 0;
  if (!logits || !target || !bins || !loss || !lse) return emb::fail(-1, "%s: null pointer", who);
  if (nbins < 2 || nbins > 1024) return emb::fail(-1, "%s: nbins=%d outside [2, 1024]", who, nbins);
  const unsigned grid = (unsigned)((rows + kWarps - 1) / kWarps);
  twohot_fwd_kernel<<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(logits, target, target2, weight2, bins,
                                                                     loss, lse, rows, nbins);
  if (cudaGetLastError() != cudaSuccess) return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}

extern "C" int emb_twohot_loss_bwd(const float* logits, const float* target, const float* target2,
                                   float weight2, const float* bins, const float* lse,
                                   const float* gloss, float* glogits, int64_t rows, int32_t nbins,
                                   void* stream) {
  const char* who = "emb_twohot_loss_bwd";
  if (rows <= 0) return 0;
  if (!logits || !target || !bins || !lse || !gloss || !glogits) return emb::fail(-1, "%s: null pointer", who);
  if (nbins < 2 || nbins > 1024) return emb::fail(-1, "%s: nbins=%d outside [2, 1024]", who, nbins);
  const unsigned grid = (unsigned)((rows + kWarps - 1) / kWarps);
  twohot_bwd_kernel<<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(logits, target, target2, weight2, bins,
                                                                     lse, gloss, glogits, rows, nbins);
  if (cudaGetLastError() != cudaSuccess) return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}

extern "C" int emb_twohot_pred(const float* logits, const float* bins, float* pred, int64_t rows,
                               int32_t nbins, void* stream) {
  const char* who = "emb_twohot_pred";
  if (rows <= 0) return 0;
  if (!logits || !bins || !pred) return emb::fail(-1, "%s: null pointer", who);
  if (nbins < 2 || nbins > 1024) return emb::fail(-1, "%s: nbins=%d outside [2, 1024]", who, nbins);
  const unsigned grid = (unsigned)((rows + kWarps - 1) / kWarps);
  twohot_pred_kernel<<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(logits, bins, pred, rows, nbins);
  if (cudaGetLastError() != cudaSuccess) return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}

extern "C" int emb_loss_reduce(const float* const* terms, const int64_t* counts, const float* scales,
                               int32_t n, float* means, float* total, uint32_t* ticket, void* stream) {
  const char* who = "emb_loss_reduce";
  if (n <= 0 || n > kMaxTerms) return emb::fail(-1, "%s: %d terms outside [1, %d]", who, n, kMaxTerms);
  if (!terms || !counts || !scales || !means || !total || !ticket) return emb::fail(-1, "%s: null pointer", who);
  ReduceArgs a;
  a.n = n;
  for (int i = 0; i < n; ++i) {
    if (!terms[i] || counts[i] <= 0) return emb::fail(-1, "%s: term %d is empty", who, i);
    a.ptr[i] = terms[i];
    a.count[i] = counts[i];
    a.scale[i] = scales[i];
  }
  loss_reduce_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(a, means, total, ticket);
  if (cudaGetLastError() != cudaSuccess) return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}
