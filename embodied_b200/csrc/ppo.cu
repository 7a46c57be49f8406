// Kernels of the ppo agent (ppo/agent.py): the advantage recurrence of ppo_loss and the
// optax chain of Agent._make_opt on the flat parameter buffer.  Both are HBM-bound.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {

// ppo/agent.py:204-212: live = (1 - term) * (1 - 1 / hor), cont = (1 - last)(1 - term) * lam,
// delta_t = rew_{t+1} + live_{t+1} val_{t+1} - val_t, adv_t = delta_t + live_{t+1} cont_{t+1} adv_{t+1},
// tar_t = adv_t + val_t.  One thread per row walks the T - 1 steps backwards (T <= a few hundred;
// the rows of a batch are independent): 5 reads + 2 writes of 4 bytes per element.
__global__ void gae_kernel(const float* __restrict__ rew, const float* __restrict__ val,
                           const uint8_t* __restrict__ last, const uint8_t* __restrict__ term,
                           float* __restrict__ adv, float* __restrict__ tar, int64_t rows, int T,
                           float keep, float lam) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* rw = rew + r * T;
  const float* vl = val + r * T;
  const uint8_t* ls = last + r * T;
  const uint8_t* tm = term + r * T;
  float carry = 0.f;
  for (int t = T - 2; t >= 0; --t) {
    const float live = tm[t + 1] ? 0.f : keep;
    const float cont = (ls[t + 1] || tm[t + 1]) ? 0.f : lam;
    // the reference forms live * val and live * cont * adv as separate roundings (no fma)
    const float delta = __fsub_rn(__fadd_rn(rw[t + 1], __fmul_rn(live, vl[t + 1])), vl[t]);
    carry = __fadd_rn(delta, __fmul_rn(__fmul_rn(live, cont), carry));
    adv[r * (T - 1) + t] = carry;
    tar[r * (T - 1) + t] = __fadd_rn(carry, vl[t]);
  }
}

constexpr int kThreads = 256;

// pass 1: per-block sums of squares of the gradient (fixed order inside a block)
__global__ void adam_sumsq_kernel(const float* __restrict__ grad, int64_t n, float* __restrict__ partial) {
  __shared__ float red[kThreads / 32];
  float s = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 g = *reinterpret_cast<const float4*>(grad + i);
      s += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
    } else {
      for (int64_t j = i; j < n; ++j) s += grad[j] * grad[j];
    }
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

// pass 2: clip_by_global_norm -> scale_by_adam -> add_decayed_weights -> scale_by_learning_rate
// (optax; Agent._make_opt, ppo/agent.py:120-131).  Every block sums the partials in the same
// order (double), so all blocks -- and all ranks holding the same gradient -- use one factor.
// state[0] = Adam's count (also the schedule's count before this update), state[1] = the norm.
__global__ void adam_update_kernel(float* __restrict__ master, const float* __restrict__ grad,
                                   float* __restrict__ mu, float* __restrict__ nu,
                                   const float* __restrict__ wdmask, int64_t n,
                                   const float* __restrict__ partial, int nblocks_sumsq,
                                   float* __restrict__ state, float lr_peak, int warmup, float clip,
                                   float eps, float wd, float b1, float b2) {
  __shared__ float s_factor, s_lr, s_c1, s_c2;
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < nblocks_sumsq; ++i) tot += (double)partial[i];
    const float norm = (float)sqrt(tot);
    const float before = state[0];                    // updates applied so far
    s_factor = norm < clip ? 1.0f : clip / norm;
    s_lr = warmup > 0 ? lr_peak * fminf(before / (float)warmup, 1.0f) : lr_peak;
    s_c1 = 1.0f - powf(b1, before + 1.0f);            // bias corrections (divisors)
    s_c2 = 1.0f - powf(b2, before + 1.0f);
    if (blockIdx.x == 0) state[1] = norm;
  }
  __syncthreads();
  const float factor = s_factor, lr = s_lr, c1 = s_c1, c2 = s_c2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float g = grad[i] * factor;
    const float m = b1 * mu[i] + (1.0f - b1) * g;
    const float v = b2 * nu[i] + (1.0f - b2) * g * g;
    mu[i] = m;
    nu[i] = v;
    float upd = (m / c1) / (sqrtf(v / c2) + eps);
    const float p = master[i];
    if (wd != 0.f && wdmask && wdmask[i] != 0.f) upd += wd * p;
    master[i] = p - lr * upd;
  }
}

// after the update (stream order): count += 1
__global__ void adam_count_kernel(float* state) { state[0] += 1.0f; }

}  // namespace

extern "C" int emb_gae_advantage(const float* rew, const float* val, const uint8_t* last,
                                 const uint8_t* term, float* adv, float* tar, int64_t rows,
                                 int32_t length, float hor, float lam, void* stream) {
  const char* who = "emb_gae_advantage";
  if (rows < 0 || length < 2) return emb::fail(-1, "%s: rows=%lld length=%d", who, (long long)rows, length);
  if (rows == 0) return 0;
  if (!rew || !val || !last || !term || !adv || !tar) return emb::fail(-1, "%s: null pointer", who);
  if (!(hor > 0.f)) return emb::fail(-1, "%s: hor=%f", who, hor);
  const int threads = 128;
  gae_kernel<<<(unsigned)((rows + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      rew, val, last, term, adv, tar, rows, length, 1.0f - 1.0f / hor, lam);
  if (cudaGetLastError() != cudaSuccess) return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}

extern "C" int emb_opt_clip_adam(float* master, const float* grad, float* mu, float* nu,
                                 const float* wdmask, int64_t n, float* scratch, int32_t scratch_len,
                                 float* state, float lr, int32_t warmup, float clip, float eps,
                                 float wd, float b1, float b2, void* stream) {
  const char* who = "emb_opt_clip_adam";
  if (n <= 0) return emb::fail(-1, "%s: n=%lld", who, (long long)n);
  if (!master || !grad || !mu || !nu || !scratch || !state) return emb::fail(-1, "%s: null pointer", who);
  if (((uintptr_t)grad & 15) != 0) return emb::fail(-1, "%s: grad must be 16-byte aligned", who);
  if (scratch_len < 1) return emb::fail(-1, "%s: scratch_len=%d", who, scratch_len);
  if (wd != 0.f && !wdmask) return emb::fail(-1, "%s: weight decay needs wdmask", who);
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t want = (n / 4 + kThreads - 1) / kThreads;
  int blocks = (int)(want < 1 ? 1 : (want > sms * 4 ? sms * 4 : want));
  if (blocks > scratch_len) blocks = scratch_len;
  cudaStream_t s = (cudaStream_t)stream;
  adam_sumsq_kernel<<<blocks, kThreads, 0, s>>>(grad, n, scratch);
  int64_t want2 = (n + kThreads - 1) / kThreads;
  const int blocks2 = (int)(want2 > sms * 8 ? sms * 8 : want2);
  adam_update_kernel<<<blocks2, kThreads, 0, s>>>(master, grad, mu, nu, wdmask, n, scratch, blocks, state,
                                                  lr, warmup, clip, eps, wd, b1, b2);
  adam_count_kernel<<<1, 1, 0, s>>>(state);
  if (cudaGetLastError() != cudaSuccess) return emb::fail_cuda(who);
  emb::count_launch();
  emb::count_launch();
  emb::count_launch();
  return 0;
}
