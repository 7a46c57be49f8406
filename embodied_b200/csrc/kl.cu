// emb_rssm_kl_fwd / _bwd: the KL / free-nats / entropy reduction of RSSM.loss
// (dreamerv3/rssm.py:120-133) as ONE pass over the posterior and prior logits
// each way, instead of ~40 element-wise launches:
//
//   a = log(unimix(softmax(post)))   b = log(unimix(softmax(prior)))    outs.py:210-216
//   kl_s = sum_c softmax(a)_c (log_softmax(a)_c - log_softmax(b)_c)     outs.py:236-240
//   dyn = rep = max(sum_s kl_s, free_nats)  (they differ in where the gradient
//   goes: dyn -> prior only, rep -> posterior only; rssm.py:125-130)
//   ent_x = sum_s -sum_c softmax(x)_c log_softmax(x)_c                   rssm.py:131-132
//
// One CTA per (b, t) row, one warp per latent (strided when S > warps), the C <= 128
// classes of a latent spread over the lanes (<= 4 per lane): every reduction over
// classes is a warp-shuffle butterfly, the sum over latents goes through shared
// memory.  HBM traffic = the two logit tensors once (fwd) / once + the two
// gradient tensors (bwd).  Logits are fp32 or bf16 and may be strided views
// ((b, t) -> b*stride_b + t*stride_t, the S*C classes contiguous).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {

constexpr int kMaxPerLane = 4;      // C <= 128

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float load_logit(const void* p, int dtype, int64_t i) {
  return dtype ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
               : reinterpret_cast<const float*>(p)[i];
}

// One latent's distribution on a warp: soft = softmax(logit), mixed = unimix(soft),
// a = log(mixed), norm = softmax(a) = mixed / sum(mixed), la = log_softmax(a).
struct Dist {
  float soft[kMaxPerLane], mixed[kMaxPerLane], norm[kMaxPerLane], la[kMaxPerLane];
};

__device__ __forceinline__ void make_dist(const void* base, int dtype, int64_t off, int C,
                                          float unimix, int lane, Dist& d) {
  float x[kMaxPerLane];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    x[i] = c < C ? load_logit(base, dtype, off + c) : -INFINITY;
    m = fmaxf(m, x[i]);
  }
  m = warp_max(m);
  float z = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    d.soft[i] = lane + 32 * i < C ? expf(x[i] - m) : 0.f;
    z += d.soft[i];
  }
  z = warp_sum(z);
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const bool on = lane + 32 * i < C;
    d.soft[i] = d.soft[i] / z;
    d.mixed[i] = on ? (1.0f - unimix) * d.soft[i] + unimix / (float)C : 0.f;
    tot += d.mixed[i];
  }
  tot = warp_sum(tot);
  const float ltot = logf(tot);
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const bool on = lane + 32 * i < C;
    d.norm[i] = d.mixed[i] / tot;
    d.la[i] = on ? logf(d.mixed[i]) - ltot : 0.f;
  }
}

struct Args {
  const void* post;
  const void* prior;
  int dtype_post, dtype_prior;
  int B, T, S, C;
  int64_t post_sb, post_st, prior_sb, prior_st;
  float unimix, free_nats;
};

__global__ void __launch_bounds__(1024)
kl_fwd_kernel(Args a, float* __restrict__ dyn, float* __restrict__ rep, float* __restrict__ kl_raw,
              float* __restrict__ ent_post, float* __restrict__ ent_prior) {
  __shared__ float acc[3][32];
  const int row = blockIdx.x, b = row / a.T, t = row - b * a.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int64_t opost = (int64_t)b * a.post_sb + (int64_t)t * a.post_st;
  const int64_t oprior = (int64_t)b * a.prior_sb + (int64_t)t * a.prior_st;
  float kl = 0.f, ep = 0.f, eq = 0.f;
  for (int s = warp; s < a.S; s += nwarps) {
    Dist p, q;
    make_dist(a.post, a.dtype_post, opost + (int64_t)s * a.C, a.C, a.unimix, lane, p);
    make_dist(a.prior, a.dtype_prior, oprior + (int64_t)s * a.C, a.C, a.unimix, lane, q);
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      kl = fmaf(p.norm[i], p.la[i] - q.la[i], kl);
      ep = fmaf(-p.norm[i], p.la[i], ep);
      eq = fmaf(-q.norm[i], q.la[i], eq);
    }
  }
  kl = warp_sum(kl); ep = warp_sum(ep); eq = warp_sum(eq);
  if (lane == 0) { acc[0][warp] = kl; acc[1][warp] = ep; acc[2][warp] = eq; }
  __syncthreads();
  if (warp == 0) {
    float v0 = lane < nwarps ? acc[0][lane] : 0.f;
    float v1 = lane < nwarps ? acc[1][lane] : 0.f;
    float v2 = lane < nwarps ? acc[2][lane] : 0.f;
    v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2);
    if (lane == 0) {
      kl_raw[row] = v0;
      const float clipped = fmaxf(v0, a.free_nats);
      dyn[row] = clipped;
      rep[row] = clipped;
      ent_post[row] = v1;
      ent_prior[row] = v2;
    }
  }
}

// Gradients: with a = log(mixed(y)), KL = sum_c softmax(a)_c (la_c - lb_c):
//   dKL/da_j = pa_j (la_j - lb_j - kl_s)            dKL/db_j = qb_j - pa_j
//   da_j/dy_k = (1-u) t_j (delta_jk - t_k) / mixed_j  (t = softmax(y)), same for b(z)
__global__ void __launch_bounds__(1024)
kl_bwd_kernel(Args a, const float* __restrict__ kl_raw, const float* __restrict__ g_dyn,
              const float* __restrict__ g_rep, float* __restrict__ g_post,
              float* __restrict__ g_prior) {
  const int row = blockIdx.x, b = row / a.T, t = row - b * a.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int64_t opost = (int64_t)b * a.post_sb + (int64_t)t * a.post_st;
  const int64_t oprior = (int64_t)b * a.prior_sb + (int64_t)t * a.prior_st;
  // torch.clamp(min=free) passes the gradient where kl >= free
  const bool open = kl_raw[row] >= a.free_nats;
  const float gd = open ? g_dyn[row] : 0.f, gr = open ? g_rep[row] : 0.f;
  for (int s = warp; s < a.S; s += nwarps) {
    Dist p, q;
    make_dist(a.post, a.dtype_post, opost + (int64_t)s * a.C, a.C, a.unimix, lane, p);
    make_dist(a.prior, a.dtype_prior, oprior + (int64_t)s * a.C, a.C, a.unimix, lane, q);
    float kl = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) kl = fmaf(p.norm[i], p.la[i] - q.la[i], kl);
    kl = warp_sum(kl);
    float w[kMaxPerLane], v[kMaxPerLane], sw = 0.f, sv = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const bool on = lane + 32 * i < a.C;
      const float da = p.norm[i] * (p.la[i] - q.la[i] - kl);
      const float db = q.norm[i] - p.norm[i];
      w[i] = on ? da * p.soft[i] / p.mixed[i] : 0.f;
      v[i] = on ? db * q.soft[i] / q.mixed[i] : 0.f;
      sw += w[i]; sv += v[i];
    }
    sw = warp_sum(sw); sv = warp_sum(sv);
    const int64_t out = ((int64_t)row * a.S + s) * a.C;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < a.C) {
        g_post[out + c] = gr * (1.0f - a.unimix) * (w[i] - p.soft[i] * sw);
        g_prior[out + c] = gd * (1.0f - a.unimix) * (v[i] - q.soft[i] * sv);
      }
    }
  }
}

int validate(const char* who, const emb_rssm_kl_args* k) {
  if (!k) return emb::fail(-1, "%s: args is NULL", who);
  if (k->B < 0 || k->T < 0 || k->S < 1 || k->C < 1)
    return emb::fail(-1, "%s: B=%d T=%d S=%d C=%d", who, k->B, k->T, k->S, k->C);
  if (k->C > 32 * kMaxPerLane) return emb::fail(-1, "%s: classes=%d > %d", who, k->C, 32 * kMaxPerLane);
  if ((k->dtype_post | k->dtype_prior) & ~1)
    return emb::fail(-1, "%s: dtype must be 0 (f32) or 1 (bf16)", who);
  return 0;
}

Args make_args(const emb_rssm_kl_args* k) {
  return Args{k->post, k->prior, k->dtype_post, k->dtype_prior, k->B, k->T, k->S, k->C,
              k->post_stride_b, k->post_stride_t, k->prior_stride_b, k->prior_stride_t,
              k->unimix, k->free_nats};
}

unsigned threads_for(int S) { return 32u * (unsigned)(S < 32 ? S : 32); }

}  // namespace

extern "C" int emb_rssm_kl_fwd(const emb_rssm_kl_args* k, float* dyn, float* rep, float* kl_raw,
                               float* ent_post, float* ent_prior, void* stream) {
  const char* who = "emb_rssm_kl_fwd";
  if (int e = validate(who, k)) return e;
  const int64_t rows = (int64_t)k->B * k->T;
  if (rows == 0) return 0;
  kl_fwd_kernel<<<(unsigned)rows, threads_for(k->S), 0, (cudaStream_t)stream>>>(
      make_args(k), dyn, rep, kl_raw, ent_post, ent_prior);
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}

extern "C" int emb_rssm_kl_bwd(const emb_rssm_kl_args* k, const float* kl_raw, const float* g_dyn,
                               const float* g_rep, float* g_post, float* g_prior, void* stream) {
  const char* who = "emb_rssm_kl_bwd";
  if (int e = validate(who, k)) return e;
  const int64_t rows = (int64_t)k->B * k->T;
  if (rows == 0) return 0;
  kl_bwd_kernel<<<(unsigned)rows, threads_for(k->S), 0, (cudaStream_t)stream>>>(
      make_args(k), kl_raw, g_dyn, g_rep, g_post, g_prior);
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}

// ---------------------------------------------------------------- lambda returns
// emb_lambda_return: the reversed recurrence of dreamerv3/agent.py:482-490 as one
// launch (one thread per row) instead of 3 element-wise launches per time step:
//   live = (1 - term[:, 1:]) * disc ;  cont = (1 - last[:, 1:]) * lam
//   ret[:, t] = rew[:, t+1] + (1 - cont_t) live_t boot[:, t+1] + live_t cont_t ret[:, t+1],
//   ret[:, L-1] := boot[:, L-1]  (not stored).  All inputs fp32 [rows][L]; ret [rows][L-1].
namespace {
__global__ void lambda_return_kernel(const float* __restrict__ last, const float* __restrict__ term,
                                     const float* __restrict__ rew, const float* __restrict__ boot,
                                     float* __restrict__ ret, int64_t rows, int L, float disc, float lam) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* la = last + r * L;
  const float* te = term + r * L;
  const float* re = rew + r * L;
  const float* bo = boot + r * L;
  float acc = bo[L - 1];
  for (int t = L - 2; t >= 0; --t) {
    const float live = (1.0f - te[t + 1]) * disc;
    const float cont = (1.0f - la[t + 1]) * lam;
    acc = (re[t + 1] + (1.0f - cont) * live * bo[t + 1]) + live * cont * acc;
    ret[r * (L - 1) + t] = acc;
  }
}
}  // namespace

extern "C" int emb_lambda_return(const float* last, const float* term, const float* rew,
                                 const float* boot, float* ret, int64_t rows, int32_t length,
                                 float disc, float lam, void* stream) {
  const char* who = "emb_lambda_return";
  if (rows < 0 || length < 1) return emb::fail(-1, "%s: rows=%lld length=%d", who, (long long)rows, length);
  if (rows == 0 || length == 1) return 0;
  lambda_return_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      last, term, rew, boot, ret, rows, length, disc, lam);
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}

// ------------------------------------------------------------- one-hot sampling
// emb_onehot_sample: stoch = one_hot(argmax(log(unimix(softmax(logit))) + gumbel))
// (embodied/jax/outs.py:210-216,252-270, forward value of the straight-through
// sample) for the no-gradient paths (imagination, policy): one launch instead of
// softmax / unimix / log / add / argmax / one_hot / casts.  One CTA per row, one warp
// per latent (warp-shuffle max / sum / arg-max over the classes).
namespace {
__global__ void __launch_bounds__(1024)
onehot_sample_kernel(const void* __restrict__ logit, int dtype, int64_t logit_stride,
                     const float* __restrict__ gumbel, int64_t gumbel_stride, int S, int C,
                     float unimix,
                     void* __restrict__ out, int out_dtype, int64_t out_stride,
                     int32_t* __restrict__ index) {
  const int row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int s = warp; s < S; s += nwarps) {
    const int64_t off = (int64_t)row * logit_stride + (int64_t)s * C;
    const float* gn = gumbel + (int64_t)row * gumbel_stride + (int64_t)s * C;
    float x[kMaxPerLane];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      x[i] = c < C ? load_logit(logit, dtype, off + c) : -INFINITY;
      m = fmaxf(m, x[i]);
    }
    m = warp_max(m);
    float e[kMaxPerLane], z = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) { e[i] = lane + 32 * i < C ? expf(x[i] - m) : 0.f; z += e[i]; }
    z = warp_sum(z);
    float best = -INFINITY;
    int arg = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float v = logf((1.0f - unimix) * (e[i] / z) + unimix / (float)C) + gn[c];
        if (v > best) { best = v; arg = c; }
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (index && lane == 0) index[(int64_t)row * S + s] = arg;
    const int64_t oo = (int64_t)row * out_stride + (int64_t)s * C;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float v = c == arg ? 1.f : 0.f;
        if (out_dtype) reinterpret_cast<__nv_bfloat16*>(out)[oo + c] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(out)[oo + c] = v;
      }
    }
  }
}
}  // namespace

extern "C" int emb_onehot_sample(const void* logit, int32_t dtype, int64_t logit_stride,
                                 const float* gumbel, int64_t gumbel_stride, int64_t rows,
                                 int32_t S, int32_t C,
                                 float unimix, void* out, int32_t out_dtype, int64_t out_stride,
                                 int32_t* index, void* stream) {
  const char* who = "emb_onehot_sample";
  if (rows < 0 || S < 1 || C < 1 || C > 32 * kMaxPerLane)
    return emb::fail(-1, "%s: rows=%lld S=%d C=%d", who, (long long)rows, S, C);
  if ((dtype | out_dtype) & ~1) return emb::fail(-1, "%s: dtype must be 0 (f32) or 1 (bf16)", who);
  if (rows == 0) return 0;
  if (gumbel_stride == 0) gumbel_stride = (int64_t)S * C;
  onehot_sample_kernel<<<(unsigned)rows, threads_for(S), 0, (cudaStream_t)stream>>>(
      logit, dtype, logit_stride, gumbel, gumbel_stride, S, C, unimix, out, out_dtype, out_stride, index);
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}
