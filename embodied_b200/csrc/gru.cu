// Forward-only fusions of the block-GRU core (dreamerv3/rssm.py:147-158) for the
// batched paths that need no gradient (imagination, agent.py:188-200, and the
// policy): the batched GEMMs stay in the (group, row, column) layout cuBLAS
// produces, and the element-wise chains between them are one kernel each.
//
//   emb_rmsnorm_grouped_fwd:  y[g][m][:] = silu(rms_norm_row_m(x[g][m][:] + bias) * scale)
//       -- the norm runs over the FULL row (all groups), input and output stay
//       grouped, so neither dynhid0's output nor dyngru's input is transposed;
//   emb_gru_gates_fwd:        deter' = u * tanh(r * c) + (1 - u) * deter  with
//       r = sigmoid(pre_r + b), c = pre_c + b, u = sigmoid(pre_u + b - 1)
//       from pre[g][m][3*Dg] (gate-major inside a group, rssm.py:153-154).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxVec = 16;       // 16-byte vectors a thread may own of one row

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  __device__ static float round(float x) { return x; }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      v[2 * i] = __low2float(h); v[2 * i + 1] = __high2float(h);
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __device__ static float round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
};

__device__ __forceinline__ float silu(float n) { return n / (1.0f + expf(-n)); }
__device__ __forceinline__ float sigmoid(float n) { return 1.0f / (1.0f + expf(-n)); }
// bf16 outputs (8 mantissa bits): fast intrinsics are far inside the rounding error
template <typename T> __device__ __forceinline__ float sigmoid_t(float n) { return sigmoid(n); }
template <> __device__ __forceinline__ float sigmoid_t<__nv_bfloat16>(float n) {
  return __fdividef(1.0f, 1.0f + __expf(-n));
}
template <typename T> __device__ __forceinline__ float tanh_t(float n) { return tanhf(n); }
template <> __device__ __forceinline__ float tanh_t<__nv_bfloat16>(float n) {
  const float e = __expf(-2.0f * fabsf(n));
  return copysignf(__fdividef(1.0f - e, 1.0f + e), n);
}

// one CTA per row m; column c = gi*Dg + j lives at x[(gi*M + m)*Dg + j].  NV = vectors a
// thread owns (compile-time: the row stays in registers at a register count that lets
// several CTAs share an SM -- the kernel is latency-bound, not bandwidth-bound).
template <typename T, int NV>
__global__ void __launch_bounds__(kThreads)
rmsnorm_grouped_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                       const float* __restrict__ bias, T* __restrict__ y, int M, int G, int Dg,
                       int act, float eps) {
  constexpr int N = Vec<T>::N;
  __shared__ float red[kThreads / 32];
  const int m = blockIdx.x, D = G * Dg, nvec = D / N;
  float v[NV][N];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = threadIdx.x + k * kThreads;
    if (vi < nvec) {
      const int c = vi * N, gi = c / Dg, j = c - gi * Dg;
      Vec<T>::load(x + ((size_t)gi * M + m) * Dg + j, v[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = threadIdx.x + k * kThreads;
    if (vi < nvec) {
      const int c = vi * N;
#pragma unroll
      for (int i = 0; i < N; i += 4) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b = *reinterpret_cast<const float4*>(bias + c + i);
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (bias) v[k][i + q] = Vec<T>::round(v[k][i + q] + bb[q]);
          ss = fmaf(v[k][i + q], v[k][i + q], ss);
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) tot += red[w];
  const float rstd = rsqrtf(tot / (float)D + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = threadIdx.x + k * kThreads;
    if (vi < nvec) {
      const int c = vi * N, gi = c / Dg, j = c - gi * Dg;
      float o[N];
#pragma unroll
      for (int i = 0; i < N; i += 4) {
        const float4 s4 = *reinterpret_cast<const float4*>(scale + c + i);
        const float sc[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float n = Vec<T>::round(v[k][i + q] * (rstd * sc[q]));
          o[i + q] = act ? silu(n) : n;
        }
      }
      Vec<T>::store(y + ((size_t)gi * M + m) * Dg + j, o);
    }
  }
}

template <typename T>
void launch_grouped(int nv, const void* x, const float* scale, const float* bias, void* y, int m, int g,
                    int dg, int act, float eps, cudaStream_t s) {
#define EMB_NV(NV)                                                                              \
  rmsnorm_grouped_kernel<T, NV><<<(unsigned)m, kThreads, 0, s>>>((const T*)x, scale, bias, (T*)y, m, \
                                                                  g, dg, act, eps)
  if (nv <= 1) EMB_NV(1);
  else if (nv <= 2) EMB_NV(2);
  else if (nv <= 4) EMB_NV(4);
  else if (nv <= 8) EMB_NV(8);
  else EMB_NV(16);
#undef EMB_NV
}

// one thread per vector of the (M, D) state
template <typename T>
__global__ void __launch_bounds__(kThreads)
gru_gates_kernel(const T* __restrict__ pre, const float* __restrict__ bias, const T* __restrict__ deter,
                 T* __restrict__ out, int64_t nvec, int M, int G, int Dg, int64_t dstride,
                 int64_t ostride) {
  constexpr int N = Vec<T>::N;
  const int vpr = G * Dg / N;
  for (int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x; o < nvec;
       o += (int64_t)gridDim.x * kThreads) {
    const int m = (int)(o / vpr);
    const int c = (int)(o - (int64_t)m * vpr) * N, gi = c / Dg, j = c - gi * Dg;
    const T* p = pre + ((size_t)gi * M + m) * 3 * Dg + j;
    const float* b = bias + (size_t)gi * 3 * Dg + j;
    float r[N], cd[N], u[N], d[N], res[N];
    Vec<T>::load(p, r);
    Vec<T>::load(p + Dg, cd);
    Vec<T>::load(p + 2 * Dg, u);
    Vec<T>::load(deter + (size_t)m * dstride + c, d);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float rs = sigmoid_t<T>(Vec<T>::round(r[i] + b[i]));
      const float cand = tanh_t<T>(Vec<T>::round(rs * Vec<T>::round(cd[i] + b[Dg + i])));
      const float up = sigmoid_t<T>(Vec<T>::round(u[i] + b[2 * Dg + i]) - 1.0f);
      res[i] = up * cand + (1.0f - up) * d[i];
    }
    Vec<T>::store(out + (size_t)m * ostride + c, res);
  }
}

int g_sms = 0;

int common(const char* who, int64_t m, int g, int dg, int dtype) {
  if (m < 0 || g < 1 || dg < 1) return emb::fail(-1, "%s: m=%lld g=%d dg=%d", who, (long long)m, g, dg);
  if (dtype != 0 && dtype != 1) return emb::fail(-1, "%s: dtype %d (0 = f32, 1 = bf16)", who, dtype);
  if (dg % (dtype ? 8 : 4)) return emb::fail(-1, "%s: dg=%d must be a multiple of %d", who, dg, dtype ? 8 : 4);
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return emb::fail_cuda(who);
  }
  return 0;
}

}  // namespace

extern "C" int emb_rmsnorm_grouped_fwd(const void* x, const float* scale, const float* bias, void* y,
                                       int64_t m, int32_t g, int32_t dg, int32_t dtype, int32_t act,
                                       float eps, void* stream) {
  const char* who = "emb_rmsnorm_grouped_fwd";
  if (int e = common(who, m, g, dg, dtype)) return e;
  const int n = dtype ? 8 : 4;
  if ((int64_t)g * dg / n > (int64_t)kMaxVec * kThreads)
    return emb::fail(-1, "%s: row of %d columns exceeds %d", who, g * dg, kMaxVec * kThreads * n);
  if (m == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int nv = (int)(((int64_t)g * dg / n + kThreads - 1) / kThreads);
  if (dtype) launch_grouped<__nv_bfloat16>(nv, x, scale, bias, y, (int)m, g, dg, act, eps, s);
  else launch_grouped<float>(nv, x, scale, bias, y, (int)m, g, dg, act, eps, s);
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}

extern "C" int emb_gru_gates_fwd(const void* pre, const float* bias, const void* deter, void* out,
                                 int64_t m, int32_t g, int32_t dg, int32_t dtype,
                                 int64_t deter_stride, int64_t out_stride, void* stream) {
  const char* who = "emb_gru_gates_fwd";
  if (int e = common(who, m, g, dg, dtype)) return e;
  if (m == 0) return 0;
  const int64_t width = (int64_t)g * dg, per = dtype ? 8 : 4;
  if (deter_stride == 0) deter_stride = width;
  if (out_stride == 0) out_stride = width;
  if (deter_stride < width || out_stride < width || deter_stride % per || out_stride % per)
    return emb::fail(-1, "%s: row strides %lld / %lld (>= %lld, multiples of %lld)", who,
                     (long long)deter_stride, (long long)out_stride, (long long)width, (long long)per);
  if (((uintptr_t)deter | (uintptr_t)out | (uintptr_t)pre) & 15)
    return emb::fail(-1, "%s: pointers must be 16-byte aligned", who);
  const int64_t nvec = m * g * dg / (dtype ? 8 : 4);
  const int64_t want = (nvec + kThreads - 1) / kThreads, cap = (int64_t)g_sms * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype)
    gru_gates_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(
        (const __nv_bfloat16*)pre, bias, (const __nv_bfloat16*)deter, (__nv_bfloat16*)out, nvec, (int)m, g, dg,
        deter_stride, out_stride);
  else
    gru_gates_kernel<float><<<grid, kThreads, 0, s>>>(
        (const float*)pre, bias, (const float*)deter, (float*)out, nvec, (int)m, g, dg, deter_stride,
        out_stride);
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}
