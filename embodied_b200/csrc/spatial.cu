// 2x2 max-pool and 2x nearest up-sampling on NHWC tensors, forward and backward,
// one HBM pass each: the spatial glue between the 5x5 convolutions of the
// dreamerv3 encoder / decoder (dreamerv3/rssm.py:239-240
// `x.reshape(B, H//2, 2, W//2, 2, C).max((2, 4))`, :336,349 `x.repeat(2,-2).repeat(2,-3)`).
// 16-byte accesses, persistent grid of 8 x SMs CTAs.  The pool keeps a 2-bit
// argmax per element (first maximum in (dy, dx) row-major order) instead of
// int64 indices: 0.25 bytes per output element.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T> struct V;
template <> struct V<float> {
  static constexpr int N = 4;
  using Idx = uint8_t;
  __device__ static void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct V<__nv_bfloat16> {
  static constexpr int N = 8;
  using Idx = uint16_t;
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
};

// y[n, h, w, :] = max over the 2x2 window of x[n, 2h.., 2w.., :]   (x: [N, 2H, 2W, C])
template <typename T>
__global__ void __launch_bounds__(kThreads)
pool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, typename V<T>::Idx* __restrict__ idx,
                int64_t nvec, int H, int W, int Cv) {
  constexpr int N = V<T>::N;
  for (int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x; o < nvec;
       o += (int64_t)gridDim.x * kThreads) {
    const int cv = (int)(o % Cv);
    int64_t p = o / Cv;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int64_t n = p / H;
    const int64_t row = (int64_t)2 * W * Cv;
    const T* base = x + ((n * 2 * H + 2 * h) * (int64_t)(2 * W) + 2 * w) * Cv * N + (int64_t)cv * N;
    float v[4][N];
    V<T>::load(base, v[0]);
    V<T>::load(base + (int64_t)Cv * N, v[1]);
    V<T>::load(base + row * N, v[2]);
    V<T>::load(base + row * N + (int64_t)Cv * N, v[3]);
    float m[N];
    uint32_t code = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float best = v[0][i];
      uint32_t at = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k][i] > best) { best = v[k][i]; at = k; }
      m[i] = best;
      code |= at << (2 * i);
    }
    V<T>::store(y + o * N, m);
    idx[o] = (typename V<T>::Idx)code;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
pool_bwd_kernel(const T* __restrict__ gy, const typename V<T>::Idx* __restrict__ idx,
                T* __restrict__ gx, int64_t nvec, int H, int W, int Cv) {
  constexpr int N = V<T>::N;
  for (int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x; o < nvec;
       o += (int64_t)gridDim.x * kThreads) {
    const int cv = (int)(o % Cv);
    int64_t p = o / Cv;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int64_t n = p / H;
    const int64_t row = (int64_t)2 * W * Cv;
    T* base = gx + ((n * 2 * H + 2 * h) * (int64_t)(2 * W) + 2 * w) * Cv * N + (int64_t)cv * N;
    float g[N];
    V<T>::load(gy + o * N, g);
    const uint32_t code = idx[o];
    float out[4][N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const uint32_t at = (code >> (2 * i)) & 3u;
#pragma unroll
      for (int k = 0; k < 4; ++k) out[k][i] = at == (uint32_t)k ? g[i] : 0.f;
    }
    V<T>::store(base, out[0]);
    V<T>::store(base + (int64_t)Cv * N, out[1]);
    V<T>::store(base + row * N, out[2]);
    V<T>::store(base + row * N + (int64_t)Cv * N, out[3]);
  }
}

// y[n, oh, ow, :] = x[n, oh/2, ow/2, :]     (x: [N, H, W, C], y: [N, 2H, 2W, C]); one thread per INPUT vector
template <typename T>
__global__ void __launch_bounds__(kThreads)
up_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t nvec, int H, int W, int Cv) {
  constexpr int N = V<T>::N;
  for (int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x; o < nvec;
       o += (int64_t)gridDim.x * kThreads) {
    const int cv = (int)(o % Cv);
    int64_t p = o / Cv;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int64_t n = p / H;
    const int64_t row = (int64_t)2 * W * Cv;
    T* base = y + ((n * 2 * H + 2 * h) * (int64_t)(2 * W) + 2 * w) * Cv * N + (int64_t)cv * N;
    const uint4 v = *reinterpret_cast<const uint4*>(x + o * N);
    *reinterpret_cast<uint4*>(base) = v;
    *reinterpret_cast<uint4*>(base + (int64_t)Cv * N) = v;
    *reinterpret_cast<uint4*>(base + row * N) = v;
    *reinterpret_cast<uint4*>(base + row * N + (int64_t)Cv * N) = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
up_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, int64_t nvec, int H, int W, int Cv) {
  constexpr int N = V<T>::N;
  for (int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x; o < nvec;
       o += (int64_t)gridDim.x * kThreads) {
    const int cv = (int)(o % Cv);
    int64_t p = o / Cv;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int64_t n = p / H;
    const int64_t row = (int64_t)2 * W * Cv;
    const T* base = gy + ((n * 2 * H + 2 * h) * (int64_t)(2 * W) + 2 * w) * Cv * N + (int64_t)cv * N;
    float a[N], b[N], c[N], d[N], s[N];
    V<T>::load(base, a);
    V<T>::load(base + (int64_t)Cv * N, b);
    V<T>::load(base + row * N, c);
    V<T>::load(base + row * N + (int64_t)Cv * N, d);
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = (a[i] + b[i]) + (c[i] + d[i]);
    V<T>::store(gx + o * N, s);
  }
}

int g_sms = 0;

int prepare(const char* who, int64_t n, int h, int w, int c, int dtype, int64_t* nvec, int* cv,
            unsigned* grid) {
  if (n < 0 || h <= 0 || w <= 0 || c <= 0) return emb::fail(-1, "%s: bad shape", who);
  if (dtype != 0 && dtype != 1) return emb::fail(-1, "%s: dtype %d (0 = f32, 1 = bf16)", who, dtype);
  const int per = dtype ? 8 : 4;
  if (c % per) return emb::fail(-1, "%s: channels=%d must be a multiple of %d", who, c, per);
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return emb::fail_cuda(who);
  }
  *cv = c / per;
  *nvec = n * h * w * (int64_t)*cv;
  const int64_t want = (*nvec + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)g_sms * 8;
  *grid = (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
  return 0;
}

}  // namespace

extern "C" int emb_maxpool2_nhwc_fwd(const void* x, void* y, void* idx, int64_t n, int32_t h,
                                     int32_t w, int32_t c, int32_t dtype, void* stream) {
  const char* who = "emb_maxpool2_nhwc_fwd";
  int64_t nvec; int cv; unsigned grid;
  if (int e = prepare(who, n, h, w, c, dtype, &nvec, &cv, &grid)) return e;
  if (nvec == 0) return 0;
  if (dtype) {
    pool_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (uint16_t*)idx, nvec, h, w, cv);
  } else {
    pool_fwd_kernel<float><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        (const float*)x, (float*)y, (uint8_t*)idx, nvec, h, w, cv);
  }
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}

extern "C" int emb_maxpool2_nhwc_bwd(const void* gy, const void* idx, void* gx, int64_t n, int32_t h,
                                     int32_t w, int32_t c, int32_t dtype, void* stream) {
  const char* who = "emb_maxpool2_nhwc_bwd";
  int64_t nvec; int cv; unsigned grid;
  if (int e = prepare(who, n, h, w, c, dtype, &nvec, &cv, &grid)) return e;
  if (nvec == 0) return 0;
  if (dtype) {
    pool_bwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)gy, (const uint16_t*)idx, (__nv_bfloat16*)gx, nvec, h, w, cv);
  } else {
    pool_bwd_kernel<float><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        (const float*)gy, (const uint8_t*)idx, (float*)gx, nvec, h, w, cv);
  }
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}

extern "C" int emb_upsample2_nhwc_fwd(const void* x, void* y, int64_t n, int32_t h, int32_t w,
                                      int32_t c, int32_t dtype, void* stream) {
  const char* who = "emb_upsample2_nhwc_fwd";
  int64_t nvec; int cv; unsigned grid;
  if (int e = prepare(who, n, h, w, c, dtype, &nvec, &cv, &grid)) return e;
  if (nvec == 0) return 0;
  if (dtype) {
    up_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, nvec, h, w, cv);
  } else {
    up_fwd_kernel<float><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        (const float*)x, (float*)y, nvec, h, w, cv);
  }
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}

extern "C" int emb_upsample2_nhwc_bwd(const void* gy, void* gx, int64_t n, int32_t h, int32_t w,
                                      int32_t c, int32_t dtype, void* stream) {
  const char* who = "emb_upsample2_nhwc_bwd";
  int64_t nvec; int cv; unsigned grid;
  if (int e = prepare(who, n, h, w, c, dtype, &nvec, &cv, &grid)) return e;
  if (nvec == 0) return 0;
  if (dtype) {
    up_bwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)gy, (__nv_bfloat16*)gx, nvec, h, w, cv);
  } else {
    up_bwd_kernel<float><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        (const float*)gy, (float*)gx, nvec, h, w, cv);
  }
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}
