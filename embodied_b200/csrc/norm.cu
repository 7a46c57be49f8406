// emb_rmsnorm_act_fwd / _bwd: rms-norm (+ silu) over the last axis as ONE pass
// over HBM each way (embodied/jax/nets.py:361-399 `Norm('rms')` followed by
// nets.act('silu'), the pattern after every Linear / Conv2D of dreamerv3).
//
//   fwd:  n = x * rsqrt(mean(x^2) + eps) * scale ;  y = act(cast(n))
//   bwd:  g_n = g_y * act'(n) ;  g_scale += sum_rows g_n * xhat ;
//         g_x = rstd * scale * g_n - xhat * rstd * mean(g_n * scale * xhat)
//
// HBM-bound: fwd reads x once and writes y once, bwd reads x and g_y once and
// writes g_x once (the second and third sweep over a row hit L1/L2: a row is at
// most 16 KiB).  One warp per row, 16-byte accesses, persistent grid of
// 2 x SMs CTAs; the per-column scale gradient is reduced warp -> CTA (shared
// memory) -> global (one atomicAdd per column per CTA).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxCols = 2048;          // bwd keeps cols/32 scale-gradient partials per lane

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
  __device__ static float silu(float n) { return n / (1.0f + expf(-n)); }
  __device__ static float dsilu(float n) {
    const float sg = 1.0f / (1.0f + expf(-n));
    return sg * (1.0f + n * (1.0f - sg));
  }
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
  // results are rounded to bf16 (2^-9): the fast exponential and reciprocal are exact enough
  __device__ static float silu(float n) { return __fdividef(n, 1.0f + __expf(-n)); }
  __device__ static float dsilu(float n) {
    const float sg = __fdividef(1.0f, 1.0f + __expf(-n));
    return sg * fmaf(n, 1.0f - sg, 1.0f);
  }
};

// sum over the `gs` (power of two) consecutive lanes that share a row
__device__ __forceinline__ float group_sum(float s, int gs) {
  for (int o = gs >> 1; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}
// the same sum without a data-dependent loop: five shuffles, strides >= gs masked out
__device__ __forceinline__ float group_sum_flat(float s, int gs) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float other = __shfl_xor_sync(0xffffffffu, s, o);
    s += o < gs ? other : 0.f;
  }
  return s;
}
// lanes per row: the smallest power of two covering cols / N vectors, at most a warp
__device__ __forceinline__ int group_size(int cols, int n) {
  int gs = 1;
  while (gs < 32 && gs * n < cols) gs <<= 1;
  return gs;
}


template <typename T>
__global__ void __launch_bounds__(kThreads)
rmsnorm_act_fwd_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                       const float* __restrict__ bias, T* __restrict__ y,
                       int64_t rows, int cols, int act, float eps) {
  constexpr int N = Vec<T>::N;
  const int gs = group_size(cols, N), rpw = 32 / gs;          // narrow rows: several per warp
  const int lane = threadIdx.x & 31, lg = lane % gs, grp = lane / gs;
  const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  for (int64_t r0 = warp * rpw; r0 < rows; r0 += nwarps * rpw) {
    const int64_t r = r0 + grp;
    const bool live = r < rows;
    const T* xr = x + (live ? r : 0) * cols;
    T* yr = y + (live ? r : 0) * cols;
    float ss = 0.f;
    for (int c = lg * N; c < cols; c += gs * N) {
      float v[N];
      Vec<T>::load(xr + c, v);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (bias) v[i] += bias[c + i];
        ss = fmaf(v[i], v[i], ss);
      }
    }
    const float rstd = rsqrtf(group_sum(ss, gs) / (float)cols + eps);
    if (!live) continue;
    for (int c = lg * N; c < cols; c += gs * N) {
      float v[N], o[N];
      Vec<T>::load(xr + c, v);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (bias) v[i] += bias[c + i];
        const float n = Vec<T>::round(v[i] * (rstd * scale[c + i]));   // cast back, then act (nets.py:397)
        o[i] = act ? Vec<T>::silu(n) : n;
      }
      Vec<T>::store(yr + c, o);
    }
  }
}

// Rows of up to 32 * NV vectors (dense layers: 1024 .. 2048 columns): one warp per row,
// the row read ONCE into registers (the generic kernel above sweeps it twice).
template <typename T, int NV>
__global__ void __launch_bounds__(kThreads)
rmsnorm_act_fwd_reg_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                           const float* __restrict__ bias, T* __restrict__ y,
                           int64_t rows, int cols, int act, float eps) {
  constexpr int N = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const T* xr = x + r * cols;
    T* yr = y + r * cols;
    float v[NV][N];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (lane + 32 * k) * N;
      if (c < cols) Vec<T>::load(xr + c, v[k]);
    }
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (lane + 32 * k) * N;
      if (c < cols) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          if (bias) v[k][i] += bias[c + i];
          ss = fmaf(v[k][i], v[k][i], ss);
        }
      }
    }
    const float rstd = rsqrtf(group_sum(ss, 32) / (float)cols + eps);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (lane + 32 * k) * N;
      if (c < cols) {
        float o[N];
#pragma unroll
        for (int i = 0; i < N; i += 4) {
          const float4 s4 = *reinterpret_cast<const float4*>(scale + c + i);
          const float sc[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float n = Vec<T>::round(v[k][i + q] * (rstd * sc[q]));
            o[i + q] = act ? Vec<T>::silu(n) : n;
          }
        }
        Vec<T>::store(yr + c, o);
      }
    }
  }
}

// Short rows (<= 256 columns: the channel axis of the convolutions, millions of rows):
// gs = power-of-two lanes per row, 32 / gs rows per warp, NV vectors per lane, the row
// read once into registers; bias of the producing convolution folded in.
template <typename T, int NV>
__global__ void __launch_bounds__(kThreads)
rmsnorm_act_fwd_short_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                             const float* __restrict__ bias, T* __restrict__ y,
                             int64_t rows, int cols, int act, float eps) {
  constexpr int N = Vec<T>::N;
  const int gs = group_size((cols + NV - 1) / NV, N), rpw = 32 / gs;
  const int lane = threadIdx.x & 31, lg = lane % gs, grp = lane / gs;
  const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  float sc[NV][N], bi[NV][N];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (lg + gs * k) * N;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      sc[k][i] = c < cols ? scale[c + i] : 0.f;
      bi[k][i] = (bias && c < cols) ? bias[c + i] : 0.f;
    }
  }
  constexpr int U = NV == 1 ? 2 : 1;      // row sets in flight per warp iteration
  const float inv_cols = 1.0f / (float)cols;
  for (int64_t r0 = warp * rpw * U; r0 < rows; r0 += nwarps * rpw * U) {
    float v[U][NV][N], rstd[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * rpw + grp;
      const T* xr = x + (r < rows ? r : 0) * cols;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (lg + gs * k) * N;
        if (c < cols) Vec<T>::load(xr + c, v[u][k]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (lg + gs * k) * N;
        if (c < cols) {
#pragma unroll
          for (int i = 0; i < N; ++i) { v[u][k][i] += bi[k][i]; ss = fmaf(v[u][k][i], v[u][k][i], ss); }
        }
      }
      rstd[u] = rsqrtf(group_sum_flat(ss, gs) * inv_cols + eps);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * rpw + grp;
      if (r >= rows) continue;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (lg + gs * k) * N;
        if (c < cols) {
          float o[N];
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const float n = Vec<T>::round(v[u][k][i] * (rstd[u] * sc[k][i]));
            o[i] = act ? Vec<T>::silu(n) : n;
          }
          Vec<T>::store(y + r * cols + c, o);
        }
      }
    }
  }
}

// PL = elements a lane may own of one row.  PL = 8 (cols <= 256: the channel
// axis of the convolutions, millions of short rows) keeps the row in registers
// -- one load of x and g_y, two shuffle reductions, one store -- runs at full
// occupancy and also reduces the bias gradient; PL = 64 (cols <= 2048: dense
// layers, ~1e3 long rows) re-reads the row from L1 between the sweeps.
template <typename T, int PL>
__global__ void __launch_bounds__(kThreads, PL <= 8 ? 3 : (PL <= 32 ? 2 : 1))
rmsnorm_act_bwd_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                       const float* __restrict__ bias, const T* __restrict__ gy,
                       T* __restrict__ gx, float* __restrict__ gscale, float* __restrict__ gbias,
                       int64_t rows, int cols, int act, float eps) {
  constexpr int N = Vec<T>::N;
  constexpr bool SMALL = PL <= 32;         // the row (x and g_y) stays in registers
  constexpr bool BIAS = PL <= 8;           // bias gradient (convolution channels only)
  constexpr int U = PL <= 8 ? 2 : 1;       // row sets in flight per warp iteration
  extern __shared__ float part_raw[];      // [kWarps][cols] (x2 with a bias)
  const int gs_ = group_size(cols, N), rpw = 32 / gs_;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, lg = lane % gs_, grp = lane / gs_;
  const int64_t warp = (int64_t)blockIdx.x * kWarps + wid;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  float gsc[PL];                           // scale-gradient partials of this lane's columns
  float gbi[BIAS ? PL : 1];                // bias-gradient partials
#pragma unroll
  for (int i = 0; i < PL; ++i) gsc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < (BIAS ? PL : 1); ++i) gbi[i] = 0.f;
  const float inv_cols = 1.0f / (float)cols;
  // convolution channels: scale and bias of this lane's columns live in registers
  float scr[BIAS ? PL : 1], bir[BIAS ? PL : 1];
  if (BIAS) {
#pragma unroll
    for (int j = 0; j < PL / N; ++j) {
      const int c = lg * N + j * gs_ * N;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        scr[j * N + i] = c < cols ? scale[c + i] : 0.f;
        bir[j * N + i] = (bias && c < cols) ? bias[c + i] : 0.f;
      }
    }
  }
  for (int64_t r0 = warp * rpw * U; r0 < rows; r0 += nwarps * rpw * U) {
    const int64_t r = r0 + grp;
    const bool live = r < rows;
    const T* xr = x + (live ? r : 0) * cols;
    const T* gr = gy + (live ? r : 0) * cols;
    T* or_ = gx + (live ? r : 0) * cols;
    if (SMALL) {
      // U row sets per iteration: all loads are issued before the first reduction
      float v[U][PL], g[U][PL], rstd[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t ru = r0 + u * rpw + grp;
        const int64_t rr = ru < rows ? ru : 0;
#pragma unroll
        for (int j = 0; j < PL / N; ++j) {
          const int c = lg * N + j * gs_ * N;
          if (c < cols) {
            Vec<T>::load(x + rr * cols + c, *reinterpret_cast<float(*)[N]>(&v[u][j * N]));
            Vec<T>::load(gy + rr * cols + c, *reinterpret_cast<float(*)[N]>(&g[u][j * N]));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < PL / N; ++j) {
          const int c = lg * N + j * gs_ * N;
          if (c < cols) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
              if (BIAS) v[u][j * N + i] += bir[j * N + i];
              ss = fmaf(v[u][j * N + i], v[u][j * N + i], ss);
            }
          }
        }
        rstd[u] = rsqrtf(group_sum_flat(ss, gs_) * inv_cols + eps);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool live_u = r0 + u * rpw + grp < rows;
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < PL / N; ++j) {
          const int c = lg * N + j * gs_ * N;
          if (c < cols) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
              const float xh = v[u][j * N + i] * rstd[u];
              const float sc = BIAS ? scr[j * N + i] : scale[c + i];
              const float gn = act ? g[u][j * N + i] * Vec<T>::dsilu(Vec<T>::round(xh * sc))
                                   : g[u][j * N + i];
              if (live_u) gsc[j * N + i] = fmaf(gn, xh, gsc[j * N + i]);
              dot = fmaf(gn * sc, xh, dot);
              v[u][j * N + i] = xh;
              g[u][j * N + i] = gn * sc;
            }
          }
        }
        const float mean = group_sum_flat(dot, gs_) * inv_cols;
        if (live_u) {
          T* out = gx + (r0 + u * rpw + grp) * cols;
#pragma unroll
          for (int j = 0; j < PL / N; ++j) {
            const int c = lg * N + j * gs_ * N;
            if (c < cols) {
              float o[N];
#pragma unroll
              for (int i = 0; i < N; ++i) {
                o[i] = rstd[u] * (g[u][j * N + i] - v[u][j * N + i] * mean);
                if (BIAS) gbi[j * N + i] += o[i];
              }
              Vec<T>::store(out + c, o);
            }
          }
        }
      }
      continue;
    }
    float ss = 0.f;
    for (int c = lg * N; c < cols; c += gs_ * N) {
      float v[N];
      Vec<T>::load(xr + c, v);
#pragma unroll
      for (int i = 0; i < N; ++i) ss = fmaf(v[i], v[i], ss);
    }
    const float rstd = rsqrtf(group_sum(ss, gs_) / (float)cols + eps);
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < PL / N; ++j) {
      const int c = lg * N + j * gs_ * N;
      if (c < cols) {
        float v[N], g[N];
        Vec<T>::load(xr + c, v);
        Vec<T>::load(gr + c, g);
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const float xh = v[i] * rstd, sc = scale[c + i];
          const float gn = act ? g[i] * Vec<T>::dsilu(Vec<T>::round(xh * sc)) : g[i];
          if (live) gsc[j * N + i] = fmaf(gn, xh, gsc[j * N + i]);
          dot = fmaf(gn * sc, xh, dot);
        }
      }
    }
    const float mean = group_sum(dot, gs_) / (float)cols;
    if (!live) continue;
    for (int c = lg * N; c < cols; c += gs_ * N) {
      float v[N], g[N], o[N];
      Vec<T>::load(xr + c, v);
      Vec<T>::load(gr + c, g);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const float xh = v[i] * rstd, sc = scale[c + i];
        const float gn = act ? g[i] * Vec<T>::dsilu(Vec<T>::round(xh * sc)) : g[i];
        o[i] = rstd * (sc * gn - xh * mean);
      }
      Vec<T>::store(or_ + c, o);
    }
  }
  // scale (and bias) gradient: row groups of a warp -> lane group 0 (shuffles) ->
  // shared [warp][col] -> one atomicAdd per column per CTA
  float* part_b = part_raw + (size_t)kWarps * cols;
#pragma unroll
  for (int j = 0; j < PL / N; ++j) {
    const int c = lg * N + j * gs_ * N;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float v = gsc[j * N + i];
      for (int o = gs_; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (c < cols && grp == 0) part_raw[(size_t)wid * cols + c + i] = v;
      if (BIAS && gbias) {
        float w = gbi[j * N + i];
        for (int o = gs_; o < 32; o <<= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (c < cols && grp == 0) part_b[(size_t)wid * cols + c + i] = w;
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += part_raw[(size_t)w * cols + c];
    atomicAdd(gscale + c, s);
    if (BIAS && gbias) {
      float sb = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) sb += part_b[(size_t)w * cols + c];
      atomicAdd(gbias + c, sb);
    }
  }
}

int g_sms = 0;

int check(const char* who, const void* x, const void* y, int64_t rows, int cols, int dtype) {
  if (rows < 0 || cols <= 0) return emb::fail(-1, "%s: rows=%lld cols=%d", who, (long long)rows, cols);
  if (dtype != 0 && dtype != 1) return emb::fail(-1, "%s: dtype %d (0 = f32, 1 = bf16)", who, dtype);
  const int n = dtype ? 8 : 4;
  if (cols % n) return emb::fail(-1, "%s: cols=%d must be a multiple of %d", who, cols, n);
  if (((uintptr_t)x | (uintptr_t)y) & 15) return emb::fail(-1, "%s: pointers must be 16-byte aligned", who);
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return emb::fail_cuda(who);
  }
  return 0;
}

constexpr int kSmallCols = 256;           // PL = 8: 32 lanes x 8 elements

unsigned grid_for(int64_t rows, int per_sm) {
  const int64_t want = (rows + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)g_sms * per_sm;
  return (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
}

template <typename T, int PL>
int launch_bwd(const char* who, const void* x, const float* scale, const float* bias, const void* gy,
               void* gx, float* gscale, float* gbias, int64_t rows, int cols, int act, float eps,
               cudaStream_t s) {
  auto fn = rmsnorm_act_bwd_kernel<T, PL>;
  static bool attr_set = false;
  const int max_cols = PL <= 8 ? kSmallCols : (PL <= 32 ? 1024 : kMaxCols);
  if (!attr_set) {
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(2 * kWarps * max_cols * sizeof(float))) != cudaSuccess)
      return emb::fail_cuda(who);
    attr_set = true;
  }
  const size_t smem = (size_t)(gbias ? 2 : 1) * kWarps * cols * sizeof(float);
  fn<<<grid_for(rows, PL <= 8 ? 3 : (PL <= 32 ? 4 : 2)), kThreads, smem, s>>>(
      (const T*)x, scale, bias, (const T*)gy, (T*)gx, gscale, gbias, rows, cols, act, eps);
  return 0;
}

}  // namespace

extern "C" int emb_rmsnorm_act_fwd(const void* x, const float* scale, const float* bias, void* y,
                                   int64_t rows, int32_t cols, int32_t dtype, int32_t act, float eps,
                                   void* stream) {
  const char* who = "emb_rmsnorm_act_fwd";
  if (int e = check(who, x, y, rows, cols, dtype)) return e;
  if (rows == 0) return 0;
  const unsigned grid = grid_for(rows, 8);
  cudaStream_t s = (cudaStream_t)stream;
  const int vec = dtype ? 8 : 4, nv = (cols / vec + 31) / 32;
  if (cols >= 32 * vec && nv <= 8) {          // one warp per row, the row in registers
#define EMB_REG(T, NV) rmsnorm_act_fwd_reg_kernel<T, NV><<<grid, kThreads, 0, s>>>( \
      (const T*)x, scale, bias, (T*)y, rows, cols, act, eps)
    if (dtype) {
      if (nv <= 1) EMB_REG(__nv_bfloat16, 1); else if (nv <= 2) EMB_REG(__nv_bfloat16, 2);
      else if (nv <= 4) EMB_REG(__nv_bfloat16, 4); else EMB_REG(__nv_bfloat16, 8);
    } else {
      if (nv <= 1) EMB_REG(float, 1); else if (nv <= 2) EMB_REG(float, 2);
      else if (nv <= 4) EMB_REG(float, 4); else EMB_REG(float, 8);
    }
#undef EMB_REG
  } else if (cols <= kSmallCols) {            // short rows: packed per warp, in registers
    if (dtype)
      rmsnorm_act_fwd_short_kernel<__nv_bfloat16, 1><<<grid, kThreads, 0, s>>>(
          (const __nv_bfloat16*)x, scale, bias, (__nv_bfloat16*)y, rows, cols, act, eps);
    else
      rmsnorm_act_fwd_short_kernel<float, 2><<<grid, kThreads, 0, s>>>(
          (const float*)x, scale, bias, (float*)y, rows, cols, act, eps);
  } else if (dtype)
    rmsnorm_act_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(
        (const __nv_bfloat16*)x, scale, bias, (__nv_bfloat16*)y, rows, cols, act, eps);
  else
    rmsnorm_act_fwd_kernel<float><<<grid, kThreads, 0, s>>>(
        (const float*)x, scale, bias, (float*)y, rows, cols, act, eps);
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}

extern "C" int emb_rmsnorm_act_bwd(const void* x, const float* scale, const float* bias,
                                   const void* gy, void* gx, float* gscale, float* gbias,
                                   int64_t rows, int32_t cols, int32_t dtype, int32_t act, float eps,
                                   void* stream) {
  const char* who = "emb_rmsnorm_act_bwd";
  if (int e = check(who, x, gx, rows, cols, dtype)) return e;
  if ((uintptr_t)gy & 15) return emb::fail(-1, "%s: gy must be 16-byte aligned", who);
  if (cols > kMaxCols) return emb::fail(-1, "%s: cols=%d > %d", who, cols, kMaxCols);
  if ((bias || gbias) && cols > kSmallCols)
    return emb::fail(-1, "%s: a bias needs cols <= %d (got %d)", who, kSmallCols, cols);
  if ((bias == nullptr) != (gbias == nullptr))
    return emb::fail(-1, "%s: bias and gbias must be given together", who);
  if (rows == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int e;
  if (cols <= kSmallCols) {
    e = dtype ? launch_bwd<__nv_bfloat16, 8>(who, x, scale, bias, gy, gx, gscale, gbias, rows, cols, act, eps, s)
              : launch_bwd<float, 8>(who, x, scale, bias, gy, gx, gscale, gbias, rows, cols, act, eps, s);
  } else if (cols <= 1024) {
    e = dtype ? launch_bwd<__nv_bfloat16, 32>(who, x, scale, bias, gy, gx, gscale, gbias, rows, cols, act, eps, s)
              : launch_bwd<float, 32>(who, x, scale, bias, gy, gx, gscale, gbias, rows, cols, act, eps, s);
  } else {
    e = dtype ? launch_bwd<__nv_bfloat16, 64>(who, x, scale, bias, gy, gx, gscale, gbias, rows, cols, act, eps, s)
              : launch_bwd<float, 64>(who, x, scale, bias, gy, gx, gscale, gbias, rows, cols, act, eps, s);
  }
  if (e) return e;
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}
