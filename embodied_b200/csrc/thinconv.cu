// The two THIN 5x5 convolutions of the dreamerv3 encoder / decoder -- the first
// encoder layer (3 image channels in, dreamerv3/rssm.py:233-238) and the decoder's
// image output (3 channels out, rssm.py:349-352; embodied/jax/nets.py:298-323
// Conv2D SAME) -- have one side too thin for an implicit-GEMM convolution: the
// library kernels run them at a few % of any roofline (2.5 / 1.9 ms forward and
// 5.7 ms backward at B*T = 1024, 64x64).  Here they become one skinny tensor-core
// GEMM over all pixels plus one of two HBM-bound rearrangements:
//
//   patches  P[p][(dy*K+dx)*C + c] = x[p + s*(dy-K/2, dx-K/2)][c]   (zero outside; s = +1 / -1)
//   tapsum   y[p][c] = sum_{dy,dx} z[p + (dy-K/2, dx-K/2)][(dy*K+dx)*C + c]
//
//   encoder conv0:  y = patches(x) @ W[(K*K*C), Cout]                    (dW by the GEMM's autograd)
//   decoder imgout: y = tapsum(x @ W[Cin, (K*K*C)])   ;   backward: g_z = patches_{s=-1}(g_y),
//                   then g_x = g_z @ W^T and dW = x^T @ g_z are plain GEMMs again.
//
// `up` = 2 folds the nearest-neighbour x2 up-sampling that precedes the decoder's
// image head (rssm.py:349) into the rearrangement: the GEMM runs on the LOW
// resolution input (4x fewer rows, no up-sampled tensor in HBM), tapsum reads
// z[(p + offset) / 2], and its backward sums the 2x2 block of every low-res pixel.
//
// Rows of P / z are padded to KP columns (a multiple of 8, zero filled by
// `patches`).  dtype 0 = fp32, 1 = bf16 (pure data movement / fp32 accumulation).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// one thread = one 16-byte vector of the output row (VEC elements)
// (H, W) = the OUTPUT grid; x lives on the (H*up, W*up) grid.  One thread = one
// 16-byte vector of the output row; 32-bit index arithmetic, K a compile-time
// constant (divisions by K become multiplies), (tap, channel) advanced incrementally.
template <typename T, int VEC, int K>
__global__ void __launch_bounds__(kThreads)
patches_kernel(const T* __restrict__ x, T* __restrict__ out, uint32_t nvec, int H, int W, int C,
               int KP, int sign, int up) {
  const uint32_t vpr = KP / VEC;
  constexpr int taps = K * K, half = K / 2;
  const int HX = H * up, WX = W * up;
  for (uint32_t o = blockIdx.x * kThreads + threadIdx.x; o < nvec; o += gridDim.x * kThreads) {
    const uint32_t p = o / vpr, v = o - p * vpr;
    const uint32_t row = p / W;
    const int xx = (int)(p - row * W), yy = (int)(row % H);
    const uint32_t img = row / H;
    const T* xi = x + (size_t)img * HX * WX * C;
    int tap = (int)(v * VEC) / C, c = (int)(v * VEC) - tap * C;
    T vals[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float val = 0.f;
      if (tap < taps) {
        const int dy = tap / K - half, dx = tap % K - half;
        if (up == 1) {
          const int sy = yy + sign * dy, sx = xx + sign * dx;
          if (sy >= 0 && sy < HX && sx >= 0 && sx < WX) val = (float)xi[(sy * WX + sx) * C + c];
        } else {
          for (int uy = 0; uy < 2; ++uy) {
            const int sy = yy * 2 + uy + sign * dy;
            if (sy < 0 || sy >= HX) continue;
            for (int ux = 0; ux < 2; ++ux) {
              const int sx = xx * 2 + ux + sign * dx;
              if (sx >= 0 && sx < WX) val += (float)xi[(sy * WX + sx) * C + c];
            }
          }
        }
      }
      vals[i] = T(val);
      if (++c == C) { c = 0; ++tap; }
    }
    *reinterpret_cast<uint4*>(out + (size_t)p * KP + v * VEC) = *reinterpret_cast<const uint4*>(vals);
  }
}

// (H, W) = the OUTPUT grid; z lives on the (H/up, W/up) grid.  One thread = one
// output pixel, all C <= 4 channels (fp32 accumulation).
template <typename T, int K>
__global__ void __launch_bounds__(kThreads)
tapsum_kernel(const T* __restrict__ z, const float* __restrict__ bias, T* __restrict__ y,
              uint32_t pixels, int H, int W, int C, int KP, int up) {
  constexpr int half = K / 2;
  const int HZ = H / up, WZ = W / up;
  for (uint32_t p = blockIdx.x * kThreads + threadIdx.x; p < pixels; p += gridDim.x * kThreads) {
    const uint32_t row = p / W;
    const int xx = (int)(p - row * W), yy = (int)(row % H);
    const uint32_t img = row / H;
    const T* zi = z + (size_t)img * HZ * WZ * KP;
    float acc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = (bias && c < C) ? bias[c] : 0.f;
#pragma unroll
    for (int dy = 0; dy < K; ++dy) {
      const int sy = yy + dy - half;
      if (sy < 0 || sy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < K; ++dx) {
        const int sx = xx + dx - half;
        if (sx < 0 || sx >= W) continue;
        const T* q = zi + (size_t)((sy / up) * WZ + sx / up) * KP + (dy * K + dx) * C;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < C) acc[c] += (float)q[c];
      }
    }
    for (int c = 0; c < C; ++c) y[(size_t)p * C + c] = T(acc[c]);
  }
}

int g_sms = 0;

int prepare(const char* who, int64_t n, int h, int w, int c, int k, int kp, int dtype) {
  if (n < 0 || h < 1 || w < 1 || c < 1 || k < 1 || !(k & 1))
    return emb::fail(-1, "%s: n=%lld h=%d w=%d c=%d k=%d", who, (long long)n, h, w, c, k);
  if (dtype != 0 && dtype != 1) return emb::fail(-1, "%s: dtype %d (0 = f32, 1 = bf16)", who, dtype);
  if (k != 3 && k != 5) return emb::fail(-1, "%s: k=%d (3 or 5)", who, k);
  if (n * (int64_t)h * w * kp >= ((int64_t)1 << 31))
    return emb::fail(-1, "%s: %lld x %d patch elements exceed the 32-bit index range", who,
                     (long long)(n * h * w), kp);
  if (kp % 8 || kp < k * k * c)
    return emb::fail(-1, "%s: kp=%d must be a multiple of 8 and >= k*k*c=%d", who, kp, k * k * c);
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return emb::fail_cuda(who);
  }
  return 0;
}

unsigned grid_for(int64_t items) {
  const int64_t want = (items + kThreads - 1) / kThreads, cap = (int64_t)g_sms * 16;
  return (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

extern "C" int emb_conv_patches_nhwc(const void* x, void* out, int64_t n, int32_t h, int32_t w,
                                     int32_t c, int32_t k, int32_t kp, int32_t sign, int32_t up,
                                     int32_t dtype, void* stream) {
  const char* who = "emb_conv_patches_nhwc";
  if (int e = prepare(who, n, h, w, c, k, kp, dtype)) return e;
  if (sign != 1 && sign != -1) return emb::fail(-1, "%s: sign=%d", who, sign);
  if (up != 1 && up != 2) return emb::fail(-1, "%s: up=%d", who, up);
  if ((uintptr_t)out & 15) return emb::fail(-1, "%s: out must be 16-byte aligned", who);
  const int64_t pixels = n * h * w;
  if (pixels == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype) {
    const int64_t nvec = pixels * (kp / 8);
    auto fn = k == 5 ? patches_kernel<__nv_bfloat16, 8, 5> : patches_kernel<__nv_bfloat16, 8, 3>;
    fn<<<grid_for(nvec), kThreads, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, (uint32_t)nvec,
                                           h, w, c, kp, sign, up);
  } else {
    const int64_t nvec = pixels * (kp / 4);
    auto fn = k == 5 ? patches_kernel<float, 4, 5> : patches_kernel<float, 4, 3>;
    fn<<<grid_for(nvec), kThreads, 0, s>>>((const float*)x, (float*)out, (uint32_t)nvec, h, w, c, kp,
                                           sign, up);
  }
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}

extern "C" int emb_conv_tapsum_nhwc(const void* z, const float* bias, void* y, int64_t n, int32_t h,
                                    int32_t w, int32_t c, int32_t k, int32_t kp, int32_t up,
                                    int32_t dtype, void* stream) {
  const char* who = "emb_conv_tapsum_nhwc";
  if (int e = prepare(who, n, h, w, c, k, kp, dtype)) return e;
  if ((up != 1 && up != 2) || h % up || w % up) return emb::fail(-1, "%s: up=%d h=%d w=%d", who, up, h, w);
  if (c > 4) return emb::fail(-1, "%s: c=%d > 4 output channels", who, c);
  const int64_t total = n * h * w;
  if (total == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype) {
    auto fn = k == 5 ? tapsum_kernel<__nv_bfloat16, 5> : tapsum_kernel<__nv_bfloat16, 3>;
    fn<<<grid_for(total), kThreads, 0, s>>>((const __nv_bfloat16*)z, bias, (__nv_bfloat16*)y,
                                            (uint32_t)total, h, w, c, kp, up);
  } else {
    auto fn = k == 5 ? tapsum_kernel<float, 5> : tapsum_kernel<float, 3>;
    fn<<<grid_for(total), kThreads, 0, s>>>((const float*)z, bias, (float*)y, (uint32_t)total, h, w, c,
                                            kp, up);
  }
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}
