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

// Both kernels work a group of R output rows of one image at a time: the input rows
// the group touches are staged in shared memory with a zero halo (coalesced reads,
// no bounds checks afterwards, each input row read ~once instead of K times), and
// the output rows are written as consecutive vectors.  Everything that depends only
// on the column of the output row -- which tile cells a thread's vector gathers --
// is computed once per kernel: these kernels are bound by instruction issue, not by
// HBM, unless the per-element work is a handful of instructions.

// (H, W) = the OUTPUT grid; x lives on the (H*UP, W*UP) grid.  Tile = the
// R*UP + K - 1 input rows [yy0*UP - K/2, ..] x (W*UP + K - 1) pixels x C, followed by
// UP rows of zeros that the padding columns (>= K*K*C) read.  Thread t owns vector
// column t % (KP/VEC) of the pixels t / (KP/VEC) + i * (threads / (KP/VEC)).
template <typename T, int VEC, int K, int UP>
__global__ void __launch_bounds__(kThreads)
patches_kernel(const T* __restrict__ x, T* __restrict__ out, int groups_total, int H, int W, int C,
               int KP, int sign, int R) {
  extern __shared__ __align__(16) unsigned char tile_raw[];
  T* tile = reinterpret_cast<T*>(tile_raw);
  constexpr int taps = K * K, half = K / 2;
  const int HX = H * UP, WX = W * UP;
  const int TWC = (WX + 2 * half) * C;
  const int TR = R * UP + K - 1;           // real rows of the tile
  const int vpr = KP / VEC, ppp = kThreads / vpr;
  const int v = threadIdx.x % vpr, px0 = threadIdx.x / vpr;
  const int gpi = (H + R - 1) / R;         // groups per image
  int offs[VEC];
  {
    int tap = (v * VEC) / C, c = v * VEC - tap * C;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      if (tap < taps) {
        const int dyi = tap / K, dxi = tap - dyi * K;
        offs[i] = (half + sign * (dyi - half)) * TWC + (half + sign * (dxi - half)) * C + c;
      } else {
        offs[i] = -1;                      // padding column
      }
      if (++c == C) { c = 0; ++tap; }
    }
  }
  for (int j = threadIdx.x; j < UP * TWC; j += kThreads) tile[TR * TWC + j] = T(0.f);
  for (int g = blockIdx.x; g < groups_total; g += gridDim.x) {
    const int img = g / gpi, yy0 = (g - img * gpi) * R;
    const int nrows = min(R, H - yy0);
    const T* xi = x + (size_t)img * HX * WX * C;
    const int y0 = yy0 * UP - half;
    __syncthreads();                       // the previous group's readers are done
    for (int tr = 0; tr < nrows * UP + K - 1; ++tr) {
      const int sy = y0 + tr;
      const bool row_ok = sy >= 0 && sy < HX;
      const T* src = xi + (size_t)(row_ok ? sy : 0) * WX * C;
      for (int j = threadIdx.x; j < TWC; j += kThreads) {
        const int jj = j - half * C;
        tile[tr * TWC + j] = (row_ok && jj >= 0 && jj < WX * C) ? src[jj] : T(0.f);
      }
    }
    __syncthreads();
    if (px0 < ppp) {
      T* orow = out + ((size_t)img * H + yy0) * W * KP + v * VEC;
      int ry = 0, xx = px0;
      while (xx >= W) { xx -= W; ++ry; }
      while (ry < nrows) {
        const T* q = tile + ry * UP * TWC + xx * UP * C;
        const T* qz = tile + TR * TWC + xx * UP * C;      // same column of the zero rows
        T vals[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const T* qi = offs[i] >= 0 ? q + offs[i] : qz;
          if (UP == 1) {
            vals[i] = qi[0];
          } else {                         // the 2x2 block of this low-resolution pixel
            float val = (float)qi[0];
            val += (float)qi[C];
            val += (float)qi[TWC];
            val += (float)qi[TWC + C];
            vals[i] = T(val);
          }
        }
        *reinterpret_cast<uint4*>(orow + ((size_t)ry * W + xx) * KP) = *reinterpret_cast<const uint4*>(vals);
        xx += ppp;
        while (xx >= W) { xx -= W; ++ry; }
      }
    }
  }
}

// (H, W) = the OUTPUT grid; z lives on the (H/UP, W/UP) grid.  Tile = the z rows the
// group's taps reach x (W/UP pixels + halo) x KP columns, copied as 4-element vectors;
// pixels are KP + 4 elements apart in the tile so that the threads of a warp (one
// output pixel each, same tap) read different banks.  One thread = one output pixel,
// all C <= 4 channels, taps summed (dy, dx)-major in fp32.
template <typename T> struct Quad;
template <> struct Quad<float> { using type = float4; };
template <> struct Quad<__nv_bfloat16> { using type = uint2; };

template <typename T, int K, int UP>
__global__ void __launch_bounds__(kThreads)
tapsum_kernel(const T* __restrict__ z, const float* __restrict__ bias, T* __restrict__ y,
              int groups_total, int H, int W, int C, int KP, int R) {
  using Q = typename Quad<T>::type;
  extern __shared__ __align__(16) unsigned char tile_raw[];
  T* tile = reinterpret_cast<T*>(tile_raw);
  constexpr int half = K / 2;
  constexpr int hal = (half + UP - 1) / UP;  // z pixels / rows reached past either end
  const int HZ = H / UP, WZ = W / UP;
  const int TS = KP + 4, KV = KP / 4;        // tile pixel stride (elements), quads per pixel
  const int TW = WZ + 2 * hal, TWS = TW * TS;
  const int gpi = (H + R - 1) / R;
  // this thread's walk over a tile row in steps of kThreads quads: (pixel, quad) pairs
  const int tx0 = threadIdx.x / KV, e0 = threadIdx.x - tx0 * KV;
  const int dtx = kThreads / KV, de = kThreads - dtx * KV;
  float b[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) b[c] = (bias && c < C) ? bias[c] : 0.f;
  for (int g = blockIdx.x; g < groups_total; g += gridDim.x) {
    const int img = g / gpi, yy0 = (g - img * gpi) * R;
    const int nrows = min(R, H - yy0);
    const T* zi = z + (size_t)img * HZ * WZ * KP;
    const int zr0 = (yy0 - half + UP * hal) / UP - hal;          // first z row (may be < 0)
    const int nz = (yy0 + nrows - 1 + half) / UP - zr0 + 1;
    __syncthreads();
    for (int zr = 0; zr < nz; ++zr) {
      const int zy = zr0 + zr;
      const bool row_ok = zy >= 0 && zy < HZ;
      const T* src = zi + (size_t)(row_ok ? zy : 0) * WZ * KP;
      int tx = tx0, e = e0;
      for (int j = threadIdx.x; j < TW * KV; j += kThreads) {
        const int zx = tx - hal;
        Q val = {};
        if (row_ok && zx >= 0 && zx < WZ) val = *reinterpret_cast<const Q*>(src + zx * KP + e * 4);
        *reinterpret_cast<Q*>(tile + zr * TWS + tx * TS + e * 4) = val;
        tx += dtx; e += de;
        if (e >= KV) { e -= KV; ++tx; }
      }
    }
    __syncthreads();
    int ry = 0, xx = threadIdx.x;
    while (xx >= W) { xx -= W; ++ry; }
    while (ry < nrows) {
      const int yy = yy0 + ry;
      float acc[4] = {b[0], b[1], b[2], b[3]};
#pragma unroll
      for (int dyi = 0; dyi < K; ++dyi) {
        const int sy = yy + dyi - half;
        if (sy < 0 || sy >= H) continue;
        const T* q = tile + (sy / UP - zr0) * TWS + dyi * K * C;
#pragma unroll
        for (int dxi = 0; dxi < K; ++dxi) {
          const T* qq = q + ((xx + dxi - half + UP * hal) / UP) * TS + dxi * C;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < C) acc[c] += (float)qq[c];
        }
      }
      T* o = y + (((size_t)img * H + yy) * W + xx) * C;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < C) o[c] = T(acc[c]);
      xx += kThreads;
      while (xx >= W) { xx -= W; ++ry; }
    }
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

constexpr size_t kMaxTile = 48 * 1024;     // shared-memory tile of one row group
constexpr int kMaxGroupRows = 8;

unsigned grid_for(int64_t groups) {
  const int64_t cap = (int64_t)g_sms * 8;
  return (unsigned)(groups < cap ? (groups < 1 ? 1 : groups) : cap);
}

size_t patches_tile(int r, int w, int c, int k, int up, int dtype) {
  return (size_t)(r * up + k - 1 + up) * (w * up + k - 1) * c * (dtype ? 2 : 4);
}

size_t tapsum_tile(int r, int w, int kp, int k, int up, int dtype) {
  const int hal = (k / 2 + up - 1) / up;
  const int nz = (r + k - 1 + up - 1) / up + 1;      // z rows a group can touch (upper bound)
  return (size_t)nz * (w / up + 2 * hal) * (kp + 4) * (dtype ? 2 : 4);
}

// rows per group: as many as the tile budget allows, at most kMaxGroupRows; 0 = none fit
template <typename F>
int group_rows(int h, F tile_bytes) {
  int r = h < kMaxGroupRows ? h : kMaxGroupRows;
  while (r > 0 && tile_bytes(r) > kMaxTile) r >>= 1;
  return r;
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
  if (n * h == 0) return 0;
  if (kp / (dtype ? 8 : 4) > kThreads) return emb::fail(-1, "%s: kp=%d too wide", who, kp);
  const int r = group_rows(h, [&](int rr) { return patches_tile(rr, w, c, k, up, dtype); });
  if (r == 0)
    return emb::fail(-1, "%s: one %d-pixel row of %d channels does not fit a %zu-byte tile", who,
                     w * up, c, kMaxTile);
  const size_t tile = patches_tile(r, w, c, k, up, dtype);
  const int64_t rows = n * ((h + r - 1) / r);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype) {
    using B = __nv_bfloat16;
    auto fn = k == 5 ? (up == 2 ? patches_kernel<B, 8, 5, 2> : patches_kernel<B, 8, 5, 1>)
                     : (up == 2 ? patches_kernel<B, 8, 3, 2> : patches_kernel<B, 8, 3, 1>);
    fn<<<grid_for(rows), kThreads, tile, s>>>((const B*)x, (B*)out, (int)rows, h, w, c, kp, sign, r);
  } else {
    auto fn = k == 5 ? (up == 2 ? patches_kernel<float, 4, 5, 2> : patches_kernel<float, 4, 5, 1>)
                     : (up == 2 ? patches_kernel<float, 4, 3, 2> : patches_kernel<float, 4, 3, 1>);
    fn<<<grid_for(rows), kThreads, tile, s>>>((const float*)x, (float*)out, (int)rows, h, w, c, kp, sign, r);
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
  if (n * h == 0) return 0;
  const int r = group_rows(h, [&](int rr) { return tapsum_tile(rr, w, kp, k, up, dtype); });
  if (r == 0)
    return emb::fail(-1, "%s: one %d-pixel row of %d channels does not fit a %zu-byte tile", who, w,
                     c, kMaxTile);
  const size_t tile = tapsum_tile(r, w, kp, k, up, dtype);
  const int64_t rows = n * ((h + r - 1) / r);
  cudaStream_t s = (cudaStream_t)stream;
  if (kp / 4 > kThreads) return emb::fail(-1, "%s: kp=%d too wide", who, kp);
  if ((uintptr_t)z & 15) return emb::fail(-1, "%s: z must be 16-byte aligned", who);
  if (dtype) {
    using B = __nv_bfloat16;
    auto fn = k == 5 ? (up == 2 ? tapsum_kernel<B, 5, 2> : tapsum_kernel<B, 5, 1>)
                     : (up == 2 ? tapsum_kernel<B, 3, 2> : tapsum_kernel<B, 3, 1>);
    fn<<<grid_for(rows), kThreads, tile, s>>>((const B*)z, bias, (B*)y, (int)rows, h, w, c, kp, r);
  } else {
    auto fn = k == 5 ? (up == 2 ? tapsum_kernel<float, 5, 2> : tapsum_kernel<float, 5, 1>)
                     : (up == 2 ? tapsum_kernel<float, 3, 2> : tapsum_kernel<float, 3, 1>);
    fn<<<grid_for(rows), kThreads, tile, s>>>((const float*)z, bias, (float*)y, (int)rows, h, w, c, kp, r);
  }
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}
