// Shared machinery of the fused RSSM scan kernels (rssm_fwd.cu, rssm_bwd.cu):
// a persistent cooperative grid (one CTA per SM) walks the T steps of
// dreamerv3's RSSM.observe (dreamerv3/rssm.py:61-92,135-159); every step is a
// chain of small-M (16 batch rows) dense layers separated by grid barriers.
// Each layer is HBM-bound weight streaming: the layer's output columns are
// split into n8 tiles over the CTAs, every weight is read exactly once per step
// by exactly one CTA, in the tensor-core B-fragment order it is consumed in.
//
// Two engines, one code path:
//   ENG_BF16  weights packed bf16 in mma.m16n8k16 B-fragment order, activations
//             built as A fragments (shared memory or, for the deter state, a
//             global fragment buffer), mma.sync with fp32 accumulation;
//   ENG_F32   weights fp32 [tile][K][8], plain FFMA -- the parity engine
//             (1e-5 vs the fp32 oracle), not a performance path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rssm {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRows = 16;          // padded batch rows (one m16 tile)
constexpr int kMaxTiles = 24;      // n8 tiles accumulated per pass
constexpr int ENG_F32 = 0;
constexpr int ENG_BF16 = 1;
constexpr int ENG_TMA = 1;          // ABI engine ids: 1 = bf16 + TMA ring, 2 = bf16 register-staged
constexpr int ENG_LEGACY = 2;

__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
// bf16 engine: the result is rounded to bf16 (8 mantissa bits) right away
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------- grid barrier
// Monotonic counter; every CTA arrives once per barrier.  All cross-CTA data is
// read with ld.global.cg (L2), never through the non-coherent L1.
struct GridBarrier {
  unsigned* counter;
  unsigned epoch;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      epoch += gridDim.x;
      __threadfence();
      atomicAdd(counter, 1u);
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      } while ((int)(v - epoch) < 0);
      __threadfence();
    }
    __syncthreads();
  }
};

// -------------------------------------------------------------- fragment maths
// mma.m16n8k16 A fragment: element (row r, k) of a 16 x 16 tile lives in
// lane (r%8)*4 + (k%8)/2, register (r/8) + 2*(k/8), half k%2.
__device__ __forceinline__ uint32_t afrag_index(int r, int k) {
  const int ks = k >> 4, kk = k & 15;
  const int lane = ((r & 7) << 2) + ((kk & 7) >> 1);
  const int reg = (r >> 3) + ((kk >> 3) << 1);
  return ((((uint32_t)ks * 32 + lane) * 4 + reg) << 1) + (kk & 1);   // in bf16 elements
}

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint4& a, const uint2& b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 "
      "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y));
}

__device__ __forceinline__ uint2 ldg_nc_u2(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldcg_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// Sum of squares of each of the 16 rows of y[16][n] (global, fp32) -> rstd[16]
// = rsqrt(mean + eps).  Warp w handles rows w and w + 8.
__device__ __forceinline__ void row_rstd(const float* y, int n, float eps, float* rstd) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kRows; r += kWarps) {
    float s = 0.f;
#pragma unroll 4
    for (int i = lane * 4; i < n; i += 128) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(y + (size_t)r * n + i));
      s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s);
      s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) rstd[r] = rsqrtf(s / (float)n + eps);
  }
}

// ------------------------------------------------------------------ tile GEMM
// out[16][nt*8] (shared, fp32) = A[16][K] @ W[:, tiles j0 .. j0+nt of this CTA)
//
// ENG_BF16: `wblk` is THIS CTA's contiguous weight block [K/16][per][32] uint2
//   (mma B fragments of its `per` n8 tiles, k16-step major); `afrag` the A
//   fragments ([K/16][32] uint4; shared memory, or global when A_GLOBAL).
//   A warp takes KB consecutive k16 steps at a time and keeps KB x TB 8-byte
//   weight loads in flight (16 per lane) before the dependent MMAs.
// ENG_F32: `aval(r, k)` yields A on the fly, `wf32` is the whole layer
//   [tiles][K][8] fp32 and `tile_global0` this call's first tile.  Lane l owns
//   row l/2 and columns (l%2)*4..+3; warp w takes k = w, w+8, ...
template <int KB, int NT, bool A_GLOBAL>
__device__ __forceinline__ void gemm_bf16(
    const uint2* __restrict__ wblk, int per, int j0, int ksteps,
    const uint4* afrag, float* red) {
  // No guards inside the loops: NT is the exact tile count, the k16 steps are
  // consumed in whole rounds of kWarps * KB, the remainder one step at a time.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ncols = NT * 8;
  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const uint2* wl = wblk + (size_t)j0 * 32 + lane;
  const int done = (ksteps / (kWarps * KB)) * (kWarps * KB);
  for (int ks0 = warp * KB; ks0 < done; ks0 += kWarps * KB) {
    uint4 a[KB];
    uint2 b[KB][NT];
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      if (A_GLOBAL) a[kb] = ldcg_u4(afrag + (size_t)(ks0 + kb) * 32 + lane);
      else a[kb] = afrag[(size_t)(ks0 + kb) * 32 + lane];
    }
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
      for (int j = 0; j < NT; ++j)
        b[kb][j] = ldg_nc_u2(wl + ((size_t)(ks0 + kb) * per + j) * 32);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
      for (int j = 0; j < NT; ++j) mma_bf16(acc[j], a[kb], b[kb][j]);
  }
  for (int ks = done + warp; ks < ksteps; ks += kWarps) {
    uint4 a;
    uint2 b[NT];
    if (A_GLOBAL) a = ldcg_u4(afrag + (size_t)ks * 32 + lane);
    else a = afrag[(size_t)ks * 32 + lane];
#pragma unroll
    for (int j = 0; j < NT; ++j) b[j] = ldg_nc_u2(wl + ((size_t)ks * per + j) * 32);
#pragma unroll
    for (int j = 0; j < NT; ++j) mma_bf16(acc[j], a, b[j]);
  }
  // every warp parks its partial sums in its own slab; tile_gemm adds the slabs
  float* slab = red + (size_t)warp * kRows * ncols;
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    float* o = slab + j * 8 + 2 * q;
    *reinterpret_cast<float2*>(o + g * ncols) = make_float2(acc[j][0], acc[j][1]);
    *reinterpret_cast<float2*>(o + (g + 8) * ncols) = make_float2(acc[j][2], acc[j][3]);
  }
}

template <int ENG, bool A_GLOBAL, typename AVal>
__device__ __forceinline__ void tile_gemm(
    const uint2* __restrict__ wblk, int per, int j0,
    const float* __restrict__ wf32, int tile_global0,
    int nt, int K, const uint4* afrag, AVal aval, float* out, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncols = nt * 8;
  if (ENG == ENG_BF16) {
    const int ksteps = K >> 4;
    switch (nt) {
#define EMB_CASE(NT, KB) \
      case NT: gemm_bf16<KB, NT, A_GLOBAL>(wblk, per, j0, ksteps, afrag, red); break;
      EMB_CASE(1, 8) EMB_CASE(2, 8) EMB_CASE(3, 4) EMB_CASE(4, 4) EMB_CASE(5, 2) EMB_CASE(6, 2)
      EMB_CASE(7, 2) EMB_CASE(8, 2) EMB_CASE(9, 1) EMB_CASE(10, 1) EMB_CASE(11, 1) EMB_CASE(12, 1)
      EMB_CASE(13, 1) EMB_CASE(14, 1) EMB_CASE(15, 1) EMB_CASE(16, 1) EMB_CASE(17, 1) EMB_CASE(18, 1)
      EMB_CASE(19, 1) EMB_CASE(20, 1) EMB_CASE(21, 1) EMB_CASE(22, 1) EMB_CASE(23, 1) EMB_CASE(24, 1)
#undef EMB_CASE
      default: break;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kRows * ncols; i += kThreads) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += red[(size_t)w * kRows * ncols + i];
      out[i] = s;
    }
  } else {
    for (int i = threadIdx.x; i < kRows * ncols; i += kThreads) out[i] = 0.f;
    __syncthreads();
    const float* wp = wf32;
    const int r = lane >> 1, h = lane & 1;
    float acc[kMaxTiles][4];
#pragma unroll
    for (int j = 0; j < kMaxTiles; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    for (int k = warp; k < K; k += kWarps) {
      const float x = aval(r, k);
#pragma unroll
      for (int j = 0; j < kMaxTiles; ++j) {
        if (j < nt) {
          const float4 wv = ldg_nc_f4(reinterpret_cast<const float4*>(
              wp + ((size_t)(tile_global0 + j) * K + k) * 8 + h * 4));
          acc[j][0] = fmaf(x, wv.x, acc[j][0]);
          acc[j][1] = fmaf(x, wv.y, acc[j][1]);
          acc[j][2] = fmaf(x, wv.z, acc[j][2]);
          acc[j][3] = fmaf(x, wv.w, acc[j][3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kMaxTiles; ++j) {
      if (j < nt) {
        float* o = out + r * ncols + j * 8 + h * 4;
        atomicAdd(o, acc[j][0]); atomicAdd(o + 1, acc[j][1]);
        atomicAdd(o + 2, acc[j][2]); atomicAdd(o + 3, acc[j][3]);
      }
    }
  }
  __syncthreads();
}

// Ask the memory system to pull this CTA's NEXT weight block (contiguous `bytes`)
// from HBM into L2 while the current phase, its barrier and the next prologue run:
// the block is then streamed at L2 latency, and HBM stays busy across barriers.
__device__ __forceinline__ void prefetch_l2(const void* p, size_t bytes) {
  const char* c = reinterpret_cast<const char*>(p);
  for (size_t off = (size_t)threadIdx.x * 128; off < bytes; off += (size_t)kThreads * 128)
    asm volatile("prefetch.global.L2 [%0];" :: "l"(c + off));
}

// Copy `n16` A fragments (uint4) from global (L2) to shared memory.
__device__ __forceinline__ void copy_frags(uint4* dst, const uint4* src, int n16) {
#pragma unroll 4
  for (int i = threadIdx.x; i < n16; i += kThreads) dst[i] = ldcg_u4(src + i);
}

// A-fragment builders (bf16 engine).  `src` is fp32 [16][n] in global memory
// written by other CTAs (read through L2); f(r, k, v) transforms element (r, k).
// Four consecutive k per thread: one 16-byte load, two 4-byte fragment stores.
template <typename F>
__device__ __forceinline__ void build_part(__nv_bfloat16* afrag, int koff, const float* src,
                                           int n, int ld, F f) {
  const int n4 = n >> 2;
#pragma unroll 4
  for (int i = threadIdx.x; i < kRows * n4; i += kThreads) {
    const int r = i / n4, k = (i - r * n4) << 2;
    const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (size_t)r * ld + k));
    *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, koff + k)) =
        __floats2bfloat162_rn(f(r, k, v.x), f(r, k + 1, v.y));
    *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, koff + k + 2)) =
        __floats2bfloat162_rn(f(r, k + 2, v.z), f(r, k + 3, v.w));
  }
}

// Build A fragments in shared memory from aval(r, k), k in [0, K).
template <typename AVal>
__device__ __forceinline__ void build_afrag(__nv_bfloat16* afrag, int K, AVal aval) {
  for (int i = threadIdx.x; i < kRows * (K >> 1); i += kThreads) {
    const int r = i / (K >> 1), k = (i - r * (K >> 1)) << 1;
    const __nv_bfloat162 v = __floats2bfloat162_rn(aval(r, k), aval(r, k + 1));
    *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, k)) = v;
  }
  __syncthreads();
}

struct NoVal {
  __device__ __forceinline__ float operator()(int, int) const { return 0.f; }
};

// Block-diagonal layers: CTA c works inside ONE group (ncta / groups CTAs per
// group), on a run of `per` units of that group.  Same formulas as scan.py
// tile_assignment().  Returns the global unit range [u0, u1).
struct GroupSplit {
  int per, u0, u1;
};
__device__ __forceinline__ GroupSplit group_split(int units_per_group, int groups) {
  const int cpg = max(1, (int)gridDim.x / groups);
  GroupSplit s;
  s.per = (units_per_group + cpg - 1) / cpg;
  const int g = blockIdx.x / cpg, j = blockIdx.x - g * cpg;
  if (g >= groups) { s.u0 = s.u1 = 0; return s; }
  s.u0 = g * units_per_group + min(units_per_group, j * s.per);
  s.u1 = g * units_per_group + min(units_per_group, (j + 1) * s.per);
  return s;
}

// CTA c's share [begin, end) of `total` work units.
__device__ __forceinline__ void cta_range(int total, int& begin, int& end) {
  const int per = (total + gridDim.x - 1) / gridDim.x;
  begin = min(total, (int)blockIdx.x * per);
  end = min(total, begin + per);
}

}  // namespace rssm
