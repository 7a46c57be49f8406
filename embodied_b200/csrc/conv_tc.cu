// emb_conv5x5_nhwc_tc: the 5x5 SAME convolutions of the dreamerv3 encoder / decoder
// (dreamerv3/rssm.py:233-240, 336-352; embodied/jax/nets.py:298-323 Conv2D) as an
// implicit GEMM on the 5th-generation tensor cores.
//
//   out[p][co] = sum_{tap=(ky,kx)} sum_ci  in[p + (ky-2, kx-2)][ci] * w[tap][co][ci]
//
// One persistent CTA per SM walks output tiles of whole image rows (contiguous in NHWC memory):
// 128 pixels where the width divides 128, otherwise the largest whole-row / whole-image tile below
// 128 (96-wide maps: 96 pixels) with the remaining MMA rows unused.  Per (tap, 64-channel block):
//   A = the 128 x 64 window of the input shifted by the tap, fetched by ONE 4-D TMA
//       tile copy (cp.async.bulk.tensor, 128-byte swizzle); the SAME padding is the
//       TMA's out-of-bounds zero fill -- no halo handling in the kernel;
//   B = the tap's [Cout][64] weight slice (K-major, packed once per optimiser step);
//   D += A @ B^T by tcgen05.mma (M = 128, N = Cout <= 256, K = 16 per instruction),
//       fp32 accumulators in TENSOR MEMORY, two accumulator buffers of 256 columns
//       so that the epilogue of tile i overlaps the MMAs of tile i + 1.
// Warp roles: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane; owns the
// TMEM allocation), warps 2..5 = epilogue (tcgen05.ld 32 lanes x 32 columns -> bf16 ->
// 64-byte stores per thread).  smem ring: 4 stages x (16 KiB A + 32 KiB B).
// The same kernel is the data-gradient of the convolution (flipped taps, channels
// swapped -- a different weight packing, see dreamerv3/ops.py).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {

constexpr int kStages = 4;
constexpr int kABytes = 128 * 128;          // 128 pixels x 64 bf16
constexpr int kBBytes = 256 * 128;          // <= 256 output channels x 64 bf16
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;

struct ConvShape {
  int tiles;           // output tiles of `tpix` pixels
  int tpix;            // pixels per tile: whole rows (or whole images), at most 128; the MMA always runs
                       // M = 128 rows -- rows >= tpix of the shared tile hold stale data and land in
                       // accumulator lanes nobody reads (96-wide maps: one row = 96 pixels per tile)
  int hw;              // H * W
  int w;               // W
  int kblocks;         // Cin / 64
  int cout;
  int taps;            // k * k
  int ksize;           // k
  int msub;            // 128-pixel sub-tiles per tile that share one weight tile (2 when cout <= 128)
  int groups;          // 1, or 4: K also runs over the 2x2 phases of an input stored on the doubled grid
  int cin;             // channels per phase
  int out_up;          // 1, or 2: the output is phase (out_py, out_px) of a tensor on the doubled grid
  int out_py, out_px;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5}], [%6];"
      :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
// Activations are always described as 5-D tensors {channels', x, phase, y, image}: a plain NHWC tensor
// has a phase axis of length 1; the four 2x2 phases of a tensor stored on the doubled grid
// (n, 2h, 2w, c) are {channels' = 2c with px*c as the channel offset, x (stride 2c), py, y, n}.
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            int c3, int c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];"
      :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// tcgen05 shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of
// 128 bytes (64 bf16 of K), 8-row swizzle atoms 1024 bytes apart (cute UMMA::SmemDescriptor:
// start >> 4 [0,14) | LBO >> 4 [16,30) | SBO >> 4 [32,46) | version 1 [46,48) | layout [61,64)).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                      // leading byte offset: unused with swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                      // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute UMMA::InstrDescriptor): D = f32, A = B = bf16, both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on `bar` once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns of tensor memory -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(uint32_t lo, uint32_t hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(lo), __uint_as_float(hi));
  return *reinterpret_cast<const uint32_t*>(&v);
}

template <int MSUB>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w,
               __nv_bfloat16* __restrict__ out, const float* __restrict__ bias, const ConvShape s) {
  extern __shared__ unsigned char smem_raw[];
  // 128-byte swizzle atoms are anchored on 1024-byte boundaries of the shared window
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;                 // [kStages]  TMA -> MMA
  uint64_t* empty = bars + kStages;      // [kStages]  MMA -> TMA
  uint64_t* tfull = bars + 2 * kStages;  // [2]        MMA -> epilogue (accumulator ready)
  uint64_t* tempty = tfull + 2;          // [2]        epilogue -> MMA (accumulator drained)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_in) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_w) : "memory");
  }
  if (warp == 1) {                        // one warp allocates (and later frees) tensor memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int iters = s.groups * s.taps * s.kblocks;
  // stage = [msub x 16 KiB of A | weight tile]: 48 KiB either way (cout <= 128 with two sub-tiles)
  constexpr uint32_t a_bytes = (uint32_t)MSUB * kABytes;
  const uint32_t stage_tx = (uint32_t)MSUB * (uint32_t)s.tpix * 128u + (uint32_t)s.cout * 128u;
  const int half = s.ksize >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < s.tiles; tile += gridDim.x) {
        int n0[MSUB], h0[MSUB];
#pragma unroll
        for (int sub = 0; sub < MSUB; ++sub) {
          const int pix0 = (tile * MSUB + sub) * s.tpix;
          n0[sub] = pix0 / s.hw;
          h0[sub] = (pix0 - n0[sub] * s.hw) / s.w;
        }
        for (int g = 0; g < s.groups; ++g) {
          const int gpy = g >> 1, gch = (g & 1) * s.cin;
          int ky = 0, kx = 0;
          for (int tap = 0; tap < s.taps; ++tap) {
            for (int kb = 0; kb < s.kblocks; ++kb) {
              mbar_wait(&empty[stage], phase ^ 1u);
              mbar_expect_tx(&full[stage], stage_tx);
              unsigned char* base = smem + stage * kStageBytes;
#pragma unroll
              for (int sub = 0; sub < MSUB; ++sub)
                tma_load_5d(base + sub * kABytes, &map_in, gch + kb * 64, kx - half, gpy, h0[sub] + ky - half,
                            n0[sub], &full[stage]);
              tma_load_3d(base + a_bytes, &map_w, kb * 64, 0, g * s.taps + tap, &full[stage]);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            if (++kx == s.ksize) { kx = 0; ++ky; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, s.cout);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t accphase = 0;
      for (int tile = blockIdx.x; tile < s.tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], accphase ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)acc * 256u;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + stage * kStageBytes);
          const uint64_t bd = umma_desc_sw128(base + a_bytes);
#pragma unroll
          for (int sub = 0; sub < MSUB; ++sub) {
            const uint64_t ad = umma_desc_sw128(base + sub * kABytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 4 x (K = 16): +32 bytes inside the swizzle atom
              umma_bf16(d + (uint32_t)sub * 128u, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc,
                        (it | k) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);      // the stage is free once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull[acc]);          // accumulator complete
        acc ^= 1;
        if (acc == 0) accphase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;                // a warp reads TMEM lanes [32 * (warp % 4), +32)
    int acc = 0;
    uint32_t accphase = 0;
    for (int tile = blockIdx.x; tile < s.tiles; tile += gridDim.x) {
      mbar_wait(&tfull[acc], accphase);
      tc_fence_after();
#pragma unroll
      for (int sub = 0; sub < MSUB; ++sub) {
        const int m = q * 32 + lane;
        if (q * 32 >= s.tpix) continue;      // a whole warp past a partially filled tile (warp-uniform)
        const bool valid = m < s.tpix;       // the TMEM loads below are warp-collective: only stores are guarded
        size_t opix = (size_t)(tile * MSUB + sub) * s.tpix + (valid ? m : 0);
        if (s.out_up == 2) {                 // pixel (n, y, x) -> (n, 2y + py, 2x + px) of the doubled grid
          const int pn = (int)(opix / s.hw), pr = (int)(opix - (size_t)pn * s.hw);
          const int py = pr / s.w, px = pr - py * s.w;
          opix = ((size_t)pn * 2 * (s.hw / s.w) + 2 * py + s.out_py) * (2 * s.w) + 2 * px + s.out_px;
        }
        __nv_bfloat16* orow = out + opix * s.cout;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u + (uint32_t)sub * 128u;
        for (int c0 = 0; c0 < s.cout; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
          if (bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(bias + c0 + j));
          }
          uint4* dst = reinterpret_cast<uint4*>(orow + c0);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                  pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) accphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 :: "r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---- weight gradient ------------------------------------------------------------
// dW[tap][ci][co] = sum over pixels p of in[p + shift(tap)][ci] * gy[p][co]: per tap a GEMM whose
// reduction axis is the PIXEL axis.  Both operands are read in their natural NHWC layout by the same
// TMA boxes as above ([64 pixels][64 channels], 128-byte swizzle) and enter the tensor core as
// MN-major operands (cute Layout_MN_SW128_Atom: the 64-channel slabs LBO bytes apart, 8-pixel
// groups SBO = 1024 bytes apart).  D = [M = 128 or 256 channels of one tensor][N <= 256 channels of
// the other], fp32 in tensor memory for the CTA's whole pixel range; red.global.add at the end.
// CTA = (tap, split): the 25 taps of a split walk the same pixels at the same time, so every
// activation tile comes from HBM once and is then served by L2 to the other 24 taps.
constexpr int kWSlab = 64 * 128;                 // one TMA box: 64 pixels x 64 channels bf16
constexpr int kWRingSlabs = 24;                  // 192 KiB ring: 24 / SLABS stages of SLABS = (M + N) / 64 slabs
// (Measured, r02: an L2 look-ahead by the centre tap of every split -- cp.async.bulk.prefetch.tensor
//  12 chunks ahead -- and a ring of 4-6 smaller stages made every layer 1.7-2x SLOWER; the ring is
//  bound by bytes in flight per SM, not by HBM latency of a leader.)

struct WgradShape {
  int chunks;          // pixel chunks in all
  int cpix;            // pixels per chunk = bw * ht * nt <= 64 (64 where the width divides 64); the K loop
                       // always covers 64 rows: rows >= cpix of a slab are never written and stay zero
  int bw;              // box width: the image width, or a divisor of it when w > 64 (chunks then walk x too)
  int splits;          // CTAs per tap
  int hw, w;           // H*W, W
  int m, n;            // channels on the M / N side (m in {128, 256})
  int ksize;
  int m_shifted;       // 1: the M-side tensor is the convolution input (shifted per tap)
  int h, ht, nt;       // image height; rows / images per 64-pixel chunk
  int gy_py, gy_ch;    // phase row and channel offset (px * cout) of gy on a doubled grid, else 0
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // between 64-element slabs along M / N
  d |= (uint64_t)(1024 >> 4) << 32;                   // between 8-row groups along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(dst), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "f"(__uint_as_float(c)),
                  "f"(__uint_as_float(d)) : "memory");
}

template <int SLABS>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_m, const __grid_constant__ CUtensorMap map_n,
                     float* __restrict__ dw, const WgradShape s) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kWStages = kWRingSlabs / SLABS;
  constexpr int kWStageBytes = SLABS * kWSlab;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWRingSlabs * kWSlab);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWStages;
  uint64_t* done = bars + 2 * kWStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_m) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_n) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (s.cpix < 64) {                       // partially filled chunks: the rows TMA never writes must be zero
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < kWRingSlabs * kWSlab / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int taps = s.ksize * s.ksize, half = s.ksize >> 1;
  const int tap = blockIdx.x % taps, split = blockIdx.x / taps;
  const int ky = tap / s.ksize, kx = tap - ky * s.ksize;
  const int c_begin = (int)((long long)s.chunks * split / s.splits);
  const int c_end = (int)((long long)s.chunks * (split + 1) / s.splits);
  const int mslabs = s.m / 64, nslabs = s.n / 64, mtiles = s.m / 128;
  const uint32_t stage_tx = (uint32_t)(mslabs + nslabs) * (uint32_t)s.cpix * 128u;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // (no divisions inside the loop: the producer's issue latency is on the critical path)
      const long long p0 = (long long)c_begin * s.cpix;
      int n0 = (int)(p0 / s.hw), h0 = (int)((p0 - (long long)n0 * s.hw) / s.w);
      int x0 = s.bw < s.w ? (int)(p0 - (long long)n0 * s.hw - (long long)h0 * s.w) : 0;
      const int dmh = s.m_shifted ? ky - half : 0, dmw = s.m_shifted ? kx - half : 0;
      const int dnh = s.m_shifted ? 0 : ky - half, dnw = s.m_shifted ? 0 : kx - half;
      // the unshifted side is gy: possibly one phase of a tensor on the doubled grid
      const int mpy = s.m_shifted ? 0 : s.gy_py, mch = s.m_shifted ? 0 : s.gy_ch;
      const int npy = s.m_shifted ? s.gy_py : 0, nch = s.m_shifted ? s.gy_ch : 0;
      if (s.bw == s.w) {
        // whole-row boxes (every width up to 64): the original loop, nothing but two adds and a compare per
        // chunk -- this lane's issue latency is on the critical path (measured: the x walk below costs 5-9 %)
        for (int c = c_begin; c < c_end; ++c) {
          const int mh = h0 + dmh, nh = h0 + dnh;
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_expect_tx(&full[stage], stage_tx);
          unsigned char* base = smem + stage * kWStageBytes;
          for (int j = 0; j < mslabs; ++j)
            tma_load_5d(base + j * kWSlab, &map_m, mch + j * 64, dmw, mpy, mh, n0, &full[stage]);
          for (int j = 0; j < nslabs; ++j)
            tma_load_5d(base + (mslabs + j) * kWSlab, &map_n, nch + j * 64, dnw, npy, nh, n0, &full[stage]);
          if (++stage == kWStages) { stage = 0; phase ^= 1u; }
          h0 += s.ht;
          if (h0 >= s.h) { h0 = 0; n0 += s.nt; }
        }
      } else {
        // rows wider than 64 pixels are walked in pieces of bw pixels
        for (int c = c_begin; c < c_end; ++c) {
          const int mh = h0 + dmh, nh = h0 + dnh, mw = x0 + dmw, nw = x0 + dnw;
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_expect_tx(&full[stage], stage_tx);
          unsigned char* base = smem + stage * kWStageBytes;
          for (int j = 0; j < mslabs; ++j)
            tma_load_5d(base + j * kWSlab, &map_m, mch + j * 64, mw, mpy, mh, n0, &full[stage]);
          for (int j = 0; j < nslabs; ++j)
            tma_load_5d(base + (mslabs + j) * kWSlab, &map_n, nch + j * 64, nw, npy, nh, n0, &full[stage]);
          if (++stage == kWStages) { stage = 0; phase ^= 1u; }
          x0 += s.bw;
          if (x0 >= s.w) {
            x0 = 0;
            if (++h0 >= s.h) { h0 = 0; ++n0; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, s.n) | (1u << 15) | (1u << 16);   // both MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + stage * kWStageBytes);
        for (int mt = 0; mt < mtiles; ++mt) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {     // K = 16 pixels = two 8-row groups = 2048 bytes
            const uint64_t ad = umma_desc_mn_sw128(base + mt * 2 * kWSlab + k * 2048, kWSlab);
            const uint64_t bd = umma_desc_mn_sw128(base + mslabs * kWSlab + k * 2048, kWSlab);
            umma_bf16(tmem_base + (uint32_t)mt * 256u, ad, bd, idesc, (c > c_begin || k) ? 1u : 0u);
          }
        }
        umma_commit(&empty[stage]);
        if (++stage == kWStages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(done);
    }
  } else if (c_begin < c_end) {
    mbar_wait(done, 0);
    tc_fence_after();
    const int q = warp & 3;
    for (int mt = 0; mt < mtiles; ++mt) {
      const int m = mt * 128 + q * 32 + lane;
      float* drow = dw + ((size_t)tap * s.m + m) * s.n;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)mt * 256u;
      for (int c0 = 0; c0 < s.n; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) red_add_v4(drow + c0 + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 :: "r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---- weight gradient, two taps per CTA (M side of 128 channels) ----------------------
// With M = 128 a tap's D tile is 128 x N: every operand byte fetched from L2 feeds only 128 or N
// multiply-adds, and the 32 x 32 layers (1 M pixels) run into the L2 -> shared-memory rate.  A CTA
// that owns TWO taps keeps two accumulators (2 x 256 TMEM columns) and fetches the UNSHIFTED
// operand (gy) once for both: 7-8 slabs per chunk instead of 10 for the same multiply-adds.
// Stage = [shared slabs | tap 0 slabs | tap 1 slabs]; the last group of an odd tap count has one tap.
template <int SLABS>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_pair_kernel(const __grid_constant__ CUtensorMap map_m, const __grid_constant__ CUtensorMap map_n,
                       float* __restrict__ dw, const WgradShape s) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kWStages = kWRingSlabs / SLABS;
  constexpr int kWStageBytes = SLABS * kWSlab;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWRingSlabs * kWSlab);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWStages;
  uint64_t* done = bars + 2 * kWStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_m) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_n) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (s.cpix < 64) {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < kWRingSlabs * kWSlab / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int taps = s.ksize * s.ksize, half = s.ksize >> 1;
  const int ngroups = (taps + 1) >> 1;
  const int group = blockIdx.x % ngroups, split = blockIdx.x / ngroups;
  const int tap0 = group * 2, ng = taps - tap0 >= 2 ? 2 : 1;
  const int c_begin = (int)((long long)s.chunks * split / s.splits);
  const int c_end = (int)((long long)s.chunks * (split + 1) / s.splits);
  const int mslabs = s.m / 64, nslabs = s.n / 64;            // m == 128: mslabs == 2
  // the shifted tensor (the convolution input) is fetched per tap, the other one (gy) once
  const int shared_slabs = s.m_shifted ? nslabs : mslabs, tap_slabs = s.m_shifted ? mslabs : nslabs;
  const uint32_t stage_tx = (uint32_t)(shared_slabs + ng * tap_slabs) * (uint32_t)s.cpix * 128u;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const long long p0 = (long long)c_begin * s.cpix;
      int n0 = (int)(p0 / s.hw), h0 = (int)((p0 - (long long)n0 * s.hw) / s.w);
      const CUtensorMap* map_sh = s.m_shifted ? &map_n : &map_m;      // unshifted: gy
      const CUtensorMap* map_tp = s.m_shifted ? &map_m : &map_n;      // shifted: x
      int dh[2], dx[2];
      for (int g = 0; g < 2; ++g) {
        const int tap = tap0 + (g < ng ? g : 0);
        const int ky = tap / s.ksize;
        dh[g] = ky - half;
        dx[g] = tap - ky * s.ksize - half;
      }
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&empty[stage], phase ^ 1u);
        mbar_expect_tx(&full[stage], stage_tx);
        unsigned char* base = smem + stage * kWStageBytes;
        for (int j = 0; j < shared_slabs; ++j)
          tma_load_5d(base + j * kWSlab, map_sh, s.gy_ch + j * 64, 0, s.gy_py, h0, n0, &full[stage]);
        for (int g = 0; g < ng; ++g)
          for (int j = 0; j < tap_slabs; ++j)
            tma_load_5d(base + (shared_slabs + g * tap_slabs + j) * kWSlab, map_tp, j * 64, dx[g], 0, h0 + dh[g], n0,
                        &full[stage]);
        if (++stage == kWStages) { stage = 0; phase ^= 1u; }
        h0 += s.ht;
        if (h0 >= s.h) { h0 = 0; n0 += s.nt; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, s.n) | (1u << 15) | (1u << 16);   // both MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + stage * kWStageBytes);
        for (int g = 0; g < ng; ++g) {
          const uint32_t tp = base + (uint32_t)(shared_slabs + g * tap_slabs) * kWSlab;
          const uint32_t a0 = s.m_shifted ? tp : base, b0 = s.m_shifted ? base : tp;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + (uint32_t)g * 256u, umma_desc_mn_sw128(a0 + k * 2048, kWSlab),
                      umma_desc_mn_sw128(b0 + k * 2048, kWSlab), idesc, (c > c_begin || k) ? 1u : 0u);
        }
        umma_commit(&empty[stage]);
        if (++stage == kWStages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(done);
    }
  } else if (c_begin < c_end) {
    mbar_wait(done, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int m = q * 32 + lane;
    for (int g = 0; g < ng; ++g) {
      float* drow = dw + ((size_t)(tap0 + g) * s.m + m) * s.n;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 256u;
      for (int c0 = 0; c0 < s.n; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) red_add_v4(drow + c0 + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 :: "r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

int g_sms = 0;

}  // namespace

// 5-D tensor map of an activation tensor with `c` channels whose GEMM pixels live on the (n, h, w)
// grid: up = 1 plain NHWC; up = 2 the tensor is stored on the doubled grid (n, 2h, 2w, c) and the
// phase (py, px) is selected by coordinates {px*c + channel, x, py, y, n}.
static int encode_act(CUtensorMap* map, const void* ptr, int64_t n, int h, int w, int c, int up, int box_w,
                      int box_h, int box_n, const char* who) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return emb::fail(-1, "%s: cuTensorMapEncodeTiled is not available from this driver", who);
  const cuuint64_t u = (cuuint64_t)up, cc = (cuuint64_t)c;
  const cuuint64_t dims[5] = {u * cc, (cuuint64_t)w, u, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t sx = u * cc * 2, sp = sx * w, sy = sp * u, sn = sy * h;
  const cuuint64_t strides[4] = {sx, sp, sy, sn};
  const cuuint32_t box[5] = {64, (cuuint32_t)box_w, 1, (cuuint32_t)box_h, (cuuint32_t)box_n};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return emb::fail(-3, "%s: cuTensorMapEncodeTiled(activations) failed with %d", who, (int)r);
  return 0;
}

static int sm_count(const char* who) {
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return emb::fail_cuda(who);
  }
  return 0;
}

// The largest tile of at most `px` consecutive pixels made of whole rows of one image (rows
// dividing the height) or of whole images (a count dividing n): box (w, rows, images).
static int tile_box(int64_t n, int h, int w, int px, int* ht, int* nt, const char* who) {
  if (w > px || w < 1 || h < 1) return emb::fail(-1, "%s: width %d must be in [1, %d]", who, w, px);
  if (h * w > px) {
    int r = px / w;
    while (h % r) --r;
    *ht = r; *nt = 1;
  } else {
    int k = px / (h * w);
    while (n % k) --k;
    *ht = h; *nt = k;
  }
  return 0;
}

extern "C" int emb_conv_nhwc_tc(const emb_conv_tc_args* a, void* stream) {
  const char* who = "emb_conv_nhwc_tc";
  if (!a) return emb::fail(-1, "%s: null args", who);
  const int64_t n = a->n;
  const int h = a->h, w = a->w, cin = a->cin, cout = a->cout, ksize = a->ksize;
  if (n <= 0) return 0;
  if (!a->in || !a->w_packed || !a->out) return emb::fail(-1, "%s: null pointer", who);
  if (ksize != 5 && ksize != 3 && ksize != 1) return emb::fail(-1, "%s: kernel size %d not in {1, 3, 5}", who, ksize);
  if (cin % 64) return emb::fail(-1, "%s: cin=%d must be a multiple of 64 (128-byte swizzled K blocks)", who, cin);
  if (cout % 32 || cout < 32 || cout > 256)
    return emb::fail(-1, "%s: cout=%d must be a multiple of 32 in [32, 256]", who, cout);
  if ((a->in_up != 1 && a->in_up != 2) || (a->out_up != 1 && a->out_up != 2) || a->out_phase < 0 ||
      a->out_phase > 3)
    return emb::fail(-1, "%s: in_up / out_up must be 1 or 2, out_phase in [0, 4)", who);
  int ht, nt;
  if (int e = tile_box(n, h, w, 128, &ht, &nt, who)) return e;
  const int tpix = ht * w * nt;
  const int64_t ntiles = n * h * w / tpix;
  if (((uintptr_t)a->in | (uintptr_t)a->w_packed | (uintptr_t)a->out) & 15)
    return emb::fail(-1, "%s: pointers must be 16-byte aligned", who);
  if (int e = sm_count(who)) return e;
  const int groups = a->in_up == 2 ? 4 : 1;
  CUtensorMap map_in, map_w;
  if (int e = encode_act(&map_in, a->in, n, h, w, cin, a->in_up, w, ht, nt, who)) return e;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)cout, (cuuint64_t)(groups * ksize * ksize)};
    const cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cout * cin * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)cout, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode_fn()(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a->w_packed),
                                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return emb::fail(-3, "%s: cuTensorMapEncodeTiled(weights) failed with %d", who, (int)r);
  }
  ConvShape s;
  s.msub = (cout <= 128 && ntiles % 2 == 0) ? 2 : 1;
  s.tiles = (int)(ntiles / s.msub);
  s.tpix = tpix;
  s.hw = h * w;
  s.w = w;
  s.kblocks = cin / 64;
  s.cout = cout;
  s.taps = ksize * ksize;
  s.ksize = ksize;
  s.groups = groups;
  s.cin = cin;
  s.out_up = a->out_up;
  s.out_py = a->out_phase >> 1;
  s.out_px = a->out_phase & 1;
  const size_t smem = (size_t)kStages * kStageBytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute((const void*)conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute((const void*)conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return emb::fail_cuda(who);
    attr_set = true;
  }
  const int grid = s.tiles < g_sms ? s.tiles : g_sms;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a->out);
  if (s.msub == 2)
    conv_tc_kernel<2><<<grid, kThreads, smem, (cudaStream_t)stream>>>(map_in, map_w, out, a->bias, s);
  else
    conv_tc_kernel<1><<<grid, kThreads, smem, (cudaStream_t)stream>>>(map_in, map_w, out, a->bias, s);
  if (cudaGetLastError() != cudaSuccess) return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}

extern "C" int emb_conv5x5_nhwc_tc(const void* in, const void* w_packed, const float* bias, void* out,
                                   int64_t n, int32_t h, int32_t w, int32_t cin, int32_t cout,
                                   int32_t ksize, void* stream) {
  emb_conv_tc_args a;
  a.in = in; a.w_packed = w_packed; a.bias = bias; a.out = out;
  a.n = n; a.h = h; a.w = w; a.cin = cin; a.cout = cout; a.ksize = ksize;
  a.in_up = 1; a.out_up = 1; a.out_phase = 0;
  return emb_conv_nhwc_tc(&a, stream);
}

// dw[tap][m][n] (fp32, accumulated into) for the SAME convolution whose input is `x` (n, h, w, cin)
// and whose output gradient is `gy`: (n, h, w, cout), or with gy_up = 2 phase gy_phase of
// (n, 2h, 2w, cout).  m_is_in = 1: m = cin, n = cout (dw is HWIO); 0: m = cout, n = cin.
extern "C" int emb_conv_wgrad_tc(const emb_conv_wgrad_tc_args* a, void* stream) {
  const char* who = "emb_conv_wgrad_tc";
  if (!a) return emb::fail(-1, "%s: null args", who);
  const int64_t n = a->n;
  const int h = a->h, w = a->w, cin = a->cin, cout = a->cout, ksize = a->ksize;
  if (n <= 0) return 0;
  if (!a->x || !a->gy || !a->dw) return emb::fail(-1, "%s: null pointer", who);
  if (ksize != 5 && ksize != 3 && ksize != 1) return emb::fail(-1, "%s: kernel size %d not in {1, 3, 5}", who, ksize);
  const int m = a->m_is_in ? cin : cout, nn = a->m_is_in ? cout : cin;
  if (m != 128 && m != 256) return emb::fail(-1, "%s: the M side has %d channels, need 128 or 256", who, m);
  // the N side may have any multiple of 8 channels (16-byte rows for TMA): its last 64-channel slab is
  // then partly out of bounds, which TMA fills with zeros, and dw has n_pad = 64 * ceil(n / 64) columns
  if (nn % 8 || nn < 8 || nn > 256) return emb::fail(-1, "%s: the N side has %d channels, need a multiple of 8 <= 256", who, nn);
  const int n_pad = (nn + 63) / 64 * 64;
  if ((a->gy_up != 1 && a->gy_up != 2) || a->gy_phase < 0 || a->gy_phase > 3)
    return emb::fail(-1, "%s: gy_up must be 1 or 2, gy_phase in [0, 4)", who);
  // chunk = box (bw, ht, nt) of at most 64 pixels: whole rows / whole images where a row fits,
  // else the widest piece of a row that divides the width
  int ht, nt, bw = w;
  if (w > 64) {
    bw = 64;
    while (w % bw) --bw;
    ht = 1; nt = 1;
  } else if (int e = tile_box(n, h, w, 64, &ht, &nt, who)) {
    return e;
  }
  const int cpix = bw * ht * nt;
  if (((uintptr_t)a->x | (uintptr_t)a->gy | (uintptr_t)a->dw) & 15)
    return emb::fail(-1, "%s: pointers must be 16-byte aligned", who);
  if (int e = sm_count(who)) return e;
  CUtensorMap map_x, map_gy;
  if (int e = encode_act(&map_x, a->x, n, h, w, cin, 1, bw, ht, nt, who)) return e;
  if (int e = encode_act(&map_gy, a->gy, n, h, w, cout, a->gy_up, bw, ht, nt, who)) return e;
  CUtensorMap* maps[2] = {a->m_is_in ? &map_x : &map_gy, a->m_is_in ? &map_gy : &map_x};
  WgradShape s;
  s.chunks = (int)(n * h * w / cpix);
  s.cpix = cpix;
  s.bw = bw;
  const int taps = ksize * ksize;
  s.splits = g_sms / taps > 0 ? g_sms / taps : 1;
  if (s.splits > s.chunks) s.splits = s.chunks;
  s.hw = h * w;
  s.w = w;
  s.m = m;
  s.n = n_pad;
  s.ksize = ksize;
  s.m_shifted = a->m_is_in ? 1 : 0;
  s.h = h;
  s.ht = ht;
  s.nt = nt;
  s.gy_py = a->gy_up == 2 ? a->gy_phase >> 1 : 0;
  s.gy_ch = a->gy_up == 2 ? (a->gy_phase & 1) * cout : 0;
  const size_t smem = (size_t)kWRingSlabs * kWSlab + 1024 + 256;
  const void* fn = nullptr;
  // two taps per CTA where the M side has 128 channels and a stage of both taps stays within 8
  // slabs (three ring stages); whole-row chunks only; EMB_WGRAD_PAIR=0 keeps one tap per CTA
  static int pair_on = -1;
  if (pair_on < 0) {
    const char* e = getenv("EMB_WGRAD_PAIR");
    pair_on = (e && e[0] == '0') ? 0 : 1;
  }
  const int pair_slabs = s.m_shifted ? n_pad / 64 + 2 * (m / 64) : m / 64 + 2 * (n_pad / 64);
  if (pair_on && m == 128 && taps > 1 && bw == w && pair_slabs <= 8) {
    const int ngroups = (taps + 1) / 2;
    s.splits = g_sms / ngroups > 0 ? g_sms / ngroups : 1;
    if (s.splits > s.chunks) s.splits = s.chunks;
    switch (pair_slabs) {
      case 5: fn = (const void*)conv_wgrad_pair_kernel<5>; break;
      case 6: fn = (const void*)conv_wgrad_pair_kernel<6>; break;
      case 7: fn = (const void*)conv_wgrad_pair_kernel<7>; break;
      default: fn = (const void*)conv_wgrad_pair_kernel<8>; break;
    }
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return emb::fail_cuda(who);
    float* dwp = a->dw;
    void* pargs[] = {maps[0], maps[1], &dwp, &s};
    if (cudaLaunchKernel(fn, dim3(ngroups * s.splits), dim3(kThreads), pargs, smem, (cudaStream_t)stream) !=
        cudaSuccess)
      return emb::fail_cuda(who);
    emb::count_launch();
    return 0;
  }
  switch ((m + n_pad) / 64) {
    case 3: fn = (const void*)conv_wgrad_tc_kernel<3>; break;
    case 4: fn = (const void*)conv_wgrad_tc_kernel<4>; break;
    case 5: fn = (const void*)conv_wgrad_tc_kernel<5>; break;
    case 6: fn = (const void*)conv_wgrad_tc_kernel<6>; break;
    case 7: fn = (const void*)conv_wgrad_tc_kernel<7>; break;
    default: fn = (const void*)conv_wgrad_tc_kernel<8>; break;
  }
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return emb::fail_cuda(who);
  float* dw = a->dw;
  void* params[] = {maps[0], maps[1], &dw, &s};
  if (cudaLaunchKernel(fn, dim3(taps * s.splits), dim3(kThreads), params, smem, (cudaStream_t)stream) !=
      cudaSuccess)
    return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}

extern "C" int emb_conv5x5_wgrad_tc(const void* x, const void* gy, float* dw, int64_t n, int32_t h,
                                    int32_t w, int32_t cin, int32_t cout, int32_t ksize,
                                    int32_t m_is_in, void* stream) {
  emb_conv_wgrad_tc_args a;
  a.x = x; a.gy = gy; a.dw = dw;
  a.n = n; a.h = h; a.w = w; a.cin = cin; a.cout = cout; a.ksize = ksize;
  a.m_is_in = m_is_in; a.gy_up = 1; a.gy_phase = 0;
  return emb_conv_wgrad_tc(&a, stream);
}
