// Weight streaming for the bf16 scan kernels (rssm_fwd_tma.cu, rssm_bwd_tma.cu):
// a dedicated producer warp walks the STATIC per-CTA weight schedule of the
// whole scan (T steps x phases) and moves it HBM -> shared memory with TMA bulk
// copies (cp.async.bulk ... mbarrier::complete_tx) through a ring of stages.
// The weights do not depend on the recurrent state, so the producer runs ahead
// of the consumers across phase boundaries and grid barriers: HBM keeps
// streaming while the eight consumer warps sit in a barrier or build the next
// A operand.  Consumers read mma B fragments from the ring (conflict-free 8-byte
// LDS), release a stage with one mbarrier arrive per warp.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rssm_common.cuh"

namespace rssm_tma {

using namespace rssm;

constexpr int kCWarps = 8;                    // consumer warps
constexpr int kCThreads = kCWarps * 32;       // 256
constexpr int kAllThreads = kCThreads + 32;   // + the producer warp
constexpr int kStageBytesDefault = 24576;   // bytes per ring stage (Ring::stage_bytes)
constexpr int kMaxPer = 48;                   // n8 tiles per CTA and layer (6 per warp)
constexpr int kAGlobal = 6;                   // A fragments per warp and stage when A lives in global memory
constexpr int kADepth = 4;                    // stages of A a warp keeps in flight (cp.async, private ring)
constexpr int kAPrivBytes = kCWarps * kADepth * kAGlobal * 512;

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
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, no tensor map: the blocks are contiguous),
// completion reported as `bytes` transaction bytes on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// The weight stream is read once per step and is larger than L2: mark its lines
// evict-first so that it does not flush the (small, latency-critical) activation
// working set the phases hand to each other through L2.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;"
      :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// generic-proxy writes (st.global / st.shared) -> visible to later async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
// barrier among the consumer warps only (the producer warp never joins)
__device__ __forceinline__ void cbar() {
  asm volatile("bar.sync 1, %0;" :: "n"(kCThreads) : "memory");
}

struct Ring {                // shared-memory ring + this thread's cursor (same sequence on both sides)
  uint64_t* full;            // [nstages] count 1 (+ tx bytes)
  uint64_t* empty;           // [nstages] count kCWarps
  unsigned char* data;       // [nstages][stage_bytes]
  int nstages;
  int stage_bytes;
  int stage;
  uint32_t phase;
};

// How the eight consumer warps share a layer of `per` n8 tiles per CTA: TG tile
// groups x KW k-lanes, at least two tiles per warp where possible (independent
// mma chains; every A fragment is read by TG warps only).
__host__ __device__ __forceinline__ int tile_groups(int per) {
  int tg = 1;
  while (tg < kCWarps && tg * 4 <= per) tg <<= 1;
  return tg;
}
// Tiles per CTA are padded (with all-zero tiles, by the packer) until the tile
// groups divide them: the mma loops then carry no guards.  `unit` = tiles that
// must stay together (3 for the GRU's gate triples).
__host__ __device__ __forceinline__ int pad_tiles(int per_raw, int unit) {
  int per = per_raw;
  while (per % tile_groups(per)) per += unit;
  return per;
}
// k16 steps per stage for a layer whose CTA block has `per` n8 tiles (256 B per
// tile and step): a multiple of the k-lanes KW; with A in global memory exactly
// kAGlobal fragments per warp and stage (register double buffer).
__device__ __forceinline__ int ksteps_per_chunk(int stage_bytes, int per, bool a_global) {
  const int kw = kCWarps / tile_groups(per);
  int kc = stage_bytes / (per * 256);
  if (a_global && kc > kAGlobal * kw) kc = kAGlobal * kw;
  kc = kc / kw * kw;
  return kc < kw ? kw : kc;          // (kw * per * 256 <= stage_bytes is checked on the host)
}

// Host mirror of ksteps_per_chunk (launch-time validation).
inline int host_ksteps_per_chunk(int stage_bytes, int per, bool a_global) {
  const int kw = kCWarps / tile_groups(per);
  int kc = stage_bytes / (per * 256);
  if (a_global && kc > kAGlobal * kw) kc = kAGlobal * kw;
  kc = kc / kw * kw;
  return kc < kw ? kw : kc;
}

// Producer side: stream `ksteps` k16 steps of one CTA block.
__device__ __forceinline__ void produce(Ring& r, const unsigned char* blk, int per, int ksteps,
                                        bool a_global) {
  const uint32_t step_bytes = (uint32_t)per * 256u;
  const int kc = ksteps_per_chunk(r.stage_bytes, per, a_global);
  const uint64_t pol = policy_evict_first();
  for (int k0 = 0; k0 < ksteps; k0 += kc) {
    const int n = min(kc, ksteps - k0);
    mbar_wait(&r.empty[r.stage], r.phase ^ 1u);
    mbar_expect_tx(&r.full[r.stage], n * step_bytes);
    bulk_g2s_hint(r.data + (size_t)r.stage * r.stage_bytes, blk + (size_t)k0 * step_bytes,
                  n * step_bytes, &r.full[r.stage], pol);
    if (++r.stage == r.nstages) { r.stage = 0; r.phase ^= 1u; }
  }
}

// ---- look-ahead producer ------------------------------------------------------
// The ring holds only a few microseconds of stream, so HBM used to idle whenever
// the consumers sat in a grid barrier, an operand build or an epilogue.  The
// schedule is static, so the producer also runs an L2 PREFETCH cursor `ahead`
// chunks in front of the ring's load cursor (cp.async.bulk.prefetch.L2): HBM keeps
// filling L2 (126 MB) through every stall, and the ring then refills at L2 rate.
struct Seg {                 // one weight block streamed per step
  const unsigned char* blk;
  int per, ksteps, a_global;
};
constexpr int kMaxSegs = 6;
static_assert(sizeof(Seg) * kMaxSegs <= 256, "segment table must fit its 256-byte slot");
struct Cursor { int t, s, k0; };

__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(bytes) : "memory");
}

// Next chunk of the periodic schedule (`nseg` segments per step, `steps` steps; segment s is
// left out of the LAST step when bit s of skip_last is set).  False at the end.
__device__ __forceinline__ bool next_chunk(Cursor& c, const Seg* segs, int nseg, int steps,
                                           uint32_t skip_last, int stage_bytes,
                                           const unsigned char*& ptr, uint32_t& bytes) {
  for (;;) {
    if (c.t >= steps) return false;
    if (c.s >= nseg) { c.s = 0; c.k0 = 0; ++c.t; continue; }
    const Seg g = segs[c.s];
    const bool skip = (c.t + 1 == steps) && ((skip_last >> c.s) & 1u);
    if (skip || c.k0 >= g.ksteps) { ++c.s; c.k0 = 0; continue; }
    const int kc = ksteps_per_chunk(stage_bytes, g.per, g.a_global != 0);
    const int n = min(kc, g.ksteps - c.k0);
    const uint32_t step_bytes = (uint32_t)g.per * 256u;
    ptr = g.blk + (size_t)c.k0 * step_bytes;
    bytes = (uint32_t)n * step_bytes;
    c.k0 += kc;
    return true;
  }
}

// The producer thread's whole life: `segs` lives in shared memory.
static __device__ __noinline__ void run_producer(Ring r, const Seg* segs, int nseg, int steps,
                                          uint32_t skip_last, int ahead) {
  Cursor ld{0, 0, 0}, pf{0, 0, 0};
  const unsigned char *ptr, *pp;
  uint32_t bytes, pb;
  bool pf_on = ahead > 0;
  for (int i = 0; i < ahead && pf_on; ++i) {
    pf_on = next_chunk(pf, segs, nseg, steps, skip_last, r.stage_bytes, pp, pb);
    if (pf_on && i >= r.nstages) bulk_prefetch_l2(pp, pb);    // the first chunks go straight to the ring
  }
  const uint64_t pol = policy_evict_first();
  while (next_chunk(ld, segs, nseg, steps, skip_last, r.stage_bytes, ptr, bytes)) {
    if (pf_on) {
      pf_on = next_chunk(pf, segs, nseg, steps, skip_last, r.stage_bytes, pp, pb);
      if (pf_on) bulk_prefetch_l2(pp, pb);
    }
    mbar_wait(&r.empty[r.stage], r.phase ^ 1u);
    mbar_expect_tx(&r.full[r.stage], bytes);
    bulk_g2s_hint(r.data + (size_t)r.stage * r.stage_bytes, ptr, bytes, &r.full[r.stage], pol);
    if (++r.stage == r.nstages) { r.stage = 0; r.phase ^= 1u; }
  }
}

// explicit shared-space loads: inside the non-inlined consume() the compiler cannot prove
// that the ring / operand pointers are shared memory and would emit generic loads
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

__device__ __forceinline__ void mma_bf16_nv(float (&c)[4], const uint4& a, const uint2& b) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 "
      "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y));
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Consumer side: out[16][per*8] (shared, fp32) (+)= A[16][ksteps*16] @ W(this CTA's block).
// per == TG * TW exactly.  Warp w owns tiles [tgi*TW, tgi*TW+TW) (tgi = w % TG)
// and the k16 steps == w / TG (mod KW) of every chunk.
// A_GLOBAL false: `afrag` = A fragments in shared memory ([ks][32] uint4).
// A_GLOBAL true:  `afrag` = A fragments in global memory (L2); every lane pulls
//   its own 16-byte fragments kADepth stages ahead with cp.async into a private
//   shared-memory ring (`apriv`, kAPrivBytes for the CTA) -- no registers held.
// `first` false accumulates onto `out` (a layer consumed in two k ranges).
// The k-lanes' partial sums meet in shared-memory slabs BEHIND `out` (out must hold
// 16 x per*8*KW floats: out_tiles() below), summed in a fixed order by lane 0's warps:
// no atomics, bit-reproducible.  Ends with a consumer barrier.
// The fragment loads of k16 step i+1 are issued before the mma of step i (the loads are
// volatile asm, so program order is issue order), and layers with few tiles per warp keep
// two accumulator sets: one warp's chunk is otherwise a serial LDS -> HMMA -> HMMA chain
// (measured 0.33 us per 24 KiB chunk with the data already in the ring -- slower than HBM).
// (Not inlined: one body per (TW, A_GLOBAL) for all call sites keeps the
// per-step instruction footprint inside the instruction cache.)
__host__ __device__ __forceinline__ int out_tiles(int per) {        // tiles of `out` a layer needs
  return per * (kCWarps / tile_groups(per));
}

template <int TW, bool A_GLOBAL>
__device__ __noinline__ uint32_t consume(Ring r, int per, int ksteps, const uint4* __restrict__ afrag,
                                         unsigned char* apriv, float* out, bool first) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int TG = tile_groups(per), KW = kCWarps / TG;
  const int tgi = warp % TG, kl = warp / TG;
  const int kc = ksteps_per_chunk(r.stage_bytes, per, A_GLOBAL);
  const int ncols = per * 8;
  const uint32_t kstride = (uint32_t)per * 256u * (uint32_t)KW;      // bytes between this warp's k16 steps
  const uint32_t wofs = ((uint32_t)(tgi * TW) * 32u + lane) * 8u + (uint32_t)kl * per * 256u;
  const uint4* ap = afrag + (size_t)kl * 32 + lane;                  // this warp's first fragment
  const size_t astride = (size_t)KW * 32;                            // uint4 between its k16 steps
  constexpr int NA = 1;                                              // accumulator sets
  float acc[NA][TW][4];
#pragma unroll
  for (int s = 0; s < NA; ++s)
#pragma unroll
    for (int j = 0; j < TW; ++j) acc[s][j][0] = acc[s][j][1] = acc[s][j][2] = acc[s][j][3] = 0.f;
  int stage = r.stage;
  uint32_t phase = r.phase;
  const int nstages = r.nstages;
  const uint32_t data_s = smem_u32(r.data);
  const int total = (ksteps - kl + KW - 1) / KW;                     // this warp's k16 steps in all
  const int per_chunk = kc / KW;
  if (A_GLOBAL) {
    constexpr uint32_t slot_bytes = kAGlobal * 512;
    unsigned char* mine = apriv + (size_t)warp * (kADepth * slot_bytes) + lane * 16;
    const int nchunks = (ksteps + kc - 1) / kc;
    int islot = 0;
    auto issue = [&](int c) {
      if (c < nchunks) {
#pragma unroll
        for (int i = 0; i < kAGlobal; ++i)
          if (i < per_chunk && c * per_chunk + i < total)
            cp_async16(mine + islot * slot_bytes + i * 512, ap + (size_t)(c * per_chunk + i) * astride);
      }
      cp_async_commit();
      if (++islot == kADepth) islot = 0;
    };
    for (int c = 0; c < kADepth - 1; ++c) issue(c);
    int cslot = 0;
    for (int c = 0; c < nchunks; ++c) {
      issue(c + kADepth - 1);
      cp_async_wait<kADepth - 1>();
      mbar_wait(&r.full[stage], phase);
      const uint32_t st = data_s + (uint32_t)stage * (uint32_t)r.stage_bytes + wofs;
      const uint32_t as = smem_u32(mine) + cslot * slot_bytes;
#pragma unroll
      for (int i = 0; i < kAGlobal; ++i) {
        // fragments beyond the chunk are zero: the multiply is harmless, its B
        // address is clamped into the stage
        const bool on = i < per_chunk && c * per_chunk + i < total;
        const uint4 av = on ? lds_u4(as + i * 512) : make_uint4(0, 0, 0, 0);
        const uint32_t bp = st + (uint32_t)(on ? i : 0) * kstride;
        uint2 b[TW];
#pragma unroll
        for (int j = 0; j < TW; ++j) b[j] = lds_u2(bp + j * 256);
#pragma unroll
        for (int j = 0; j < TW; ++j) mma_bf16_nv(acc[0][j], av, b[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&r.empty[stage]);
      if (++stage == nstages) { stage = 0; phase ^= 1u; }
      if (++cslot == kADepth) cslot = 0;
    }
    cp_async_wait<0>();
  } else {
    for (int j0 = 0; j0 * KW < ksteps; j0 += per_chunk) {
      const int n = min(per_chunk, total - j0);                      // <= 0 for a lane past the tail
      mbar_wait(&r.full[stage], phase);
      const uint32_t st = data_s + (uint32_t)stage * (uint32_t)r.stage_bytes + wofs;
      const uint32_t a0 = smem_u32(ap) + (uint32_t)j0 * (uint32_t)astride * 16u;
#pragma unroll 2
      for (int i = 0; i < n; ++i) {
        const uint4 av = lds_u4(a0 + (uint32_t)i * (uint32_t)astride * 16u);
        const uint32_t bp = st + (uint32_t)i * kstride;
        uint2 b[TW];
#pragma unroll
        for (int j = 0; j < TW; ++j) b[j] = lds_u2(bp + j * 256);
#pragma unroll
        for (int j = 0; j < TW; ++j) mma_bf16_nv(acc[0][j], av, b[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&r.empty[stage]);
      if (++stage == nstages) { stage = 0; phase ^= 1u; }
    }
  }
  if (NA == 2) {
#pragma unroll
    for (int j = 0; j < TW; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[0][j][e] += acc[NA - 1][j][e];
  }
  const int g = lane >> 2, q = lane & 3;
  if (KW > 1) {
    // k-lanes 1.. leave their partial tiles in the slabs behind `out`; k-lane 0 sums them in order
    if (kl > 0) {
      float* slab = out + (size_t)kl * kRows * ncols;
#pragma unroll
      for (int j = 0; j < TW; ++j) {
        float* o = slab + (tgi * TW + j) * 8 + 2 * q;
        *reinterpret_cast<float2*>(o + g * ncols) = make_float2(acc[0][j][0], acc[0][j][1]);
        *reinterpret_cast<float2*>(o + (g + 8) * ncols) = make_float2(acc[0][j][2], acc[0][j][3]);
      }
    }
    cbar();
  }
  if (kl == 0) {
#pragma unroll
    for (int j = 0; j < TW; ++j) {
      float* o = out + (tgi * TW + j) * 8 + 2 * q;
      float2 lo = make_float2(acc[0][j][0], acc[0][j][1]), hi = make_float2(acc[0][j][2], acc[0][j][3]);
      for (int k = 1; k < KW; ++k) {
        const float2 plo = *reinterpret_cast<const float2*>(o + (size_t)k * kRows * ncols + g * ncols);
        const float2 phi = *reinterpret_cast<const float2*>(o + (size_t)k * kRows * ncols + (g + 8) * ncols);
        lo.x += plo.x; lo.y += plo.y; hi.x += phi.x; hi.y += phi.y;
      }
      if (!first) {
        const float2 plo = *reinterpret_cast<const float2*>(o + g * ncols);
        const float2 phi = *reinterpret_cast<const float2*>(o + (g + 8) * ncols);
        lo.x += plo.x; lo.y += plo.y; hi.x += phi.x; hi.y += phi.y;
      }
      *reinterpret_cast<float2*>(o + g * ncols) = lo;
      *reinterpret_cast<float2*>(o + (g + 8) * ncols) = hi;
    }
  }
  cbar();
  return (uint32_t)stage | (phase << 8);
}

#define EMB_CONSUME(AG, ring, per, ksteps, afrag, apriv, out, first)                                   \
  {                                                                                                    \
    uint32_t cur_;                                                                                     \
    switch ((per) / rssm_tma::tile_groups(per)) {                                                      \
      case 1: cur_ = rssm_tma::consume<1, AG>(ring, per, ksteps, afrag, apriv, out, first); break;     \
      case 2: cur_ = rssm_tma::consume<2, AG>(ring, per, ksteps, afrag, apriv, out, first); break;     \
      case 3: cur_ = rssm_tma::consume<3, AG>(ring, per, ksteps, afrag, apriv, out, first); break;     \
      case 4: cur_ = rssm_tma::consume<4, AG>(ring, per, ksteps, afrag, apriv, out, first); break;     \
      case 5: cur_ = rssm_tma::consume<5, AG>(ring, per, ksteps, afrag, apriv, out, first); break;     \
      default: cur_ = rssm_tma::consume<6, AG>(ring, per, ksteps, afrag, apriv, out, first); break;    \
    }                                                                                                  \
    ring.stage = (int)(cur_ & 0xff);                                                                   \
    ring.phase = cur_ >> 8;                                                                            \
  }

// Grid barrier for the consumer warps (monotonic counter, one arrival per CTA).
// Every thread first orders its generic-proxy global writes before later TMA reads.
struct GridBarrierC {
  unsigned* counter;
  unsigned epoch;
  __device__ __forceinline__ void sync() {
    fence_proxy_async();
    cbar();
    if (threadIdx.x == 0) {
      // release / acquire at gpu scope instead of two full fences: the bar.sync above
      // orders the CTA's writes before thread 0's release, the one below hands the
      // acquired view to the other threads (cross-CTA data is read through L2)
      epoch += gridDim.x;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(counter) : "memory");
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      } while ((int)(v - epoch) < 0);
    }
    cbar();
  }
};

}  // namespace rssm_tma
