// emb_rssm_observe_fwd, bf16 engine: the T-step recurrent scan of RSSM.observe
// (dreamerv3/rssm.py:61-92 `_observe`, :135-159 `_core`, embodied/jax/outs.py:208-270)
// as ONE persistent cooperative kernel whose weights arrive by TMA.
//
// Per CTA: eight consumer warps + one producer warp (rssm_tma.cuh).  The
// producer streams this CTA's weight blocks of every phase of every step, in
// order, HBM -> shared-memory ring, never waiting for a grid barrier; consumers
// do mma.sync m16n8k16 on the ring's B fragments and the phase's A fragments.
//
// Per step t (phases separated by grid barriers; every weight read once):
//   P4  yhid  = [keep*deter_g, x0, x1] @ dynhid0[g][:Dg+2H] + hid_pre[t]    (25.2 M weights at size200m)
//         A = three TMA bulk copies of prebuilt fragment buffers (32 KiB each)
//   P5  gates = silu(rms(yhid))_g @ dyngru[g] + b ; GRU -> deter_t           (25.2 M)
//   P1  yobs  = deter_t @ obs0[:D] + pre_tok_t ;  y0' = keep'*(deter_t @ dynin0) + b   (16.8 M)
//         A = deter fragments in global memory (L2), prefetched per stage
//   P2  logit = silu(rms(yobs)) @ obslogit + b ; then x0' = silu(rms(y0')) by 16 row CTAs   (2.1 M)
//   P3  16 row CTAs: idx = argmax(log unimix(softmax(logit)) + gumbel) ;
//       y1' = keep' * sum_s dynin1[s*C+idx_s] + b ; x1' = silu(rms(y1'))  (row gather, no GEMM)
// Hoisted by the caller (no dependence on the recurrent state): the action
// branch's contribution to dynhid0 and its bias (hid_pre = x2 @ dynhid0[g][Dg+2H:] + b),
// pre_tok = tokens @ obs0[D:] + b, and step 0's y0 / y1.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"
#include "rssm_common.cuh"
#include "rssm_tma.cuh"

namespace emb_tma {
// A region of the forward kernel: P4's [deter_g | x0 | x1] fragments, P5 / P2's
// consumer-built operands, and P1's private cp.async rings.
__host__ __device__ inline size_t a_region_bytes(const emb_rssm_fwd_args& a) {
  const int Dg = a.D / a.G, Kh = Dg + 2 * a.H;
  int amax = Kh > a.H ? Kh : a.H;
  if (Dg > amax) amax = Dg;
  size_t n = (size_t)rssm::kRows * amax * 2;
  if (n < (size_t)rssm_tma::kAPrivBytes) n = rssm_tma::kAPrivBytes;
  return (n + 127) & ~(size_t)127;
}
}  // namespace emb_tma

namespace {

using namespace rssm;
using namespace rssm_tma;

struct Plan {                 // static work split of one CTA; identical on producer and consumers
  int per_hid, per_gru, per_ph1, per_log;          // n8 tiles per CTA block, padded (gru: 3 per unit)
  int wk;
  int raw_ph1;
  int ks_hid, ks_gru, ks_ph1, ks_log;              // k16 steps
  bool on_hid, on_gru, on_log;
  int u0_hid, u1_hid, u0_gru, u1_gru, u0_log, u1_log;
  const unsigned char *blk_hid, *blk_gru, *blk_ph1, *blk_log;
};

// CTA -> batch row it serves in the sampling phase (or -1): CTAs WITHOUT tiles in the
// block-diagonal layers, so that sampling runs underneath the other CTAs' dynhid0; the last
// 16 CTAs when there are too few of those.
__device__ __forceinline__ bool rows_on_spare_ctas(const emb_rssm_fwd_args& a) {
  const int ncta = gridDim.x, G = a.G, units = a.D / a.G / 8;
  const int cpg = max(1, ncta / G), per = (units + cpg - 1) / cpg, used = (units + per - 1) / per;
  return (cpg - used) * G + (ncta - G * cpg) >= kRows;
}
__device__ __forceinline__ int row_of_cta(const emb_rssm_fwd_args& a, int cta) {
  const int ncta = gridDim.x, G = a.G, units = a.D / a.G / 8;
  const int cpg = max(1, ncta / G), per = (units + cpg - 1) / cpg, used = (units + per - 1) / per;
  const int spare = cpg - used, tail = ncta - G * cpg;
  int row = ncta - 1 - cta;
  if (rows_on_spare_ctas(a)) {
    const int g = cta / cpg, j = cta - g * cpg;
    if (g >= G) row = G * spare + (cta - G * cpg);
    else row = j >= used ? g * spare + (j - used) : -1;
  }
  return row >= 0 && row < kRows ? row : -1;
}

// `wk` = this CTA's index among the CTAs that do not serve a row (they share the two dense
// layers obs0 | dynin0 and obslogit; a row CTA passes gridDim.x and owns no tiles).
__device__ __forceinline__ Plan make_plan(const emb_rssm_fwd_args& a, int wk) {
  Plan p;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int D = a.D, H = a.H, Dg = a.D / a.G, SC = a.S * a.C, Kh = Dg + 2 * H;
  const GroupSplit sh = group_split(Dg / 8, a.G), sg = group_split(Dg / 8, a.G);
  // ownership by the raw run length, block stride by the padded one (scan.py pack_matrix)
  p.per_hid = pad_tiles(sh.per, 1); p.u0_hid = sh.u0; p.u1_hid = sh.u1; p.on_hid = sh.u0 < sh.u1;
  p.per_gru = pad_tiles(sg.per * 3, 3); p.u0_gru = sg.u0; p.u1_gru = sg.u1; p.on_gru = sg.u0 < sg.u1;
  p.raw_ph1 = (2 * H / 8 + ncta - 1) / ncta;
  p.per_ph1 = pad_tiles(p.raw_ph1, 1);
  const int tiles_log = SC / 8, raw_log = (tiles_log + ncta - 1) / ncta;
  p.per_log = pad_tiles(raw_log, 1);
  p.wk = wk;
  p.u0_log = min(tiles_log, wk * raw_log); p.u1_log = min(tiles_log, p.u0_log + raw_log);
  p.on_log = p.u0_log < p.u1_log;
  p.ks_hid = Kh / 16; p.ks_gru = Dg / 16; p.ks_ph1 = D / 16; p.ks_log = H / 16;
  p.blk_hid = reinterpret_cast<const unsigned char*>(a.w_hid) + (size_t)cta * p.ks_hid * p.per_hid * 256;
  p.blk_gru = reinterpret_cast<const unsigned char*>(a.w_gru) + (size_t)cta * p.ks_gru * p.per_gru * 256;
  const int wb = min(wk, ncta - 1);
  p.blk_ph1 = reinterpret_cast<const unsigned char*>(a.w_ph1) + (size_t)wb * p.ks_ph1 * p.per_ph1 * 256;
  p.blk_log = reinterpret_cast<const unsigned char*>(a.w_logit) + (size_t)wb * p.ks_log * p.per_log * 256;
  return p;
}

// any-of over the consumer threads (named barrier + shared flag)
__device__ __forceinline__ bool __syncthreads_or_consumers(bool pred, float* scratch) {
  int* flag = reinterpret_cast<int*>(scratch + 16);
  if (threadIdx.x == 0) *flag = 0;
  cbar();
  if (pred) atomicOr(flag, 1);
  cbar();
  return *flag != 0;
}

// P1 covers yobs (H columns) and, except at the last step, y0' (H more).
__device__ __forceinline__ void ph1_range(const emb_rssm_fwd_args& a, const Plan& p, bool last,
                                          int& u0, int& u1) {
  const int total = (last ? a.H : 2 * a.H) / 8;
  u0 = min(total, p.wk * p.raw_ph1);
  u1 = min(total, u0 + p.raw_ph1);
}

// One row of H values held in registers by the 256 consumer threads (thread i
// owns columns 4i..4i+3 of every 1024-column group): x = silu(rms(y) * scale)
// -> A fragments of row `row` in `dstA`; also stores y (fp32) and its rstd.
constexpr int kRowGroups = 4;                      // H <= 4096
struct RowVals { float4 v[kRowGroups]; };

__device__ __noinline__ void finish_row(const RowVals y, int H, int row, const float* __restrict__ scale,
                                           float eps, float* red, float* ydst, float* rstd_dst,
                                           __nv_bfloat16* dstA) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float s = 0.f;
#pragma unroll
  for (int gI = 0; gI < kRowGroups; ++gI) {
    const int c = gI * kCThreads * 4 + tid * 4;
    if (c < H) {
      const float4 v = y.v[gI];
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      if (ydst) *reinterpret_cast<float4*>(ydst + c) = v;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  cbar();                                          // `red` may still be read by a previous call
  if (lane == 0) red[warp] = s;
  cbar();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kCWarps; ++w) tot += red[w];
  const float rstd = rsqrtf(tot / (float)H + eps);
  if (tid == 0) *rstd_dst = rstd;
#pragma unroll
  for (int gI = 0; gI < kRowGroups; ++gI) {
    const int c = gI * kCThreads * 4 + tid * 4;
    if (c < H) {
      const float4 v = y.v[gI];
      const float4 sc = *reinterpret_cast<const float4*>(scale + c);
      *reinterpret_cast<__nv_bfloat162*>(dstA + afrag_index(row, c)) = __floats2bfloat162_rn(
          silu_fast(v.x * (rstd * sc.x)), silu_fast(v.y * (rstd * sc.y)));
      *reinterpret_cast<__nv_bfloat162*>(dstA + afrag_index(row, c + 2)) = __floats2bfloat162_rn(
          silu_fast(v.z * (rstd * sc.z)), silu_fast(v.w * (rstd * sc.w)));
    }
  }
}

// A fragments (shared) of f(y[r][k]) for a [16][n] fp32 global slice, every
// thread's loads issued before any use (n * 16 / 1024 <= 16 float4 per thread).
template <typename F>
__device__ __forceinline__ void build_a(__nv_bfloat16* afrag, const float* src, int n, int ld, F f) {
  constexpr int kV = 8;                              // loads in flight per thread and round
  const int n4 = n >> 2, count = kRows * n4;
  for (int base = 0; base < count; base += kV * kCThreads) {
    float4 v[kV];
#pragma unroll
    for (int j = 0; j < kV; ++j) {
      const int i = base + threadIdx.x + j * kCThreads;
      if (i < count) {
        const int r = i / n4, k = (i - r * n4) << 2;
        v[j] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)r * ld + k));
      }
    }
#pragma unroll
    for (int j = 0; j < kV; ++j) {
      const int i = base + threadIdx.x + j * kCThreads;
      if (i < count) {
        const int r = i / n4, k = (i - r * n4) << 2;
        *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, k)) =
            __floats2bfloat162_rn(f(r, k, v[j].x), f(r, k + 1, v[j].y));
        *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, k + 2)) =
            __floats2bfloat162_rn(f(r, k + 2, v[j].z), f(r, k + 3, v[j].w));
      }
    }
  }
}

// A fragments (shared) of f(y[r][k]) from a [16][n] fp32 tile ALREADY staged in
// shared memory by TMA (one bulk-copy round trip instead of dependent L2 loads).
template <typename F>
__device__ __forceinline__ void convert_staged(__nv_bfloat16* afrag, const float* stage, int n, F f) {
  const int n4 = n >> 2;
  for (int i = threadIdx.x; i < kRows * n4; i += kCThreads) {
    const int r = i / n4, k = (i - r * n4) << 2;
    const float4 v = reinterpret_cast<const float4*>(stage)[i];
    *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, k)) =
        __floats2bfloat162_rn(f(r, k, v.x), f(r, k + 1, v.y));
    *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, k + 2)) =
        __floats2bfloat162_rn(f(r, k + 2, v.z), f(r, k + 3, v.w));
  }
}

// 8 consumer warps + the weight-producer warp + one operand-fetch warp
constexpr int kFwdThreads = kAllThreads + 32;

__global__ void __launch_bounds__(kFwdThreads, 1)
rssm_fwd_tma_kernel(const __grid_constant__ emb_rssm_fwd_args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int T = a.T, D = a.D, H = a.H, S = a.S, C = a.C, G = a.G;
  const int Dg = D / G, SC = S * C, Kh = Dg + 2 * H;
  const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x;
  const size_t RH = (size_t)kRows * H, RD = (size_t)kRows * D, RSC = (size_t)kRows * SC;

  // ---- shared memory: [barriers 256 B][segment table 256 B][out 16 x maxper*8 f32][stats][A region][ring]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* out = reinterpret_cast<float*>(smem_raw + 512);   // [256, 512): the producer's segment table
  const int maxper = (a.tma_cfg >> 8) & 0xff;
  float* rstd_a = out + kRows * maxper * 8;
  float* red = rstd_a + kRows;                                   // [kCWarps] + scratch
  int* sidx = reinterpret_cast<int*>(red + 32);                  // [S] sampled classes of a row
  // constants of this CTA's columns, staged once (the shared-memory carve-out leaves
  // no L1: every global read of a scale / bias would be an L2 round trip per use)
  float* c_shid = reinterpret_cast<float*>(sidx + 128);          // [Dg]  dynhid0norm scale, own group
  float* c_sobs = c_shid + Dg;                                   // [H]   obs0norm scale
  float* c_bgru = c_sobs + H;                                    // [3*Dg] dyngru bias, own group
  unsigned char* abase = reinterpret_cast<unsigned char*>(c_bgru + 3 * Dg);
  __nv_bfloat16* afrag = reinterpret_cast<__nv_bfloat16*>(abase);
  uint4* afrag4 = reinterpret_cast<uint4*>(abase);
  Ring ring;
  ring.nstages = a.tma_cfg & 0xff;
  ring.stage_bytes = ((a.tma_cfg >> 16) & 0xff) * 1024;
  ring.stage = 0;
  ring.phase = 0;
  ring.full = bars;
  ring.empty = bars + 12;
  uint64_t* afull01 = bars + 24;                                 // P4's A: deter slice + x0 landed
  uint64_t* afull2 = bars + 25;                                  // P4's A: x1 landed
  uint64_t* astage = bars + 26;                                  // fp32 tile of yhid / yobs landed
  ring.data = abase + emb_tma::a_region_bytes(a);

  if (tid == 0) {
    for (int i = 0; i < ring.nstages; ++i) { mbar_init(&ring.full[i], 1); mbar_init(&ring.empty[i], kCWarps); }
    mbar_init(afull01, 1);
    mbar_init(afull2, 1);
    mbar_init(astage, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int myrow = row_of_cta(a, cta);
  const bool rowcta = myrow >= 0;
  int wk = cta;                                    // (row CTAs keep their dense tiles in the fallback)
  if (rows_on_spare_ctas(a)) {
    wk = ncta;
    if (!rowcta) {
      wk = 0;
      for (int c = 0; c < cta; ++c) wk += row_of_cta(a, c) < 0;
    }
  }
  const Plan p = make_plan(a, wk);
  {
    const int g0 = p.on_gru ? p.u0_gru / (Dg / 8) : 0;
    for (int i = tid; i < Dg; i += kFwdThreads) c_shid[i] = a.s_hid[g0 * Dg + i];
    for (int i = tid; i < H; i += kFwdThreads) c_sobs[i] = a.s_obs[i];
    for (int i = tid; i < 3 * Dg; i += kFwdThreads) c_bgru[i] = a.b_gru[(size_t)g0 * 3 * Dg + i];
  }
  __syncthreads();
  __nv_bfloat16* deterA = reinterpret_cast<__nv_bfloat16*>(a.deterA);
  __nv_bfloat16* x0A = deterA + 2 * RD;                          // [16*H] fragments of x0
  __nv_bfloat16* x1A = x0A + RH;                                 // [16*H] fragments of x1
  const int ks01 = (Dg + H) / 16, ks2 = H / 16;                  // P4's two k ranges

  // =========================================================== producer warp
  // (the per-step schedule is static: see rssm_tma.cuh run_producer; the obs0 | dynin0 block
  //  of a CTA that only owns y0' columns is left out of the last step, like its consumers)
  Seg* segs = reinterpret_cast<Seg*>(smem_raw + 256);
  if (tid == kCThreads) {
    int n = 0;
    uint32_t skip_last = 0;
    if (p.on_hid) {
      segs[n++] = Seg{p.blk_hid, p.per_hid, ks01, 0};
      segs[n++] = Seg{p.blk_hid + (size_t)ks01 * p.per_hid * 256, p.per_hid, ks2, 0};
    }
    if (p.on_gru) segs[n++] = Seg{p.blk_gru, p.per_gru, p.ks_gru, 0};
    int u0, u1, l0, l1;
    ph1_range(a, p, false, u0, u1);
    ph1_range(a, p, true, l0, l1);
    if (u0 < u1) {
      if (!(l0 < l1)) skip_last |= 1u << n;
      segs[n++] = Seg{p.blk_ph1, p.per_ph1, p.ks_ph1, 1};
    }
    if (p.on_log) segs[n++] = Seg{p.blk_log, p.per_log, p.ks_log, 0};
    run_producer(ring, segs, n, T, skip_last, (a.tma_cfg >> 24) & 0x7f);
  }
  // Tenth warp (one lane): fetches P4's x1 fragments the moment the 16 row CTAs
  // have published them (release / acquire counter a.barrier[32], +16 per step) -- the
  // sampling phase of step t-1 runs underneath the first two k ranges of step t's dynhid0.
  if (tid == kAllThreads && p.on_hid) {
    const unsigned* flag = a.barrier + 32;
    const uint32_t xpart_ = (uint32_t)H * kRows * 2, part_ = (uint32_t)Dg * kRows * 2;
    for (int t = 0; t < T; ++t) {
      const unsigned want = (unsigned)kRows * (unsigned)(t + 1);
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      } while ((int)(v - want) < 0);
      fence_proxy_async();
      mbar_expect_tx(afull2, xpart_);
      bulk_g2s(abase + part_ + xpart_, x1A, xpart_, afull2);
    }
  }
  if (tid >= kCThreads) return;

  // ========================================================== consumer warps
  GridBarrierC bar{a.barrier, 0};
  uint32_t aphase = 0, sphase = 0;
  const int warp = tid >> 5, lane = tid & 31;
  // TMA staging of the fp32 tiles the dyngru / obslogit operands are built from:
  // fragments in the first 32 n bytes of the A region, the fp32 tile behind them
  const size_t aregion = emb_tma::a_region_bytes(a);
  const bool stage_hid = aregion >= (size_t)kRows * Dg * 6, stage_obs = aregion >= (size_t)kRows * H * 6;
  const int gh = p.on_hid ? p.u0_hid / (Dg / 8) : 0;
  const uint32_t part = (uint32_t)Dg * kRows * 2, xpart = (uint32_t)H * kRows * 2;

  // P4's A operand, part 1: [deter_{t-1} group slice | x0] (both final long before P4 starts)
  auto issue_a01 = [&](int t) {
    if (p.on_hid && tid == 0) {
      mbar_expect_tx(afull01, part + xpart);
      bulk_g2s(abase, deterA + (size_t)((t + 1) & 1) * RD + (size_t)gh * Dg * kRows, part, afull01);
      bulk_g2s(abase + part, x0A, xpart, afull01);
    }
  };
  // row `myrow` of a [16][H] global matrix -> registers
  auto load_row = [&](const float* yrow) {
    RowVals y;
#pragma unroll
    for (int gI = 0; gI < kRowGroups; ++gI) {
      const int c = gI * kCThreads * 4 + tid * 4;
      y.v[gI] = c < H ? __ldcg(reinterpret_cast<const float4*>(yrow + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return y;
  };

  // ---- prologue: deter0 -> fragments (slot 1 = "(t-1)&1" of step 0); x0, x1 of step 0
  for (size_t i = (size_t)cta * kCThreads + tid; i < RD; i += (size_t)ncta * kCThreads) {
    const int r = (int)(i / D), k = (int)(i - (size_t)r * D);
    deterA[RD + afrag_index(r, k)] = __float2bfloat16_rn(a.deter0[i]);
  }
  if (rowcta) {
    finish_row(load_row(a.y0 + (size_t)myrow * H), H, myrow, a.s0, a.eps, red, nullptr,
               a.rstd + myrow, x0A);
    finish_row(load_row(a.y1 + (size_t)myrow * H), H, myrow, a.s1, a.eps, red, nullptr,
               a.rstd + kRows + myrow, x1A);
  }
  // x1 fragments of the next step are final: publish (generic writes -> later TMA reads)
  auto publish_x1 = [&]() {
    fence_proxy_async();
    cbar();
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(a.barrier + 32) : "memory");
  };
  if (rowcta) publish_x1();
  bar.sync();
  issue_a01(0);

  // phase marks of CTA 0 (set 0) and of the CTA serving row 0 (set 1): timing[2][T][16]
#define MARK(i)                                                              \
  if (a.timing && (cta == 0 || myrow == 0) && tid == 0) {               \
    unsigned long long now_;                                                 \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_));                  \
    a.timing[((size_t)(cta ? T : 0) + t) * 16 + (i)] = now_;                 \
  }
#define MARKALL(i)                                                           \
  if (a.timing && t == 20 && tid == 0) {                                     \
    unsigned long long now_;                                                 \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_));                  \
    a.timing[(size_t)2 * T * 16 + cta * 4 + (i)] = now_;                     \
  }
  for (int t = 0; t < T; ++t) {
    MARK(0)
    const float* keep = a.keep + (size_t)t * kRows;
    const float* keep_next = a.keep + (size_t)(t + 1) * kRows;
    const float* deter_prev = t == 0 ? a.deter0 : a.deter + (size_t)(t - 1) * RD;
    float* yhid = a.yhid + (size_t)t * RD;
    const bool last = t + 1 == T;

    // ------------------------------------------------------------------ P4
    if (p.on_hid) {
      mbar_wait(afull01, aphase);   // (part 2 of A, x1, is fetched by the operand-fetch warp)
      // reset rows: deter_{t-1} enters as zero (rssm.py:76-77); rare, so patched in place
      float kp = tid < kRows ? ldcg(keep + tid) : 1.f;
      if (__syncthreads_or_consumers(kp == 0.f, red)) {
        for (int i = tid; i < (Dg / 16) * 32; i += kCThreads) {
          uint4 v = afrag4[i];
          const int r = (i & 31) >> 2;
          if (ldcg(keep + r) == 0.f) { v.x = 0; v.z = 0; }
          if (ldcg(keep + r + 8) == 0.f) { v.y = 0; v.w = 0; }
          afrag4[i] = v;
        }
        cbar();
      }
      MARK(1)
      MARKALL(0)
      EMB_CONSUME(false, ring, p.per_hid, ks01, afrag4, nullptr, out, true)
      MARKALL(1)
      mbar_wait(afull2, aphase);
      aphase ^= 1u;
      MARKALL(2)
      EMB_CONSUME(false, ring, p.per_hid, ks2, afrag4 + (size_t)ks01 * 32, nullptr, out, false)
      MARKALL(3)
      const int ncols = p.per_hid * 8, nvalid = (p.u1_hid - p.u0_hid) * 8;
      const float* pre = a.hid_pre + (size_t)t * RD;
      // epilogue: + hoisted action branch and bias -> yhid ; row sums of squares -> sumsq[t]
      float sq = 0.f;                                   // thread i: row i / 16, 16 threads per row
      {
        const int r = tid >> 4, c0 = tid & 15;
        for (int cb = c0; cb < nvalid; cb += 64) {             // four L2 loads in flight per thread
          float pv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = cb + 16 * e;
            pv[e] = c < nvalid ? __ldcg(pre + (size_t)r * D + p.u0_hid * 8 + c) : 0.f;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = cb + 16 * e;
            if (c < nvalid) {
              const float v = out[r * ncols + c] + pv[e];
              yhid[(size_t)r * D + p.u0_hid * 8 + c] = v;
              sq = fmaf(v, v, sq);
            }
          }
        }
#pragma unroll
        for (int o = 8; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (c0 == 0) atomicAdd(a.sumsq + (size_t)t * kRows + r, sq);
      }
    }
    MARK(2)
    bar.sync();
    MARK(3)

    // ------------------------------------------------------------------ P5
    float* deter = a.deter + (size_t)t * RD;
    if (p.on_gru) {
      const int upg = Dg / 8;
      const int g = p.u0_gru / upg;
      float* stage = reinterpret_cast<float*>(abase + (size_t)kRows * Dg * 2);
      if (stage_hid && tid == 0) {          // this group's slice of yhid: 16 row segments by TMA
        mbar_expect_tx(astage, (uint32_t)kRows * Dg * 4);
        for (int r = 0; r < kRows; ++r)
          bulk_g2s(stage + (size_t)r * Dg, yhid + (size_t)r * D + g * Dg, (uint32_t)Dg * 4, astage);
      }
      if (tid < kRows)
        rstd_a[tid] = rsqrtf(ldcg(a.sumsq + (size_t)t * kRows + tid) / (float)D + a.eps);
      cbar();
      if (stage_hid) {
        mbar_wait(astage, sphase);
        sphase ^= 1u;
        convert_staged(afrag, stage, Dg, [&](int r, int k, float v) {
          return silu_fast(v * (rstd_a[r] * c_shid[k])); });
      } else {
        build_a(afrag, yhid + g * Dg, Dg, D, [&](int r, int k, float v) {
          return silu_fast(v * (rstd_a[r] * c_shid[k])); });
      }
      cbar();
      MARK(12)
      const int ncols = p.per_gru * 8, nu = p.u1_gru - p.u0_gru;
      EMB_CONSUME(false, ring, p.per_gru, p.ks_gru, afrag4, nullptr, out, true)
      MARK(13)
      // epilogue: GRU gates (rssm.py:152-158); columns of a unit: [reset 8 | cand 8 | update 8].
      // Loads of all of a thread's elements first (L2 latency once), then the maths.
      constexpr int kE = 4;
      const int count = kRows * nu * 8;
      for (int base = 0; base < count; base += kCThreads * kE) {
        float bz[kE][3], oldv[kE], kp[kE];
#pragma unroll
        for (int e = 0; e < kE; ++e) {
          const int i = base + e * kCThreads + tid;
          if (i < count) {
            const int r = i / (nu * 8), c = i - r * (nu * 8);
            const int jj = (p.u0_gru - g * upg + (c >> 3)) * 8 + (c & 7);
            const float* bg = c_bgru + jj;
            bz[e][0] = bg[0]; bz[e][1] = bg[Dg]; bz[e][2] = bg[2 * Dg];
            oldv[e] = ldcg(deter_prev + (size_t)r * D + g * Dg + jj);
            kp[e] = ldcg(keep + r);
          }
        }
#pragma unroll
        for (int e = 0; e < kE; ++e) {
          const int i = base + e * kCThreads + tid;
          if (i < count) {
            const int r = i / (nu * 8), c = i - r * (nu * 8);
            const int u = c >> 3, nn = c & 7;
            const int col = g * Dg + (p.u0_gru - g * upg + u) * 8 + nn;
            const float* o = out + r * ncols + u * 24 + nn;
            const float rs = sigmoid_f(o[0] + bz[e][0]);
            const float cpre = o[8] + bz[e][1];
            const float cand = tanhf(rs * cpre);
            const float up = sigmoid_f(o[16] + bz[e][2] - 1.0f);
            const float nw = up * cand + (1.0f - up) * (kp[e] * oldv[e]);
            deter[(size_t)r * D + col] = nw;
            float* gs = a.gates + (size_t)t * 4 * RD + (size_t)r * D + col;
            gs[0] = rs; gs[RD] = cand; gs[2 * RD] = up; gs[3 * RD] = cpre;
            deterA[(size_t)(t & 1) * RD + afrag_index(r, col)] = __float2bfloat16_rn(nw);
          }
        }
      }
    }
    MARK(4)
    bar.sync();
    MARK(5)

    // ------------------------------------------------------------------ P1
    float* yobs = a.yobs + (size_t)t * RH;
    {
      int u0, u1;
      ph1_range(a, p, last, u0, u1);
      if (u0 < u1) {
        const uint4* dA = reinterpret_cast<const uint4*>(deterA + (size_t)(t & 1) * RD);
        const int ncols = p.per_ph1 * 8, nvalid = (u1 - u0) * 8;
        // thread i: row i / 16; yobs columns also feed the row's sum of squares (P2's norm)
        const int r = tid >> 4, c0 = tid & 15;
        EMB_CONSUME(true, ring, p.per_ph1, p.ks_ph1, dA, abase, out, true)
        float sq = 0.f;
        for (int c = c0; c < nvalid; c += 16) {
          const int col = u0 * 8 + c;
          if (col < H) {
            const float v = out[r * ncols + c] + ldcg(a.pre_tok + (size_t)t * RH + (size_t)r * H + col);
            yobs[(size_t)r * H + col] = v;
            sq = fmaf(v, v, sq);
          } else {
            a.y0[(size_t)(t + 1) * RH + (size_t)r * H + col - H] =
                ldcg(keep_next + r) * out[r * ncols + c] + a.b0[col - H];
          }
        }
#pragma unroll
        for (int o = 8; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (c0 == 0 && u0 * 8 < H) atomicAdd(a.sumsq_obs + (size_t)t * kRows + r, sq);
      }
    }
    MARK(6)
    bar.sync();
    MARK(7)

    // ------------------------------------------------------------------ P2
    float* logit = a.logit + (size_t)t * RSC;
    if (p.on_log) {
      float* stage = reinterpret_cast<float*>(abase + (size_t)kRows * H * 2);
      if (stage_obs && tid == 0) {          // yobs[t]: one contiguous tile by TMA
        mbar_expect_tx(astage, (uint32_t)kRows * H * 4);
        bulk_g2s(stage, yobs, (uint32_t)kRows * H * 4, astage);
      }
      if (tid < kRows) {
        const float rs = rsqrtf(ldcg(a.sumsq_obs + (size_t)t * kRows + tid) / (float)H + a.eps);
        rstd_a[tid] = rs;
        if (cta == 0) a.rstd[(size_t)t * 3 * kRows + 2 * kRows + tid] = rs;
      }
      cbar();
      if (stage_obs) {
        mbar_wait(astage, sphase);
        sphase ^= 1u;
        convert_staged(afrag, stage, H, [&](int r, int k, float v) {
          return silu_fast(v * (rstd_a[r] * c_sobs[k])); });
      } else {
        build_a(afrag, yobs, H, H, [&](int r, int k, float v) {
          return silu_fast(v * (rstd_a[r] * c_sobs[k])); });
      }
      cbar();
      MARK(14)
      EMB_CONSUME(false, ring, p.per_log, p.ks_log, afrag4, nullptr, out, true)
      MARK(15)
      const int ncols = p.per_log * 8, nvalid = (p.u1_log - p.u0_log) * 8;
      for (int i = tid; i < kRows * ncols; i += kCThreads) {
        const int r = i / ncols, c = i - r * ncols;
        if (c >= nvalid) continue;
        const int col = p.u0_log * 8 + c;
        logit[(size_t)r * SC + col] = out[i] + a.b_logit[col];
      }
    }
    if (!last && rowcta)
      finish_row(load_row(a.y0 + (size_t)(t + 1) * RH + (size_t)myrow * H), H, myrow, a.s0, a.eps, red,
                 nullptr, a.rstd + (size_t)(t + 1) * 3 * kRows + myrow, x0A);
    MARK(8)
    bar.sync();
    MARK(9)
    if (!last) issue_a01(t + 1);       // next step's [deter_t | x0'] are final: fetch them during P3

    // ------------------------------------------------------------------ P3
    // Row r is handled by one CTA: each warp samples four latents at a time
    // (softmax, unimix, Gumbel arg-max; all loads first), then the whole CTA sums
    // the S sampled dynin1 rows (the one-hot matmul of rssm.py:143 is a row
    // gather) in registers, norms, and leaves x1' as A fragments.
    if (rowcta) {
      const int r = myrow;
      const bool live = r < a.B;
      if (live) {
        // LP lanes per latent (<= kNC classes per lane): a warp samples 32 / LP latents at once
        // (size200m: 8 lanes x 8 classes, the 32 latents of a row in ONE round of the 8 warps)
        constexpr int kNC = 8;
        const int LP = C > 128 ? 32 : (C > 64 ? 16 : (C > 32 ? 8 : (C > 16 ? 4 : (C > 8 ? 2 : 1))));
        const int sub = lane / LP, ll = lane % LP, nsub = 32 / LP;
        const float* gum = a.gumbel + (size_t)t * RSC + (size_t)r * SC;
        const float* lrow = logit + (size_t)r * SC;
        // the row's logits and noise by TMA into the x1 slot of the A region (free
        // until the next dynhid0 phase): the sampling rounds then read shared memory
        const bool staged = aregion >= (size_t)part + xpart + (size_t)SC * 8;
        if (staged) {
          float* sl = reinterpret_cast<float*>(abase + part + xpart);
          if (tid == 0) {
            mbar_expect_tx(astage, (uint32_t)SC * 8);
            bulk_g2s(sl, lrow, (uint32_t)SC * 4, astage);
            bulk_g2s(sl + SC, gum, (uint32_t)SC * 4, astage);
          }
          mbar_wait(astage, sphase);
          sphase ^= 1u;
          lrow = sl;
          gum = sl + SC;
        }
#pragma unroll 1
        for (int sb = warp * nsub; sb < S; sb += kCWarps * nsub) {
          const int sv = sb + sub;
          const bool lat = sv < S;
          float lv[kNC], gv[kNC];
#pragma unroll
          for (int i = 0; i < kNC; ++i) {
            const int c = ll + LP * i;
            const bool on = lat && c < C;
            lv[i] = on ? (staged ? lrow[sv * C + c] : ldcg(lrow + (size_t)sv * C + c)) : -INFINITY;
            gv[i] = on ? (staged ? gum[sv * C + c] : ldcg(gum + (size_t)sv * C + c)) : 0.f;
          }
          float m = lv[0];
#pragma unroll
          for (int i = 1; i < kNC; ++i) m = fmaxf(m, lv[i]);
          for (int o = LP >> 1; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          float e[kNC], z = 0.f;
#pragma unroll
          for (int i = 0; i < kNC; ++i) { e[i] = (lat && ll + LP * i < C) ? __expf(lv[i] - m) : 0.f; z += e[i]; }
          for (int o = LP >> 1; o; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
          const float rz = 1.0f / z, um = 1.0f - a.unimix, uc = a.unimix / (float)C;
          float best = -INFINITY;
          int arg = 0x7fffffff;
#pragma unroll
          for (int i = 0; i < kNC; ++i) {
            const int c = ll + LP * i;
            if (lat && c < C) {
              const float pr = e[i] * rz;
              a.probs[(size_t)t * RSC + (size_t)r * SC + (size_t)sv * C + c] = pr;
              const float v = __logf(um * pr + uc) + gv[i];
              if (v > best || (v == best && c < arg)) { best = v; arg = c; }
            }
          }
          for (int o = LP >> 1; o; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
          }
          if (lat && ll == 0) {
            a.index[(size_t)t * kRows * S + (size_t)r * S + sv] = arg;
            sidx[sv] = arg;
          }
        }
      }
      cbar();
      if (!last) {
        // y1'[r] = b1 + keep' * sum_s dynin1[s*C + idx_s]
        const float kn = live ? ldcg(keep_next + r) : 0.f;
        const __nv_bfloat16* w1 = reinterpret_cast<const __nv_bfloat16*>(a.w_in1);
        RowVals y;
#pragma unroll
        for (int gI = 0; gI < kRowGroups; ++gI) {
          const int c = gI * kCThreads * 4 + tid * 4;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          if (c < H && kn != 0.f) {
#pragma unroll 32
            for (int sv = 0; sv < S; ++sv) {
              const uint2 q = __ldg(reinterpret_cast<const uint2*>(w1 + ((size_t)sv * C + sidx[sv]) * H + c));
              const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&q.x);
              const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
              s0 += __low2float(lo); s1 += __high2float(lo);
              s2 += __low2float(hi); s3 += __high2float(hi);
            }
          }
          if (c < H) {
            const float4 b = *reinterpret_cast<const float4*>(a.b1 + c);
            y.v[gI] = make_float4(b.x + kn * s0, b.y + kn * s1, b.z + kn * s2, b.w + kn * s3);
          } else {
            y.v[gI] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        finish_row(y, H, r, a.s1, a.eps, red, a.y1 + (size_t)(t + 1) * RH + (size_t)r * H,
                   a.rstd + (size_t)(t + 1) * 3 * kRows + kRows + r, x1A);
        publish_x1();
      }
    }
    MARK(10)
    MARK(11)       // no grid barrier here: dynhid0's last k range waits for the x1 counter instead
  }
#undef MARK
}

int g_sms_tma = 0;

}  // namespace

namespace emb_tma {

size_t fwd_smem_bytes(const emb_rssm_fwd_args& a, int maxper, int stage_bytes, int* nstages) {
  size_t fixed = 512 + sizeof(float) * (rssm::kRows * maxper * 8 + rssm::kRows + 32) + 128 * sizeof(int);
  fixed += sizeof(float) * (4 * (a.D / a.G) + a.H);             // staged constants
  fixed = (fixed + 127) & ~(size_t)127;
  fixed += a_region_bytes(a);
  const size_t cap = 227 * 1024 - 128;
  int n = fixed + 2 * (size_t)stage_bytes <= cap ? (int)((cap - fixed) / stage_bytes) : 0;
  if (n > 12) n = 12;
  if (*nstages > 0 && *nstages < n) n = *nstages;      // caller's cap (EMB_TMA_STAGES)
  *nstages = n;
  return fixed + (size_t)n * stage_bytes + 128;
}

int launch_fwd(const emb_rssm_fwd_args& a, void* stream, bool dry) {
  const char* who = "emb_rssm_observe_fwd";
  const int Dg = a.D / a.G;
  if (!a.hid_pre) return emb::fail(-1, "%s: the bf16 engine needs hid_pre (hoisted action branch)", who);
  if (g_sms_tma == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms_tma, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return emb::fail_cuda(who);
  }
  if (a.ncta < rssm::kRows || a.ncta > g_sms_tma)
    return emb::fail(-1, "%s: ncta=%d outside [16, %d SMs] (cooperative grid)", who, a.ncta, g_sms_tma);
  // tiles per CTA of every layer must fit the consumers' accumulators
  auto cdiv = [](int x, int y) { return (x + y - 1) / y; };
  const int cpg = a.ncta / a.G > 1 ? a.ncta / a.G : 1;
  const int per_hid = rssm_tma::pad_tiles(cdiv(Dg / 8, cpg), 1);
  const int per_gru = rssm_tma::pad_tiles(3 * cdiv(Dg / 8, cpg), 3);
  const int per_ph1 = rssm_tma::pad_tiles(cdiv(2 * a.H / 8, a.ncta), 1);
  const int per_log = rssm_tma::pad_tiles(cdiv(a.S * a.C / 8, a.ncta), 1);
  if (per_gru > rssm_tma::kMaxPer || per_hid > rssm_tma::kMaxPer || per_ph1 > rssm_tma::kMaxPer ||
      per_log > rssm_tma::kMaxPer)
    return emb::fail(-1, "%s: %d/%d/%d/%d tiles per CTA exceed %d (model too wide for %d CTAs)", who,
                     per_hid, per_gru, per_ph1, per_log, rssm_tma::kMaxPer, a.ncta);
  if (a.S > 128) return emb::fail(-1, "%s: stoch=%d > 128", who, a.S);
  if (a.C > 256) return emb::fail(-1, "%s: classes=%d > 256", who, a.C);
  if (a.H > 4096 || a.H % 4) return emb::fail(-1, "%s: hidden=%d must be <= 4096 and a multiple of 4", who, a.H);
  if (!a.sumsq_obs) return emb::fail(-1, "%s: the bf16 engine needs sumsq_obs", who);
  // tuning knobs (diagnostics): ring stage size in KiB and a cap on the stage count
  int stage_bytes = rssm_tma::kStageBytesDefault, stage_cap = 0;
  if (const char* e = getenv("EMB_TMA_STAGE_KB")) stage_bytes = atoi(e) * 1024;
  if (const char* e = getenv("EMB_TMA_STAGES")) stage_cap = atoi(e);
  if (stage_bytes < 8192 || stage_bytes > 65536 || stage_bytes % 1024)
    return emb::fail(-1, "%s: EMB_TMA_STAGE_KB out of range", who);
  // deter layer (A in global memory): every consumer warp needs >= 1 k16 step in every stage
  if ((a.D / 16) % (rssm_tma::kCWarps / rssm_tma::tile_groups(per_ph1)))
    return emb::fail(-1, "%s: D/16=%d must be a multiple of the %d k-lanes of the deter layer", who,
                     a.D / 16, rssm_tma::kCWarps / rssm_tma::tile_groups(per_ph1));
  emb_rssm_fwd_args copy = a;
  // `out` holds a layer's tile plus the slabs its k-lanes reduce through (rssm_tma.cuh out_tiles)
  int nstages = stage_cap, maxper = 0;
  for (int per : {per_hid, per_gru, per_ph1, per_log})
    if (rssm_tma::out_tiles(per) > maxper) maxper = rssm_tma::out_tiles(per);
  const size_t smem = fwd_smem_bytes(a, maxper, stage_bytes, &nstages);
  if (nstages < 2)
    return emb::fail(-1, "%s: A operand (%d columns) leaves no room for the weight ring", who, Dg + 2 * a.H);
  // L2 prefetch distance in ring chunks.  Measured (r02, size200m): 0 -> 61.5 us/step, 8 -> 68.2,
  // 16 -> 69.0, 32 -> 71.2: the stalls are latency chains, not a starved stream, and the
  // prefetched lines displace the activations the phases exchange through L2.  Off by default.
  int ahead = 0;
  if (const char* e = getenv("EMB_TMA_PREFETCH")) ahead = atoi(e);
  if (ahead < 0) ahead = 0;
  if (ahead > 127) ahead = 127;
  copy.tma_cfg = nstages | (maxper << 8) | ((stage_bytes / 1024) << 16) | (ahead << 24);
  if (dry) return 0;                      // emb_rssm_tma_fits: validation and sizing only
  const void* fn = (const void*)rssm_fwd_tma_kernel;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return emb::fail_cuda(who);
  void* params[] = {&copy};
  if (cudaLaunchCooperativeKernel(fn, dim3(a.ncta), dim3(kFwdThreads), params, smem,
                                  (cudaStream_t)stream) != cudaSuccess)
    return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}

}  // namespace emb_tma
