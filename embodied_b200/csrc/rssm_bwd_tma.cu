// emb_rssm_observe_bwd, bf16 engine: back-propagation through time of the fused
// RSSM scan (the reverse of rssm_fwd_tma.cu; dreamerv3/rssm.py:61-92,135-159
// differentiated) with the weight stream on the TMA ring of rssm_tma.cuh.
//
// Per step t = T-1 .. 0 the gradient is pushed back through the five in-scan
// layers with TRANSPOSED packed weights (one HBM pass over every weight per
// step, evict-first in L2); parameter gradients are formed afterwards by the
// host from the per-step layer gradients this kernel leaves behind
// (embodied_b200/dreamerv3/scan.py scan_backward).
//
//   B1  g_stoch = G_stoch[t] + (keep' * g_y1') @ dynin1^T                        (2.1 M)
//       every CTA owns WHOLE latents, so the unimix-softmax jacobian runs in its epilogue:
//       g_logit = G_logit[t] + (1-eps) p (g_stoch - sum_c p g_stoch), left as fp32 and as bf16
//       A fragments (in the g_stoch scratch buffer)
//   B2  g_xo = g_logit @ obslogit^T ; the operand is ONE TMA bulk copy of those fragments      (2.1 M)
//   B3  g_deter = G_deter[t] + carry + [g_yobs | keep' * g_y0'] @ [obs0[:D] | dynin0]^T    (16.8 M)
//       + GRU gate backward in the epilogue -> g_gates
//   B4  g_h     = g_gates_g @ dyngru[g]^T ; row dots for the rms-norm backward   (25.2 M)
//   B5  g_in    = g_yhid_g @ dynhid0[g][:Dg+2H]^T -> carry (deter part), g_x0 / g_x1 (25.2 M;
//       the action rows are hoisted: the host forms g_x2 with one GEMM)
// (primes = step t+1;  g_y* = rms-norm + silu backward of g_x*.)
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
// A region of the backward kernel: the widest operand (3Dg, 2H or SC columns of bf16
// fragments), B3's [U | g_y0' | Y] (6H bytes per row) and B5's [U | Y] (4Dg).
__host__ __device__ inline size_t bwd_a_region_bytes(int D, int G, int H, int SC) {
  const int Dg = D / G;
  size_t n = (size_t)rssm::kRows * 3 * Dg * 2;
  const size_t cands[4] = {(size_t)rssm::kRows * 2 * H * 2, (size_t)rssm::kRows * SC * 2,
                           (size_t)rssm::kRows * H * 6, (size_t)rssm::kRows * Dg * 4};
  for (int i = 0; i < 4; ++i) if (cands[i] > n) n = cands[i];
  return (n + 127) & ~(size_t)127;
}
}  // namespace emb_tma

namespace {

using namespace rssm;
using namespace rssm_tma;

__device__ __forceinline__ float dsilu_fast(float n) {          // d silu(n) / dn
  const float sg = __fdividef(1.0f, 1.0f + __expf(-n));
  return sg * (1.0f + n * (1.0f - sg));
}
// rms-norm + silu backward (nets.py:374-383): with n = y * rstd * s,
//   g_n = g_x * silu'(n),   g_y = rstd * s * g_n - y * coef,   coef = rstd^3 * mean_k(g_n s y)
__device__ __forceinline__ float norm_bwd_elem(float gx, float y, float s, float rstd, float coef) {
  const float gn = gx * dsilu_fast(y * rstd * s);
  return rstd * s * gn - y * coef;
}

// A fragments (shared) from two fp32 [16][n] global sources: f(r, k, v1, v2)
template <typename F>
__device__ __forceinline__ void build2(__nv_bfloat16* afrag, int koff, const float* s1, int ld1,
                                       const float* s2, int ld2, int n, F f) {
  const int n4 = n >> 2;
#pragma unroll 4
  for (int i = threadIdx.x; i < kRows * n4; i += kCThreads) {
    const int r = i / n4, k = (i - r * n4) << 2;
    const float4 v = __ldcg(reinterpret_cast<const float4*>(s1 + (size_t)r * ld1 + k));
    const float4 w = __ldcg(reinterpret_cast<const float4*>(s2 + (size_t)r * ld2 + k));
    *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, koff + k)) =
        __floats2bfloat162_rn(f(r, k, v.x, w.x), f(r, k + 1, v.y, w.y));
    *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, koff + k + 2)) =
        __floats2bfloat162_rn(f(r, k + 2, v.z, w.z), f(r, k + 3, v.w, w.w));
  }
}
template <typename F>
__device__ __forceinline__ void build1(__nv_bfloat16* afrag, int koff, const float* src, int n, int ld, F f) {
  const int n4 = n >> 2;
#pragma unroll 4
  for (int i = threadIdx.x; i < kRows * n4; i += kCThreads) {
    const int r = i / n4, k = (i - r * n4) << 2;
    const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (size_t)r * ld + k));
    *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, koff + k)) =
        __floats2bfloat162_rn(f(r, k, v.x), f(r, k + 1, v.y));
    *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, koff + k + 2)) =
        __floats2bfloat162_rn(f(r, k + 2, v.z), f(r, k + 3, v.w));
  }
}

// A = U - coef[row] * Y on `n16` 16-byte fragment vectors in shared memory (in place in
// U).  Fragment layout (rssm_common.cuh afrag_index): within a k16 block lane = vec % 32,
// the four 32-bit registers hold rows (lane/4, lane/4 + 8, lane/4, lane/4 + 8).
__device__ __forceinline__ void combine_frags(uint4* u, const uint4* y, int n16, const float* coef) {
  for (int i = threadIdx.x; i < n16; i += kCThreads) {
    const int r0 = (i & 31) >> 2;
    const float c0 = coef[r0], c1 = coef[r0 + 8];
    uint4 uv = u[i];
    const uint4 yv = y[i];
    uint32_t* up = reinterpret_cast<uint32_t*>(&uv);
    const uint32_t* yp = reinterpret_cast<const uint32_t*>(&yv);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float c = (q & 1) ? c1 : c0;
      const float2 uf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&up[q]));
      const float2 yf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yp[q]));
      const __nv_bfloat162 o = __floats2bfloat162_rn(uf.x - c * yf.x, uf.y - c * yf.y);
      up[q] = *reinterpret_cast<const uint32_t*>(&o);
    }
    u[i] = uv;
  }
}

constexpr int kRowGroups = 4;         // a row CTA holds H <= 4096 columns, 4 per thread and group

struct Plan {
  int per[5], ks[5], u0[5], u1[5];        // B1..B5: padded tiles per CTA, k16 steps, owned tiles
  const unsigned char* blk[5];
};

__device__ __forceinline__ Plan make_plan(const emb_rssm_bwd_args& a) {
  Plan p;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int D = a.D, H = a.H, Dg = D / a.G, SC = a.S * a.C, Kh = Dg + 2 * H;
  const int tiles[3] = {SC / 8, H / 8, D / 8};
  const int kdim[5] = {H, SC, 2 * H, 3 * Dg, Dg};
  const void* w[5] = {a.wt_in1, a.wt_logit, a.wt_ph1, a.wt_gru, a.wt_hid};
  int unit1 = a.C;                                  // tiles per B1 unit = lcm(C, 8) / 8: whole latents
  while (unit1 % 8) unit1 += a.C;
  unit1 /= 8;
  for (int i = 0; i < 3; ++i) {
    const int unit = i == 0 ? unit1 : 1;
    const int raw = (tiles[i] / unit + ncta - 1) / ncta * unit;
    p.per[i] = pad_tiles(raw, unit);
    p.u0[i] = min(tiles[i], cta * raw);
    p.u1[i] = min(tiles[i], p.u0[i] + raw);
  }
  const GroupSplit s4 = group_split(Dg / 8, a.G), s5 = group_split(Kh / 8, a.G);
  p.per[3] = pad_tiles(s4.per, 1); p.u0[3] = s4.u0; p.u1[3] = s4.u1;
  p.per[4] = pad_tiles(s5.per, 1); p.u0[4] = s5.u0; p.u1[4] = s5.u1;
  for (int i = 0; i < 5; ++i) {
    p.ks[i] = kdim[i] / 16;
    p.blk[i] = reinterpret_cast<const unsigned char*>(w[i]) + (size_t)cta * p.ks[i] * p.per[i] * 256;
  }
  return p;
}

__global__ void __launch_bounds__(kAllThreads, 1)
rssm_bwd_tma_kernel(const __grid_constant__ emb_rssm_bwd_args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int T = a.T, D = a.D, H = a.H, S = a.S, C = a.C, G = a.G;
  const int Dg = D / G, SC = S * C, Kh = Dg + 2 * H;
  const int tid = threadIdx.x;
  const size_t RH = (size_t)kRows * H, RD = (size_t)kRows * D, RSC = (size_t)kRows * SC;

  // ---- shared memory: [barriers 256 B][segment table 256 B][out 16 x maxper*8 f32][4 x 16 row statistics][A region][ring]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* out = reinterpret_cast<float*>(smem_raw + 512);   // [256, 512): the producer's segment table
  const int maxper = (a.hoist_x2 >> 8) & 0xff;
  float* st = out + kRows * maxper * 8;
  float *rstd_a = st, *coef_a = st + 16, *rstd_b = st + 32, *coef_b = st + 48;
  unsigned char* abase = reinterpret_cast<unsigned char*>(st + 64);
  __nv_bfloat16* afrag = reinterpret_cast<__nv_bfloat16*>(abase);
  uint4* afrag4 = reinterpret_cast<uint4*>(abase);
  Ring ring;
  ring.nstages = a.hoist_x2 & 0xff;
  ring.stage_bytes = ((a.hoist_x2 >> 16) & 0xff) * 1024;
  ring.stage = 0;
  ring.phase = 0;
  ring.full = bars;
  ring.empty = bars + 12;
  uint64_t* astage = bars + 24;                       // B2's operand (fragments of g_logit) landed
  ring.data = abase + emb_tma::bwd_a_region_bytes(D, G, H, SC);
  if (tid == 0) {
    for (int i = 0; i < ring.nstages; ++i) { mbar_init(&ring.full[i], 1); mbar_init(&ring.empty[i], kCWarps); }
    mbar_init(astage, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const Plan p = make_plan(a);

  // =========================================================== producer warp
  Seg* segs = reinterpret_cast<Seg*>(smem_raw + 256);
  if (tid == kCThreads) {
    int n = 0;
    for (int i = 0; i < 5; ++i)
      if (p.u0[i] < p.u1[i]) segs[n++] = Seg{p.blk[i], p.per[i], p.ks[i], 0};
    run_producer(ring, segs, n, T, 0u, (a.hoist_x2 >> 24) & 0x7f);
  }
  if (tid >= kCThreads) return;

  // ========================================================== consumer warps
  GridBarrierC bar{a.barrier, 0};
  uint32_t sphase = 0;
  __nv_bfloat16* glA = reinterpret_cast<__nv_bfloat16*>(a.g_stoch);   // [SC/16][32][8] fragments of g_logit
  // bf16 fragment hand-offs between phases (a.frag_scratch): producers' epilogues leave the
  // next phase's A operand ready, consumers fetch it with TMA bulk copies.  For a normalised
  // layer the operand is g_y = U - coef * Y (U = rstd s g_n, Y = y): U and Y are written
  // before the row dots (coef) are complete and combined in shared memory afterwards.
  __nv_bfloat16* ggA = reinterpret_cast<__nv_bfloat16*>(a.frag_scratch);   // [G][3Dg] x 16 rows: g_gates
  __nv_bfloat16* UhA = ggA + 3 * RD;                                       // [D]   dynhid0 norm: U
  __nv_bfloat16* YhA = UhA + RD;                                           // [D]   dynhid0 norm: Y
  __nv_bfloat16* UoA = YhA + RD;                                           // [H]   obs0 norm: U
  __nv_bfloat16* YoA = UoA + RH;                                           // [H]   obs0 norm: Y
  __nv_bfloat16* gy0A = YoA + RH;                                          // [H]   keep' * g_y0'
  __nv_bfloat16* gy1A = gy0A + RH;                                         // [H]   keep' * g_y1'
  float* gxp = a.gx_part;                            // [G][16][2H] per-group partials of g_[x0|x1]
  unsigned* row_flag = a.barrier + 1;                // rows whose g_y0' / g_y1' fragments are ready
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int myrow = ncta - 1 - cta;                  // rows 0..15 are served by the LAST 16 CTAs

  // row statistics of a normalised layer at step `ts`: slot 0 x0, 1 x1, 2 xo
  auto load_stats = [&](int ts, int slot, int n, float* rstd, float* coef) {
    if (tid < kRows) {
      const float rs = a.rstd[(size_t)ts * 3 * kRows + slot * kRows + tid];
      rstd[tid] = rs;
      coef[tid] = rs * rs * rs * ldcg(a.dots + ((size_t)ts * 4 + slot) * kRows + tid) / (float)n;
    }
  };
  // sum out[r][c0..c1) per row (16 threads per row) and add it to a.dots[ts][slot][r]
  auto add_row_dots = [&](int ts, int slot, int ncols, int c0, int c1) {
    const int r = tid >> 4, l = tid & 15;
    float sum = 0.f;
    for (int c = c0 + l; c < c1; c += 16) sum += out[r * ncols + c];
#pragma unroll
    for (int o = 8; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (l == 0 && c0 < c1) atomicAdd(a.dots + ((size_t)ts * 4 + slot) * kRows + r, sum);
  };

  // Row r of the gradients that came back through dynhid0 at step `ts`: sum the per-group
  // partials (no atomics), store g_x0 / g_x1 [ts] for the host, and -- with the complete row
  // in registers, so the row dots need no cross-CTA reduction either -- leave keep * g_y0 and
  // keep * g_y1 (rms-norm + silu backward) as operand fragments for B3 and B1.
  auto row_work = [&](int ts, bool frags) {
    const int r = myrow;
    const float* part0 = gxp + (size_t)r * 2 * H;
    const size_t gstride = (size_t)kRows * 2 * H;
    const float rs0 = a.rstd[(size_t)ts * 3 * kRows + r], rs1 = a.rstd[(size_t)ts * 3 * kRows + kRows + r];
    const float kn = ldcg(a.keep + (size_t)ts * kRows + r);
    float d0 = 0.f, d1 = 0.f;
    float4 gx0[kRowGroups], gx1[kRowGroups], y0v[kRowGroups], y1v[kRowGroups];
#pragma unroll
    for (int q = 0; q < kRowGroups; ++q) {
      const int k = q * kCThreads * 4 + tid * 4;
      gx0[q] = gx1[q] = y0v[q] = y1v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < H) {
        for (int g = 0; g < G; ++g) {
          const float4 u = __ldcg(reinterpret_cast<const float4*>(part0 + g * gstride + k));
          const float4 v = __ldcg(reinterpret_cast<const float4*>(part0 + g * gstride + H + k));
          gx0[q].x += u.x; gx0[q].y += u.y; gx0[q].z += u.z; gx0[q].w += u.w;
          gx1[q].x += v.x; gx1[q].y += v.y; gx1[q].z += v.z; gx1[q].w += v.w;
        }
        *reinterpret_cast<float4*>(a.g_x0 + (size_t)ts * RH + (size_t)r * H + k) = gx0[q];
        *reinterpret_cast<float4*>(a.g_x1 + (size_t)ts * RH + (size_t)r * H + k) = gx1[q];
        if (frags) {
          y0v[q] = *reinterpret_cast<const float4*>(a.y0 + (size_t)ts * RH + (size_t)r * H + k);
          y1v[q] = *reinterpret_cast<const float4*>(a.y1 + (size_t)ts * RH + (size_t)r * H + k);
          const float4 s0 = *reinterpret_cast<const float4*>(a.s0 + k);
          const float4 s1 = *reinterpret_cast<const float4*>(a.s1 + k);
          // g_n * s * y, g_n = g_x * silu'(y * rstd * s)
          d0 += gx0[q].x * dsilu_fast(y0v[q].x * rs0 * s0.x) * s0.x * y0v[q].x +
                gx0[q].y * dsilu_fast(y0v[q].y * rs0 * s0.y) * s0.y * y0v[q].y +
                gx0[q].z * dsilu_fast(y0v[q].z * rs0 * s0.z) * s0.z * y0v[q].z +
                gx0[q].w * dsilu_fast(y0v[q].w * rs0 * s0.w) * s0.w * y0v[q].w;
          d1 += gx1[q].x * dsilu_fast(y1v[q].x * rs1 * s1.x) * s1.x * y1v[q].x +
                gx1[q].y * dsilu_fast(y1v[q].y * rs1 * s1.y) * s1.y * y1v[q].y +
                gx1[q].z * dsilu_fast(y1v[q].z * rs1 * s1.z) * s1.z * y1v[q].z +
                gx1[q].w * dsilu_fast(y1v[q].w * rs1 * s1.w) * s1.w * y1v[q].w;
        }
      }
    }
    if (!frags) return;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      d0 += __shfl_xor_sync(0xffffffffu, d0, o);
      d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    }
    cbar();
    if ((tid & 31) == 0) { st[tid >> 5] = d0; st[8 + (tid >> 5)] = d1; }     // rstd_a / coef_a slots
    cbar();
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int w = 0; w < kCWarps; ++w) { t0 += st[w]; t1 += st[8 + w]; }
    const float c0 = rs0 * rs0 * rs0 * t0 / (float)H, c1 = rs1 * rs1 * rs1 * t1 / (float)H;
#pragma unroll
    for (int q = 0; q < kRowGroups; ++q) {
      const int k = q * kCThreads * 4 + tid * 4;
      if (k < H) {
        const float4 s0 = *reinterpret_cast<const float4*>(a.s0 + k);
        const float4 s1 = *reinterpret_cast<const float4*>(a.s1 + k);
        *reinterpret_cast<__nv_bfloat162*>(gy0A + afrag_index(r, k)) = __floats2bfloat162_rn(
            kn * norm_bwd_elem(gx0[q].x, y0v[q].x, s0.x, rs0, c0), kn * norm_bwd_elem(gx0[q].y, y0v[q].y, s0.y, rs0, c0));
        *reinterpret_cast<__nv_bfloat162*>(gy0A + afrag_index(r, k + 2)) = __floats2bfloat162_rn(
            kn * norm_bwd_elem(gx0[q].z, y0v[q].z, s0.z, rs0, c0), kn * norm_bwd_elem(gx0[q].w, y0v[q].w, s0.w, rs0, c0));
        *reinterpret_cast<__nv_bfloat162*>(gy1A + afrag_index(r, k)) = __floats2bfloat162_rn(
            kn * norm_bwd_elem(gx1[q].x, y1v[q].x, s1.x, rs1, c1), kn * norm_bwd_elem(gx1[q].y, y1v[q].y, s1.y, rs1, c1));
        *reinterpret_cast<__nv_bfloat162*>(gy1A + afrag_index(r, k + 2)) = __floats2bfloat162_rn(
            kn * norm_bwd_elem(gx1[q].z, y1v[q].z, s1.z, rs1, c1), kn * norm_bwd_elem(gx1[q].w, y1v[q].w, s1.w, rs1, c1));
      }
    }
    fence_proxy_async();
    cbar();
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(row_flag) : "memory");
  };

#define MARK(i)                                                              \
  if (a.timing && blockIdx.x == 0 && tid == 0) {                             \
    unsigned long long now_;                                                 \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_));                  \
    a.timing[(size_t)t * 16 + (i)] = now_;                                   \
  }
  for (int t = T - 1; t >= 0; --t) {
    MARK(0)
    const float* keep = a.keep + (size_t)t * kRows;
    const float* keep_next = a.keep + (size_t)(t + 1) * kRows;
    const float* deter_prev = t == 0 ? a.deter0 : a.deter + (size_t)(t - 1) * RD;
    const float* y0 = a.y0 + (size_t)t * RH;
    const float* y1 = a.y1 + (size_t)t * RH;
    const float* y0n = a.y0 + (size_t)(t + 1) * RH;         // step t+1 (zeros at t = T-1)
    const float* y1n = a.y1 + (size_t)(t + 1) * RH;
    const float* gx0n = a.g_x0 + (size_t)(t + 1) * RH;
    const float* gx1n = a.g_x1 + (size_t)(t + 1) * RH;
    const float* yobs = a.yobs + (size_t)t * RH;
    const float* yhid = a.yhid + (size_t)t * RD;
    const float* gates = a.gates + (size_t)t * 4 * RD;
    float* g_xo = a.g_xo + (size_t)t * RH;
    float* g_h = a.g_h + (size_t)t * RD;
    float* g_gates = a.g_gates + (size_t)t * 3 * RD;
    float* g_logit = a.g_logit + (size_t)t * RSC;

    // ------------------------------------------------------------------ B1
    // 16 otherwise idle CTAs first finish step t+1's gradients wrt x0 / x1 (row_work); the
    // CTAs owning latents wait for the 16 rows (release / acquire counter) and fetch their
    // operand keep' * g_y1' by TMA.
    if (myrow >= 0 && myrow < kRows) row_work(t + 1, true);
    if (p.u0[0] < p.u1[0]) {
      // epilogue inputs first: their L2 latency hides behind the flag wait and the GEMM
      const int ncols = p.per[0] * 8, nvalid = (p.u1[0] - p.u0[0]) * 8;
      const int r = tid >> 4, l = tid & 15;
      const float* pr = a.probs + (size_t)t * RSC + (size_t)r * SC;
      const float* Gs = a.G_stoch + (size_t)t * RSC + (size_t)r * SC;
      const float* Gl = a.G_logit + (size_t)t * RSC + (size_t)r * SC;
      constexpr int kPre = 8;                      // C <= 128: 8 classes per thread and latent
      float ppre[kPre], gspre[kPre], glpre[kPre];
      const bool one_latent = nvalid == C;
      if (one_latent) {
#pragma unroll
        for (int i = 0; i < kPre; ++i) {
          const int c = l + 16 * i, col = p.u0[0] * 8 + c;
          ppre[i] = c < C ? pr[col] : 0.f;
          gspre[i] = c < C ? Gs[col] : 0.f;
          glpre[i] = c < C ? Gl[col] : 0.f;
        }
      }
      if (tid == 0) {
        const unsigned target = (unsigned)kRows * (unsigned)(T - t);
        unsigned v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(row_flag) : "memory");
        } while ((int)(v - target) < 0);
        fence_proxy_async();
        mbar_expect_tx(astage, (uint32_t)RH * 2);
        bulk_g2s(abase, gy1A, (uint32_t)RH * 2, astage);
      }
      mbar_wait(astage, sphase);
      sphase ^= 1u;
      EMB_CONSUME(false, ring, p.per[0], p.ks[0], afrag4, nullptr, out, true)
      // epilogue: g_stoch -> g_logit for the owned latents (16 threads per row)
      const float um = 1.0f - a.unimix;
      if (one_latent) {
        const int col0 = p.u0[0] * 8;
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < kPre; ++i) {
          const int c = l + 16 * i;
          if (c < C) { gspre[i] += out[r * ncols + c]; dot = fmaf(ppre[i], gspre[i], dot); }
        }
#pragma unroll
        for (int o = 8; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
        for (int i = 0; i < kPre; ++i) {
          const int c = l + 16 * i;
          if (c < C) {
            const float v = glpre[i] + um * ppre[i] * (gspre[i] - dot);
            g_logit[(size_t)r * SC + col0 + c] = v;
            glA[afrag_index(r, col0 + c)] = __float2bfloat16_rn(v);
          }
        }
      } else {
        for (int c0 = 0; c0 < nvalid; c0 += C) {         // one latent at a time
          const int col0 = p.u0[0] * 8 + c0;
          float dot = 0.f;
          for (int c = l; c < C; c += 16)
            dot = fmaf(pr[col0 + c], out[r * ncols + c0 + c] + Gs[col0 + c], dot);
#pragma unroll
          for (int o = 8; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
          for (int c = l; c < C; c += 16) {
            const float g = out[r * ncols + c0 + c] + Gs[col0 + c];
            const float v = Gl[col0 + c] + um * pr[col0 + c] * (g - dot);
            g_logit[(size_t)r * SC + col0 + c] = v;
            glA[afrag_index(r, col0 + c)] = __float2bfloat16_rn(v);
          }
        }
      }
    }
    MARK(1)
    bar.sync();
    MARK(2)

    // ------------------------------------------------------------------ B2
    // g_xo = g_logit @ obslogit^T ; row dots of the obs0 norm backward
    if (p.u0[1] < p.u1[1]) {
      if (tid == 0) {
        mbar_expect_tx(astage, (uint32_t)kRows * SC * 2);
        bulk_g2s(abase, glA, (uint32_t)kRows * SC * 2, astage);
      }
      load_stats(t, 2, H, rstd_a, coef_a);       // only rstd is used here
      cbar();
      mbar_wait(astage, sphase);
      sphase ^= 1u;
      EMB_CONSUME(false, ring, p.per[1], p.ks[1], afrag4, nullptr, out, true)
      const int ncols = p.per[1] * 8, nvalid = (p.u1[1] - p.u0[1]) * 8;
      for (int i = tid; i < kRows * ncols; i += kCThreads) {
        const int r = i / ncols, c = i - r * ncols;
        float prod = 0.f;
        if (c < nvalid) {
          const int col = p.u0[1] * 8 + c;
          const float gx = out[i];
          g_xo[(size_t)r * H + col] = gx;
          const float y = yobs[(size_t)r * H + col], sc = a.s_obs[col];
          const float gn = gx * dsilu_fast(y * rstd_a[r] * sc);
          prod = gn * sc * y;                                       // g_n * s * y
          UoA[afrag_index(r, col)] = __float2bfloat16_rn(rstd_a[r] * sc * gn);
          YoA[afrag_index(r, col)] = __float2bfloat16_rn(y);
        }
        out[i] = prod;
      }
      cbar();
      add_row_dots(t, 2, ncols, 0, nvalid);
    }
    MARK(3)
    bar.sync();
    MARK(4)

    // ------------------------------------------------------------------ B3
    if (p.u0[2] < p.u1[2]) {
      // operand [g_yobs | keep' g_y0']: U_obs, keep' g_y0' and Y_obs by TMA (Y behind the 2H columns)
      if (tid == 0) {
        mbar_expect_tx(astage, (uint32_t)RH * 2 * 3);
        bulk_g2s(abase, UoA, (uint32_t)RH * 2, astage);
        bulk_g2s(abase + RH * 2, gy0A, (uint32_t)RH * 2, astage);
        bulk_g2s(abase + RH * 4, YoA, (uint32_t)RH * 2, astage);
      }
      load_stats(t, 2, H, rstd_a, coef_a);
      cbar();
      mbar_wait(astage, sphase);
      sphase ^= 1u;
      combine_frags(afrag4, reinterpret_cast<const uint4*>(abase + RH * 4), (H / 16) * 32, coef_a);
      cbar();
      EMB_CONSUME(false, ring, p.per[2], p.ks[2], afrag4, nullptr, out, true)
      // epilogue: GRU backward (rssm.py:152-158) for this CTA's columns.  Compact element
      // index over the valid columns; the eight inputs of kE elements are requested before any
      // is used (one L2 round trip per batch).
      const int ncols = p.per[2] * 8, nvalid = (p.u1[2] - p.u0[2]) * 8;
      const int count = kRows * nvalid;
      constexpr int kE = 4;
      for (int base = 0; base < count; base += kCThreads * kE) {
        float in[kE][8];
#pragma unroll
        for (int e = 0; e < kE; ++e) {
          const int i = base + e * kCThreads + tid;
          if (i < count) {
            const int r = i / nvalid, c = i - r * nvalid;
            const size_t at = (size_t)r * D + p.u0[2] * 8 + c;
            in[e][0] = a.G_deter[(size_t)t * RD + at];
            in[e][1] = ldcg(a.gd_carry + at);
            in[e][2] = gates[at]; in[e][3] = gates[RD + at]; in[e][4] = gates[2 * RD + at];
            in[e][5] = gates[3 * RD + at];
            in[e][6] = deter_prev[at];
            in[e][7] = ldcg(keep + r);
          }
        }
#pragma unroll
        for (int e = 0; e < kE; ++e) {
          const int i = base + e * kCThreads + tid;
          if (i < count) {
            const int r = i / nvalid, c = i - r * nvalid;
            const int col = p.u0[2] * 8 + c;
            const size_t at = (size_t)r * D + col;
            const float gd = out[r * ncols + c] + in[e][0] + in[e][1];
            const float rs = in[e][2], cand = in[e][3], up = in[e][4], cpre = in[e][5];
            const float old = in[e][7] * in[e][6];
            const float g_u = gd * (cand - old), g_c = gd * up;
            a.gd_tmp[at] = gd * (1.0f - up);                 // direct path into keep*deter_{t-1}
            const float g_rc = g_c * (1.0f - cand * cand);
            const int g = col / Dg, jj = col - g * Dg;
            float* gg = g_gates + (size_t)r * 3 * D + (size_t)g * 3 * Dg + jj;
            const float g0 = g_rc * cpre * rs * (1.0f - rs);  // reset gate, pre-sigmoid
            const float g1 = g_rc * rs;                       // candidate, pre-tanh
            const float g2 = g_u * up * (1.0f - up);          // update gate, pre-sigmoid
            gg[0] = g0; gg[Dg] = g1; gg[2 * Dg] = g2;
            __nv_bfloat16* ga = ggA + (size_t)g * 3 * Dg * kRows;         // B4's operand, group g
            ga[afrag_index(r, jj)] = __float2bfloat16_rn(g0);
            ga[afrag_index(r, Dg + jj)] = __float2bfloat16_rn(g1);
            ga[afrag_index(r, 2 * Dg + jj)] = __float2bfloat16_rn(g2);
          }
        }
      }
    }
    MARK(5)
    bar.sync();
    MARK(6)

    // ------------------------------------------------------------------ B4
    if (p.u0[3] < p.u1[3]) {
      const int tpg = Dg / 8;
      const int g = p.u0[3] / tpg;
      if (tid == 0) {
        mbar_expect_tx(astage, (uint32_t)kRows * 3 * Dg * 2);
        bulk_g2s(abase, ggA + (size_t)g * 3 * Dg * kRows, (uint32_t)kRows * 3 * Dg * 2, astage);
      }
      if (tid < kRows) rstd_a[tid] = rsqrtf(a.sumsq[(size_t)t * kRows + tid] / (float)D + a.eps);
      cbar();
      mbar_wait(astage, sphase);
      sphase ^= 1u;
      EMB_CONSUME(false, ring, p.per[3], p.ks[3], afrag4, nullptr, out, true)
      const int ncols = p.per[3] * 8, nvalid = (p.u1[3] - p.u0[3]) * 8;
      {
        constexpr int kE = 4;
        const int total = kRows * ncols;
        for (int base = 0; base < total; base += kCThreads * kE) {
          float yv[kE], sv[kE];
#pragma unroll
          for (int e = 0; e < kE; ++e) {
            const int i = base + e * kCThreads + tid;
            const int r = i / ncols, c = i - r * ncols;
            const bool on = i < total && c < nvalid;
            yv[e] = on ? yhid[(size_t)r * D + p.u0[3] * 8 + c] : 0.f;
            sv[e] = on ? a.s_hid[p.u0[3] * 8 + c] : 0.f;
          }
#pragma unroll
          for (int e = 0; e < kE; ++e) {
            const int i = base + e * kCThreads + tid;
            if (i >= total) continue;
            const int r = i / ncols, c = i - r * ncols;
            float prod = 0.f;
            if (c < nvalid) {
              const int col = p.u0[3] * 8 + c;
              const size_t at = (size_t)r * D + col;
              const float gh = out[i];
              g_h[at] = gh;
              const float gn = gh * dsilu_fast(yv[e] * rstd_a[r] * sv[e]);
              prod = gn * sv[e] * yv[e];                              // g_n * s * y
              UhA[afrag_index(r, col)] = __float2bfloat16_rn(rstd_a[r] * sv[e] * gn);
              YhA[afrag_index(r, col)] = __float2bfloat16_rn(yv[e]);
            }
            out[i] = prod;
          }
        }
      }
      cbar();
      add_row_dots(t, 3, ncols, 0, nvalid);
    }
    MARK(7)
    bar.sync();
    MARK(8)

    // ------------------------------------------------------------------ B5
    if (p.u0[4] < p.u1[4]) {
      const int tpg = Kh / 8;
      const int g = p.u0[4] / tpg;
      if (tid < kRows) {
        const float rs = rsqrtf(a.sumsq[(size_t)t * kRows + tid] / (float)D + a.eps);
        rstd_a[tid] = rs;
        coef_a[tid] = rs * rs * rs * ldcg(a.dots + ((size_t)t * 4 + 3) * kRows + tid) / (float)D;
      }
      if (tid == 0) {       // this group's slices of U and Y (fragment order: contiguous)
        mbar_expect_tx(astage, (uint32_t)kRows * Dg * 2 * 2);
        bulk_g2s(abase, UhA + (size_t)g * Dg * kRows, (uint32_t)kRows * Dg * 2, astage);
        bulk_g2s(abase + (size_t)kRows * Dg * 2, YhA + (size_t)g * Dg * kRows, (uint32_t)kRows * Dg * 2, astage);
      }
      cbar();
      mbar_wait(astage, sphase);
      sphase ^= 1u;
      combine_frags(afrag4, reinterpret_cast<const uint4*>(abase + (size_t)kRows * Dg * 2), (Dg / 16) * 32, coef_a);
      cbar();
      EMB_CONSUME(false, ring, p.per[4], p.ks[4], afrag4, nullptr, out, true)
      const int ncols = p.per[4] * 8, nvalid = (p.u1[4] - p.u0[4]) * 8;
      const int n0 = (p.u0[4] - g * tpg) * 8;            // first column within the group's Kh inputs
      for (int i = tid; i < kRows * ncols; i += kCThreads) {
        const int r = i / ncols, c = i - r * ncols;
        if (c >= nvalid) continue;
        const int n = n0 + c;
        const float v = out[i];
        if (n < Dg) {
          const size_t at = (size_t)r * D + (size_t)g * Dg + n;
          a.gd_carry[at] = ldcg(keep + r) * (ldcg(a.gd_tmp + at) + v);
        } else {
          // [x0 | x1] columns: this group's partial; the row CTAs sum the G partials next phase
          gxp[((size_t)g * kRows + r) * 2 * H + (n - Dg)] = v;
        }
      }
    }
    MARK(9)
    bar.sync();
    MARK(10)
  }
  if (myrow >= 0 && myrow < kRows) row_work(0, false);      // gradients wrt step 0's x0 / x1 (inputs)
#undef MARK
}

int g_sms_bwd = 0;

}  // namespace

namespace emb_tma {

int launch_bwd(const emb_rssm_bwd_args& a, void* stream, bool dry) {
  const char* who = "emb_rssm_observe_bwd";
  const int Dg = a.D / a.G, SC = a.S * a.C, Kh = Dg + 2 * a.H;
  if (!a.hoist_x2) return emb::fail(-1, "%s: the bf16 TMA engine needs hoist_x2 = 1", who);
  if (g_sms_bwd == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms_bwd, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return emb::fail_cuda(who);
  }
  if (a.ncta < rssm::kRows || a.ncta > g_sms_bwd)
    return emb::fail(-1, "%s: ncta=%d outside [16, %d SMs] (cooperative grid)", who, a.ncta, g_sms_bwd);
  auto cdiv = [](int x, int y) { return (x + y - 1) / y; };
  const int cpg = a.ncta / a.G > 1 ? a.ncta / a.G : 1;
  int unit1 = a.C;
  while (unit1 % 8) unit1 += a.C;
  unit1 /= 8;
  if ((SC / 8) % unit1) return emb::fail(-1, "%s: S*C/8=%d is not a multiple of %d tiles", who, SC / 8, unit1);
  const int per[5] = {rssm_tma::pad_tiles(cdiv(SC / 8 / unit1, a.ncta) * unit1, unit1),
                      rssm_tma::pad_tiles(cdiv(a.H / 8, a.ncta), 1),
                      rssm_tma::pad_tiles(cdiv(a.D / 8, a.ncta), 1), rssm_tma::pad_tiles(cdiv(Dg / 8, cpg), 1),
                      rssm_tma::pad_tiles(cdiv(Kh / 8, cpg), 1)};
  int maxper = 0;
  for (int i = 0; i < 5; ++i) {
    if (per[i] > rssm_tma::kMaxPer)
      return emb::fail(-1, "%s: %d tiles per CTA exceed %d (model too wide for %d CTAs)", who, per[i],
                       rssm_tma::kMaxPer, a.ncta);
    // `out` holds a layer's tile plus the slabs its k-lanes reduce through (rssm_tma.cuh out_tiles)
    if (rssm_tma::out_tiles(per[i]) > maxper) maxper = rssm_tma::out_tiles(per[i]);
  }
  int stage_bytes = rssm_tma::kStageBytesDefault, stage_cap = 0;
  if (const char* e = getenv("EMB_TMA_STAGE_KB")) stage_bytes = atoi(e) * 1024;
  if (const char* e = getenv("EMB_TMA_STAGES")) stage_cap = atoi(e);
  if (stage_bytes < 8192 || stage_bytes > 65536 || stage_bytes % 1024)
    return emb::fail(-1, "%s: EMB_TMA_STAGE_KB out of range", who);
  size_t fixed = 512 + sizeof(float) * (rssm::kRows * maxper * 8 + 64);
  fixed += bwd_a_region_bytes(a.D, a.G, a.H, SC);
  const size_t cap = 227 * 1024 - 128;
  int n = fixed + 2 * (size_t)stage_bytes <= cap ? (int)((cap - fixed) / stage_bytes) : 0;
  if (n > 12) n = 12;
  if (stage_cap > 0 && stage_cap < n) n = stage_cap;
  if (n < 2) return emb::fail(-1, "%s: no room for the weight ring next to the A operands", who);
  if (!a.frag_scratch || !a.gx_part)
    return emb::fail(-1, "%s: the bf16 TMA engine needs frag_scratch and gx_part", who);
  if (a.H > 4096 || a.H % 4) return emb::fail(-1, "%s: hidden=%d must be <= 4096 and a multiple of 4", who, a.H);
  const size_t smem = fixed + (size_t)n * stage_bytes + 128;
  emb_rssm_bwd_args copy = a;
  // L2 prefetch distance in ring chunks.  Measured (r02, size200m): 0 -> 61.5 us/step, 8 -> 68.2,
  // 16 -> 69.0, 32 -> 71.2: the stalls are latency chains, not a starved stream, and the
  // prefetched lines displace the activations the phases exchange through L2.  Off by default.
  int ahead = 0;
  if (const char* e = getenv("EMB_TMA_PREFETCH")) ahead = atoi(e);
  if (ahead < 0) ahead = 0;
  if (ahead > 127) ahead = 127;
  copy.hoist_x2 = n | (maxper << 8) | ((stage_bytes / 1024) << 16) | (ahead << 24);   // kernel-side ring configuration
  if (dry) return 0;                      // emb_rssm_tma_fits: validation and sizing only
  const void* fn = (const void*)rssm_bwd_tma_kernel;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return emb::fail_cuda(who);
  void* params[] = {&copy};
  if (cudaLaunchCooperativeKernel(fn, dim3(a.ncta), dim3(rssm_tma::kAllThreads), params, smem,
                                  (cudaStream_t)stream) != cudaSuccess)
    return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}

}  // namespace emb_tma
