// emb_rssm_observe_bwd: back-propagation through time of the fused RSSM scan
// (the reverse of rssm_fwd.cu; dreamerv3/rssm.py:61-92,135-159 differentiated).
//
// Only the sequential chain runs here: per step t = T-1 .. 0 the gradient is
// pushed back through the five in-scan layers with TRANSPOSED weights, again one
// HBM pass over every weight per step.  Parameter gradients are NOT formed in the
// loop: the kernel leaves the per-step upstream gradients of every layer
// (g_xo, g_logit, g_gates, g_h, g_x0, g_x1, g_x2) in [T][16][..] buffers and the
// host turns them into dW = A^T G with (B*T)-row tensor-core GEMMs afterwards.
//
//   B1  g_stoch = G_stoch[t] + (keep' * g_y1') @ dynin1^T                        (2.1 M)
//   B2  g_logit = G_logit[t] + unimix-softmax-jacobian(g_stoch) ; g_xo = g_logit @ obslogit^T (2.1 M)
//   B3  g_deter = G_deter[t] + carry + [g_yobs | keep' * g_y0'] @ [obs0[:D] | dynin0]^T    (16.8 M)
//       + GRU gate backward in the epilogue -> g_gates
//   B4  g_h     = g_gates_g @ dyngru[g]^T ; row dots for the rms-norm backward   (25.2 M)
//   B5  g_in    = g_yhid_g @ dynhid0[g]^T -> carry (deter part), g_x0/g_x1/g_x2 (summed over groups) (33.5 M)
// (primes = step t+1;  g_y* = rms-norm + silu backward of g_x*, recomputed where
// it is consumed.)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"
#include "rssm_common.cuh"

namespace {

using namespace rssm;

template <bool FAST>
__device__ __forceinline__ float dsilu(float n) {          // d silu(n) / dn
  const float sg = FAST ? __fdividef(1.0f, 1.0f + __expf(-n)) : 1.0f / (1.0f + expf(-n));
  return sg * (1.0f + n * (1.0f - sg));
}

// rms-norm + silu backward (nets.py:374-383):  with n = y * rstd * s,
//   g_n = g_x * silu'(n),   g_y = rstd * s * g_n - y * coef,   coef = rstd^3 * mean_k(g_n s y)
// The row statistic `dot = sum_k g_n s y` is accumulated by the PRODUCER of g_x
// (atomics into a.dots), so consumers apply this element-wise.
template <bool FAST>
__device__ __forceinline__ float norm_bwd_elem(float gx, float y, float s, float rstd, float coef) {
  const float gn = gx * dsilu<FAST>(y * rstd * s);
  return rstd * s * gn - y * coef;
}

// A fragments from two fp32 [16][n] sources: f(r, k, v1, v2), four k per thread.
template <typename F>
__device__ __forceinline__ void build_part2(__nv_bfloat16* afrag, int koff, const float* s1, int ld1,
                                            const float* s2, int ld2, int n, F f) {
  const int n4 = n >> 2;
#pragma unroll 4
  for (int i = threadIdx.x; i < kRows * n4; i += kThreads) {
    const int r = i / n4, k = (i - r * n4) << 2;
    const float4 v = __ldcg(reinterpret_cast<const float4*>(s1 + (size_t)r * ld1 + k));
    const float4 w = __ldcg(reinterpret_cast<const float4*>(s2 + (size_t)r * ld2 + k));
    *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, koff + k)) =
        __floats2bfloat162_rn(f(r, k, v.x, w.x), f(r, k + 1, v.y, w.y));
    *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, koff + k + 2)) =
        __floats2bfloat162_rn(f(r, k + 2, v.z, w.z), f(r, k + 3, v.w, w.w));
  }
}

template <int ENG>
__global__ void __launch_bounds__(kThreads, 1)
rssm_bwd_kernel(const __grid_constant__ emb_rssm_bwd_args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool BF = ENG == ENG_BF16;
  const int T = a.T, D = a.D, H = a.H, S = a.S, C = a.C, G = a.G;
  const int Dg = D / G, SC = S * C, Kh = Dg + (a.hoist_x2 ? 2 : 3) * H;
  float* out = reinterpret_cast<float*>(smem_raw);                   // [16][kMaxTiles*8]
  float* st = out + kRows * kMaxTiles * 8;                           // 4 x [16] row statistics
  float *rstd_a = st, *coef_a = st + 16, *rstd_b = st + 32, *coef_b = st + 48;
  __nv_bfloat16* afrag = reinterpret_cast<__nv_bfloat16*>(st + 64);
  uint4* afrag4 = reinterpret_cast<uint4*>(afrag);
  auto red_after = [&](int K) -> float* {
    return reinterpret_cast<float*>(afrag + (BF ? (size_t)kRows * K : 0));
  };

  GridBarrier bar{a.barrier, 0};
  const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x;
  const size_t RH = (size_t)kRows * H, RD = (size_t)kRows * D, RSC = (size_t)kRows * SC;

  // work split; must match scan.py pack_bwd()
  const int tiles_b1 = SC / 8, per_b1 = (tiles_b1 + ncta - 1) / ncta;
  const int tiles_b2 = H / 8, per_b2 = (tiles_b2 + ncta - 1) / ncta;
  const int tiles_b3 = D / 8, per_b3 = (tiles_b3 + ncta - 1) / ncta;
  const GroupSplit sp_b4 = group_split(Dg / 8, G), sp_b5 = group_split(Kh / 8, G);
  const int per_b4 = sp_b4.per, per_b5 = sp_b5.per;
  const uint2* blk_b1 = reinterpret_cast<const uint2*>(a.wt_in1) + (size_t)cta * (H / 16) * per_b1 * 32;
  const uint2* blk_b2 = reinterpret_cast<const uint2*>(a.wt_logit) + (size_t)cta * (SC / 16) * per_b2 * 32;
  const uint2* blk_b3 = reinterpret_cast<const uint2*>(a.wt_ph1) + (size_t)cta * (2 * H / 16) * per_b3 * 32;
  const uint2* blk_b4 = reinterpret_cast<const uint2*>(a.wt_gru) + (size_t)cta * (3 * Dg / 16) * per_b4 * 32;
  const uint2* blk_b5 = reinterpret_cast<const uint2*>(a.wt_hid) + (size_t)cta * (Dg / 16) * per_b5 * 32;
  const size_t bytes_b1 = (size_t)(H / 16) * per_b1 * 256, bytes_b2 = (size_t)(SC / 16) * per_b2 * 256;
  const size_t bytes_b3 = (size_t)(2 * H / 16) * per_b3 * 256, bytes_b4 = (size_t)(3 * Dg / 16) * per_b4 * 256;
  const size_t bytes_b5 = (size_t)(Dg / 16) * per_b5 * 256;
  const float* wf_b1 = reinterpret_cast<const float*>(a.wt_in1);
  const float* wf_b2 = reinterpret_cast<const float*>(a.wt_logit);
  const float* wf_b3 = reinterpret_cast<const float*>(a.wt_ph1);
  const float* wf_b4 = reinterpret_cast<const float*>(a.wt_gru);
  const float* wf_b5 = reinterpret_cast<const float*>(a.wt_hid);

  // row statistics of a normalised layer at step `ts`: slot 0 x0, 1 x1, 2 xo
  auto load_stats = [&](int ts, int slot, int n, float* rstd, float* coef) {
    if (tid < kRows) {
      const float rs = a.rstd[(size_t)ts * 3 * kRows + slot * kRows + tid];
      rstd[tid] = rs;
      coef[tid] = rs * rs * rs * ldcg(a.dots + ((size_t)ts * 4 + slot) * kRows + tid) / (float)n;
    }
  };
  // sum out[r][c0..c1) per row and add it to a.dots[ts][slot][r]
  auto add_row_dots = [&](int ts, int slot, int ncols, int c0, int c1) {
    if (tid < kRows && c0 < c1) {
      float sum = 0.f;
      for (int c = c0; c < c1; ++c) sum += out[tid * ncols + c];
      atomicAdd(a.dots + ((size_t)ts * 4 + slot) * kRows + tid, sum);
    }
  };

  for (int t = T - 1; t >= 0; --t) {
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
    // g_stoch[t] = G_stoch[t] + (keep' * g_y1') @ dynin1^T      (scratch: a.g_stoch)
    {
      if (BF) { prefetch_l2(blk_b2, bytes_b2); prefetch_l2(blk_b3, bytes_b3); }   // one phase ahead
      const int u0 = min(tiles_b1, cta * per_b1), u1 = min(tiles_b1, u0 + per_b1);
      if (u0 < u1) {
        load_stats(t + 1, 1, H, rstd_a, coef_a);
        __syncthreads();
        auto aval = [&](int r, int k) -> float {
          return ldcg(keep_next + r) * norm_bwd_elem<false>(
              ldcg(gx1n + (size_t)r * H + k), y1n[(size_t)r * H + k], a.s1[k], rstd_a[r], coef_a[r]);
        };
        if (BF) {
          build_part2(afrag, 0, gx1n, H, y1n, H, H, [&](int r, int k, float gx, float y) {
            return ldcg(keep_next + r) * norm_bwd_elem<true>(gx, y, a.s1[k], rstd_a[r], coef_a[r]); });
          __syncthreads();
        }
        for (int base = u0; base < u1; base += kMaxTiles) {
          const int nt = min(kMaxTiles, u1 - base);
          tile_gemm<ENG, false>(blk_b1, per_b1, base - u0, wf_b1, base, nt, H, afrag4, aval, out, red_after(H));
          const int ncols = nt * 8;
          for (int i = tid; i < kRows * ncols; i += kThreads) {
            const int r = i / ncols, c = i - r * ncols;
            const int col = base * 8 + c;
            a.g_stoch[(size_t)r * SC + col] = out[i] + a.G_stoch[(size_t)t * RSC + (size_t)r * SC + col];
          }
          __syncthreads();
        }
      }
    }
    bar.sync();

    // ------------------------------------------------------------------ B2
    // g_logit = G_logit + (1-eps) p (g_stoch - sum_c p g_stoch) ;  g_xo = g_logit @ obslogit^T
    {
      const int u0 = min(tiles_b2, cta * per_b2), u1 = min(tiles_b2, u0 + per_b2);
      // every CTA needs the full g_logit rows as its A operand; row r is also
      // written out (for dW) by CTA ncta-1-r.
      const int wrow = ncta - 1 - cta;
      if (u0 < u1 || (wrow >= 0 && wrow < kRows)) {
        // bf16 engine: the rows go straight into A fragments; fp32 engine: fp32 rows in shared memory
        float* gl = reinterpret_cast<float*>(afrag);            // fp32 engine only: [16][SC]
        const float* pr = a.probs + (size_t)t * RSC;
        const float* Gl = a.G_logit + (size_t)t * RSC;
        // one thread per (row, latent): all loads are independent 16-byte loads
        for (int grp = tid; grp < kRows * S; grp += kThreads) {
          const int r = grp / S, sv = grp - r * S;
          const size_t o = (size_t)r * SC + (size_t)sv * C;
          float dot = 0.f;
          for (int c = 0; c < C; c += 4) {
            const float4 p = *reinterpret_cast<const float4*>(pr + o + c);
            const float4 g = __ldcg(reinterpret_cast<const float4*>(a.g_stoch + o + c));
            dot += p.x * g.x + p.y * g.y + p.z * g.z + p.w * g.w;
          }
          for (int c = 0; c < C; c += 4) {
            const float4 p = *reinterpret_cast<const float4*>(pr + o + c);
            const float4 g = __ldcg(reinterpret_cast<const float4*>(a.g_stoch + o + c));
            const float4 e = *reinterpret_cast<const float4*>(Gl + o + c);
            const float um = 1.0f - a.unimix;
            float4 v;
            v.x = e.x + um * p.x * (g.x - dot); v.y = e.y + um * p.y * (g.y - dot);
            v.z = e.z + um * p.z * (g.z - dot); v.w = e.w + um * p.w * (g.w - dot);
            if (BF) {
              const int k = sv * C + c;
              *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, k)) = __floats2bfloat162_rn(v.x, v.y);
              *reinterpret_cast<__nv_bfloat162*>(afrag + afrag_index(r, k + 2)) = __floats2bfloat162_rn(v.z, v.w);
            } else {
              *reinterpret_cast<float4*>(gl + o + c) = v;
            }
            if (wrow == r) *reinterpret_cast<float4*>(g_logit + o + c) = v;
          }
        }
        if (u0 < u1) load_stats(t, 2, H, rstd_a, coef_a);       // only rstd is used here
        __syncthreads();
        auto aval = [&](int r, int k) -> float { return gl[(size_t)r * SC + k]; };
        for (int base = u0; base < u1; base += kMaxTiles) {
          const int nt = min(kMaxTiles, u1 - base);
          tile_gemm<ENG, false>(blk_b2, per_b2, base - u0, wf_b2, base, nt, SC, afrag4, aval, out, red_after(SC));
          const int ncols = nt * 8;
          for (int i = tid; i < kRows * ncols; i += kThreads) {
            const int r = i / ncols, c = i - r * ncols;
            const int col = base * 8 + c;
            const float gx = out[i];
            g_xo[(size_t)r * H + col] = gx;
            const float y = yobs[(size_t)r * H + col], sc = a.s_obs[col];
            out[i] = gx * dsilu<BF>(y * rstd_a[r] * sc) * sc * y;      // g_n * s * y
          }
          __syncthreads();
          add_row_dots(t, 2, ncols, 0, ncols);
          __syncthreads();
        }
      }
    }
    bar.sync();

    // ------------------------------------------------------------------ B3
    // g_deter = G_deter[t] + carry + [g_yobs | keep' g_y0'] @ [obs0[:D] | dynin0]^T, then the
    // GRU backward (rssm.py:152-158) for the same columns.
    {
      if (BF) prefetch_l2(blk_b4, bytes_b4);
      const int u0 = min(tiles_b3, cta * per_b3), u1 = min(tiles_b3, u0 + per_b3);
      if (u0 < u1) {
        load_stats(t, 2, H, rstd_a, coef_a);
        load_stats(t + 1, 0, H, rstd_b, coef_b);
        __syncthreads();
        auto aval = [&](int r, int k) -> float {
          if (k < H)
            return norm_bwd_elem<false>(ldcg(g_xo + (size_t)r * H + k), yobs[(size_t)r * H + k],
                                        a.s_obs[k], rstd_a[r], coef_a[r]);
          k -= H;
          return ldcg(keep_next + r) * norm_bwd_elem<false>(
              ldcg(gx0n + (size_t)r * H + k), y0n[(size_t)r * H + k], a.s0[k], rstd_b[r], coef_b[r]);
        };
        if (BF) {
          build_part2(afrag, 0, g_xo, H, yobs, H, H, [&](int r, int k, float gx, float y) {
            return norm_bwd_elem<true>(gx, y, a.s_obs[k], rstd_a[r], coef_a[r]); });
          build_part2(afrag, H, gx0n, H, y0n, H, H, [&](int r, int k, float gx, float y) {
            return ldcg(keep_next + r) * norm_bwd_elem<true>(gx, y, a.s0[k], rstd_b[r], coef_b[r]); });
          __syncthreads();
        }
        for (int base = u0; base < u1; base += kMaxTiles) {
          const int nt = min(kMaxTiles, u1 - base);
          tile_gemm<ENG, false>(blk_b3, per_b3, base - u0, wf_b3, base, nt, 2 * H, afrag4, aval, out, red_after(2 * H));
          const int ncols = nt * 8;
          for (int i = tid; i < kRows * ncols; i += kThreads) {
            const int r = i / ncols, c = i - r * ncols;
            const int col = base * 8 + c;
            const size_t at = (size_t)r * D + col;
            const float gd = out[i] + a.G_deter[(size_t)t * RD + at] + ldcg(a.gd_carry + at);
            const float rs = gates[at], cand = gates[RD + at], up = gates[2 * RD + at], cpre = gates[3 * RD + at];
            const float old = ldcg(keep + r) * deter_prev[at];
            const float g_u = gd * (cand - old), g_c = gd * up;
            a.gd_tmp[at] = gd * (1.0f - up);                 // direct path into keep*deter_{t-1}
            const float g_rc = g_c * (1.0f - cand * cand);
            const int g = col / Dg, jj = col - g * Dg;
            float* gg = g_gates + (size_t)r * 3 * D + (size_t)g * 3 * Dg + jj;
            gg[0] = g_rc * cpre * rs * (1.0f - rs);          // reset gate, pre-sigmoid
            gg[Dg] = g_rc * rs;                              // candidate, pre-tanh
            gg[2 * Dg] = g_u * up * (1.0f - up);             // update gate, pre-sigmoid
          }
          __syncthreads();
        }
      }
    }
    bar.sync();

    // ------------------------------------------------------------------ B4
    // g_h_g = g_gates_g @ dyngru[g]^T ; row dots of the dynhid0 norm backward
    {
      if (BF) prefetch_l2(blk_b5, bytes_b5);
      const int u0 = sp_b4.u0, u1 = sp_b4.u1;
      const int tpg = Dg / 8;
      if (u0 < u1 && tid < kRows)
        rstd_a[tid] = rsqrtf(a.sumsq[(size_t)t * kRows + tid] / (float)D + a.eps);
      __syncthreads();
      for (int tile = u0; tile < u1;) {
        const int g = tile / tpg;
        const int seg_end = min(u1, (g + 1) * tpg);
        const float* src = g_gates + (size_t)g * 3 * Dg;
        auto aval = [&](int r, int k) -> float { return ldcg(src + (size_t)r * 3 * D + k); };
        if (BF) {
          build_part(afrag, 0, src, 3 * Dg, 3 * D, [&](int, int, float v) { return v; });
          __syncthreads();
        }
        for (int base = tile; base < seg_end; base += kMaxTiles) {
          const int nt = min(kMaxTiles, seg_end - base);
          tile_gemm<ENG, false>(blk_b4, per_b4, base - u0, wf_b4, base, nt, 3 * Dg, afrag4, aval, out, red_after(3 * Dg));
          const int ncols = nt * 8;
          for (int i = tid; i < kRows * ncols; i += kThreads) {
            const int r = i / ncols, c = i - r * ncols;
            const int col = base * 8 + c;
            const size_t at = (size_t)r * D + col;
            const float gh = out[i];
            g_h[at] = gh;
            const float y = yhid[at], sc = a.s_hid[col];
            out[i] = gh * dsilu<BF>(y * rstd_a[r] * sc) * sc * y;    // g_n * s * y
          }
          __syncthreads();
          add_row_dots(t, 3, ncols, 0, ncols);
          __syncthreads();
        }
        tile = seg_end;
      }
    }
    bar.sync();

    // ------------------------------------------------------------------ B5
    // g_in_g = g_yhid_g @ dynhid0[g]^T : deter part -> carry, x0/x1/x2 parts summed over groups
    {
      if (BF && t > 0) prefetch_l2(blk_b1, bytes_b1);
      const int u0 = sp_b5.u0, u1 = sp_b5.u1;
      const int tpg = Kh / 8;
      if (u0 < u1 && tid < kRows) {
        const float rs = rsqrtf(a.sumsq[(size_t)t * kRows + tid] / (float)D + a.eps);
        rstd_a[tid] = rs;
        coef_a[tid] = rs * rs * rs * ldcg(a.dots + ((size_t)t * 4 + 3) * kRows + tid) / (float)D;
        rstd_b[tid] = a.rstd[(size_t)t * 3 * kRows + tid];             // y0[t]
        coef_b[tid] = a.rstd[(size_t)t * 3 * kRows + kRows + tid];     // y1[t] (rstd, not a coef)
      }
      __syncthreads();
      for (int tile = u0; tile < u1;) {
        const int g = tile / tpg;
        const int seg_end = min(u1, (g + 1) * tpg);
        auto aval = [&](int r, int k) -> float {
          const size_t at = (size_t)r * D + (size_t)g * Dg + k;
          return norm_bwd_elem<false>(ldcg(g_h + at), yhid[at], a.s_hid[g * Dg + k], rstd_a[r], coef_a[r]);
        };
        if (BF) {
          build_part2(afrag, 0, g_h + (size_t)g * Dg, D, yhid + (size_t)g * Dg, D, Dg,
                      [&](int r, int k, float gx, float y) {
            return norm_bwd_elem<true>(gx, y, a.s_hid[g * Dg + k], rstd_a[r], coef_a[r]); });
          __syncthreads();
        }
        for (int base = tile; base < seg_end; base += kMaxTiles) {
          const int nt = min(kMaxTiles, seg_end - base);
          tile_gemm<ENG, false>(blk_b5, per_b5, base - u0, wf_b5, base, nt, Dg, afrag4, aval, out, red_after(Dg));
          const int ncols = nt * 8;
          const int n0 = (base - g * tpg) * 8;               // first column within the group's Kh inputs
          for (int i = tid; i < kRows * ncols; i += kThreads) {
            const int r = i / ncols, c = i - r * ncols;
            const int n = n0 + c;
            const float v = out[i];
            float prod = 0.f;
            if (n < Dg) {
              const size_t at = (size_t)r * D + (size_t)g * Dg + n;
              a.gd_carry[at] = ldcg(keep + r) * (ldcg(a.gd_tmp + at) + v);
            } else {
              const int m = n - Dg;                          // [x0 | x1 | x2]
              const int which = m / H, k = m - which * H;
              float* dst = which == 0 ? a.g_x0 : (which == 1 ? a.g_x1 : a.g_x2);
              atomicAdd(dst + (size_t)t * RH + (size_t)r * H + k, v);
              if (which == 0) {
                const float y = y0[(size_t)r * H + k], sc = a.s0[k];
                prod = v * dsilu<BF>(y * rstd_b[r] * sc) * sc * y;
              } else if (which == 1) {
                const float y = y1[(size_t)r * H + k], sc = a.s1[k];
                prod = v * dsilu<BF>(y * coef_b[r] * sc) * sc * y;
              }
            }
            out[i] = prod;
          }
          __syncthreads();
          // columns [Dg, Dg+H) feed dots slot 0 (x0), [Dg+H, Dg+2H) slot 1 (x1)
          add_row_dots(t, 0, ncols, max(0, Dg - n0), min(ncols, Dg + H - n0));
          add_row_dots(t, 1, ncols, max(0, Dg + H - n0), min(ncols, Dg + 2 * H - n0));
          __syncthreads();
        }
        tile = seg_end;
      }
    }
    bar.sync();
  }
}

size_t bwd_smem_bytes(const emb_rssm_bwd_args& a) {
  const size_t n = sizeof(float) * (kRows * kMaxTiles * 8 + 64);
  const int Dg = a.D / a.G, SC = a.S * a.C;
  if (a.engine == rssm::ENG_F32) return n + (size_t)kRows * SC * sizeof(float);
  int kmax = 3 * Dg;
  if (2 * a.H > kmax) kmax = 2 * a.H;
  if (SC > kmax) kmax = SC;
  return n + (size_t)kRows * kmax * 2 + (size_t)kWarps * kRows * kMaxTiles * 8 * sizeof(float);
}

int g_sms = 0;

}  // namespace

namespace emb_legacy {
size_t bwd_smem(const emb_rssm_bwd_args& a) { return bwd_smem_bytes(a); }
}
namespace emb_tma {
int launch_bwd(const emb_rssm_bwd_args& a, void* stream, bool dry);   // rssm_bwd_tma.cu
}

extern "C" int emb_rssm_observe_bwd(const emb_rssm_bwd_args* args, void* stream) {
  const char* who = "emb_rssm_observe_bwd";
  if (!args) return emb::fail(-1, "%s: args is NULL", who);
  const emb_rssm_bwd_args& a = *args;
  if (a.B < 1 || a.B > kRows) return emb::fail(-1, "%s: B=%d outside [1,16]", who, a.B);
  if (a.T < 1) return emb::fail(-1, "%s: T=%d < 1", who, a.T);
  if (a.G < 1 || a.D % a.G || (a.D / a.G) % 16 || a.H % 16 || (a.S * a.C) % 16 || a.D % 16)
    return emb::fail(-1, "%s: D/G, H and S*C must be multiples of 16", who);
  if (a.engine != rssm::ENG_F32 && a.engine != rssm::ENG_TMA && a.engine != rssm::ENG_LEGACY)
    return emb::fail(-1, "%s: engine %d", who, a.engine);
  if (a.engine == rssm::ENG_TMA) return emb_tma::launch_bwd(a, stream, false);
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return emb::fail_cuda(who);
  }
  if (a.ncta < 1 || a.ncta > g_sms)
    return emb::fail(-1, "%s: ncta=%d outside [1, %d SMs] (cooperative grid)", who, a.ncta, g_sms);
  const size_t smem = bwd_smem_bytes(a);
  if (smem > 227 * 1024)
    return emb::fail(-1, "%s: needs %zu bytes of shared memory (> 227 KiB)", who, smem);
  const void* fn = a.engine != rssm::ENG_F32 ? (const void*)rssm_bwd_kernel<rssm::ENG_BF16>
                                             : (const void*)rssm_bwd_kernel<rssm::ENG_F32>;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return emb::fail_cuda(who);
  emb_rssm_bwd_args copy = a;
  void* params[] = {&copy};
  if (cudaLaunchCooperativeKernel(fn, dim3(a.ncta), dim3(kThreads), params, smem,
                                  (cudaStream_t)stream) != cudaSuccess)
    return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}
