// emb_rssm_observe_fwd: the T-step recurrent scan of RSSM.observe as ONE
// persistent cooperative kernel (dreamerv3/rssm.py:61-92 `_observe`, :135-159
// `_core`, embodied/jax/outs.py:208-270 unimix + straight-through sample).
//
// Per step t (phases separated by grid barriers; every weight streamed once):
//   P4  yhid  = [keep*deter_g, x0, x1, x2] @ dynhid0[g] + b         (33.5 M weights at size200m)
//   P5  gates = silu(rms(yhid))_g @ dyngru[g] + b ; GRU -> deter_t   (25.2 M)
//   P1  yobs  = deter_t @ obs0[:D] + pre_tok_t ;  y0' = keep'*(deter_t @ dynin0) + b   (16.8 M)
//   P2  logit = silu(rms(yobs)) @ obslogit + b                       (2.1 M)
//   P3  idx   = argmax(log unimix(softmax(logit)) + gumbel) ;  y1' = keep' * sum_s dynin1[s*C+idx_s] + b
// Hoisted by the caller (they do not depend on the recurrent state): x2 =
// silu(rms(dynin2(action))), pre_tok = tokens @ obs0[D:] + b, and step 0's y0/y1.
// The one-hot stoch never materialises: dynin1 is a row gather.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"
#include "rssm_common.cuh"

namespace {

using namespace rssm;

struct Dims {
  int B, T, D, H, S, C, G, Dg, SC;
  float unimix, eps;
};

template <int ENG>
__global__ void __launch_bounds__(kThreads, 1)
rssm_fwd_kernel(const __grid_constant__ emb_rssm_fwd_args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Dims d = {a.B, a.T, a.D, a.H, a.S, a.C, a.G, a.D / a.G, a.S * a.C, a.unimix, a.eps};
  const int Kh = d.Dg + 3 * d.H;                 // dynhid0 input width per group
  constexpr bool BF = ENG == ENG_BF16;
  // shared memory carve-up
  float* out = reinterpret_cast<float*>(smem_raw);                   // [16][kMaxTiles*8]
  float* rstd_a = out + kRows * kMaxTiles * 8;                       // [16]
  float* rstd_b = rstd_a + kRows;                                    // [16]
  __nv_bfloat16* afrag = reinterpret_cast<__nv_bfloat16*>(rstd_b + kRows);
  uint4* afrag4 = reinterpret_cast<uint4*>(afrag);
  // per-warp partial sums live right behind the phase's A fragments (K * 32 B)
  auto red_after = [&](int K) -> float* {
    return reinterpret_cast<float*>(afrag + (BF ? (size_t)kRows * K : 0));
  };
  auto act = [](float x) -> float { return BF ? silu_fast(x) : silu_f(x); };

  GridBarrier bar{a.barrier, 0};
  const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x;
  const size_t RH = (size_t)kRows * d.H, RD = (size_t)kRows * d.D, RSC = (size_t)kRows * d.SC;

  // static work split (must match scan.py pack(): per = ceil(tiles / ncta))
  const GroupSplit sp_hid = group_split(d.Dg / 8, d.G), sp_gru = group_split(d.Dg / 8, d.G);
  const int per_hid = sp_hid.per, per_gru = sp_gru.per;
  const int tiles_ph1 = 2 * d.H / 8, per_ph1 = (tiles_ph1 + ncta - 1) / ncta;
  const int tiles_log = d.SC / 8, per_log = (tiles_log + ncta - 1) / ncta;
  const uint2* blk_hid = reinterpret_cast<const uint2*>(a.w_hid) + (size_t)cta * (Kh / 16) * per_hid * 32;
  const uint2* blk_gru = reinterpret_cast<const uint2*>(a.w_gru) + (size_t)cta * (d.Dg / 16) * per_gru * 3 * 32;
  const uint2* blk_ph1 = reinterpret_cast<const uint2*>(a.w_ph1) + (size_t)cta * (d.D / 16) * per_ph1 * 32;
  const uint2* blk_log = reinterpret_cast<const uint2*>(a.w_logit) + (size_t)cta * (d.H / 16) * per_log * 32;
  const float* wf_hid = reinterpret_cast<const float*>(a.w_hid);
  const float* wf_gru = reinterpret_cast<const float*>(a.w_gru);
  const float* wf_ph1 = reinterpret_cast<const float*>(a.w_ph1);
  const float* wf_log = reinterpret_cast<const float*>(a.w_logit);
  __nv_bfloat16* deterA = reinterpret_cast<__nv_bfloat16*>(a.deterA);
  const size_t bytes_hid = (size_t)(Kh / 16) * per_hid * 256, bytes_gru = (size_t)(d.Dg / 16) * per_gru * 3 * 256;
  const size_t bytes_ph1 = (size_t)(d.D / 16) * per_ph1 * 256, bytes_log = (size_t)(d.H / 16) * per_log * 256;
  __nv_bfloat16* x0A = deterA + 2 * RD;                       // [16*H] x0 fragments
  // x0 = silu(rms(y0)) is built ONCE per step, row r by CTA ncta-1-r, instead
  // of by every CTA in its P4 prologue.
  auto build_x0 = [&](const float* y0v, int tsave) {
    const int r = ncta - 1 - cta;
    if (!BF || r < 0 || r >= kRows) return;
    float s = 0.f;
    for (int i = tid * 4; i < d.H; i += kThreads * 4) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(y0v + (size_t)r * d.H + i));
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) out[tid >> 5] = s;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < kWarps; ++w) tot += out[w];
    const float rstd = rsqrtf(tot / (float)d.H + d.eps);
    if (tid == 0) a.rstd[(size_t)tsave * 3 * kRows + r] = rstd;
    for (int k = tid * 2; k < d.H; k += kThreads * 2) {
      const float2 v = __ldcg(reinterpret_cast<const float2*>(y0v + (size_t)r * d.H + k));
      *reinterpret_cast<__nv_bfloat162*>(x0A + afrag_index(r, k)) = __floats2bfloat162_rn(
          silu_fast(v.x * (rstd * a.s0[k])), silu_fast(v.y * (rstd * a.s0[k + 1])));
    }
    __syncthreads();
  };

  if (BF) {
    // deter0 -> A fragments (slot 1 = "(t-1) & 1" of step 0)
    for (size_t i = (size_t)cta * kThreads + tid; i < RD; i += (size_t)ncta * kThreads) {
      const int r = (int)(i / d.D), k = (int)(i - (size_t)r * d.D);
      deterA[RD + afrag_index(r, k)] = __float2bfloat16_rn(a.deter0[i]);
    }
    build_x0(a.y0, 0);
    prefetch_l2(blk_hid, bytes_hid);
    bar.sync();
  }

#define MARK(i)                                                              \
  if (a.timing && cta == 0 && tid == 0) {                                    \
    unsigned long long now_;                                                 \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_));                  \
    a.timing[(size_t)t * 16 + (i)] = now_;                                   \
  }
  for (int t = 0; t < d.T; ++t) {
    MARK(0)
    const float* keep = a.keep + (size_t)t * kRows;
    const float* keep_next = a.keep + (size_t)(t + 1) * kRows;
    const float* deter_prev = t == 0 ? a.deter0 : a.deter + (size_t)(t - 1) * RD;
    const float* y0 = a.y0 + (size_t)t * RH;
    const float* y1 = a.y1 + (size_t)t * RH;
    const float* x2 = reinterpret_cast<const float*>(a.x2) + (size_t)t * RH;   // fp32 engine view
    float* yhid = a.yhid + (size_t)t * RD;
    const bool last = t + 1 == d.T;

    // ------------------------------------------------------------------ P4
    {
      if (BF) prefetch_l2(blk_gru, bytes_gru);                // P5's weights, one phase ahead
      const int u0 = sp_hid.u0, u1 = sp_hid.u1;
      const int tpg = d.Dg / 8;                   // tiles per group
      if (u0 < u1) {
        if (!BF) row_rstd(y0, d.H, d.eps, rstd_a);
        row_rstd(y1, d.H, d.eps, rstd_b);
        __syncthreads();
        if (cta == 0 && tid < kRows) {
          if (!BF) a.rstd[(size_t)t * 3 * kRows + tid] = rstd_a[tid];
          a.rstd[(size_t)t * 3 * kRows + kRows + tid] = rstd_b[tid];
        }
        if (BF) {   // the group-independent part of A: x0 (prebuilt), x1, x2 (prebuilt by the host)
          copy_frags(afrag4 + (d.Dg / 16) * 32, reinterpret_cast<const uint4*>(x0A), (d.H / 16) * 32);
          build_part(afrag, d.Dg + d.H, y1, d.H, d.H, [&](int r, int k, float v) {
            return silu_fast(v * (rstd_b[r] * a.s1[k])); });
          copy_frags(afrag4 + ((d.Dg + 2 * d.H) / 16) * 32,
                     reinterpret_cast<const uint4*>(a.x2) + (size_t)t * (d.H / 16) * 32, (d.H / 16) * 32);
        }
      }
      MARK(1)
      for (int tile = u0; tile < u1;) {
        const int g = tile / tpg;
        const int seg_end = min(u1, (g + 1) * tpg);
        auto aval = [&](int r, int k) -> float {
          if (k < d.Dg) return ldcg(keep + r) * ldcg(deter_prev + (size_t)r * d.D + g * d.Dg + k);
          k -= d.Dg;
          if (k < d.H) return silu_f(ldcg(y0 + (size_t)r * d.H + k) * (rstd_a[r] * a.s0[k]));
          k -= d.H;
          if (k < d.H) return silu_f(ldcg(y1 + (size_t)r * d.H + k) * (rstd_b[r] * a.s1[k]));
          k -= d.H;
          return ldcg(x2 + (size_t)r * d.H + k);
        };
        if (BF) {   // deter slice of group g: copy fragments, zero the reset rows
          const uint4* src = reinterpret_cast<const uint4*>(deterA + (size_t)((t + 1) & 1) * RD) +
              (size_t)(g * d.Dg / 16) * 32;
          for (int i = tid; i < (d.Dg / 16) * 32; i += kThreads) {
            uint4 v = ldcg_u4(src + i);
            const int r = (i & 31) >> 2;
            if (ldcg(keep + r) == 0.f) { v.x = 0; v.z = 0; }
            if (ldcg(keep + r + 8) == 0.f) { v.y = 0; v.w = 0; }
            afrag4[i] = v;
          }
          __syncthreads();
        }
        for (int base = tile; base < seg_end; base += kMaxTiles) {
          const int nt = min(kMaxTiles, seg_end - base);
          tile_gemm<ENG, false>(blk_hid, per_hid, base - u0, wf_hid, base, nt, Kh, afrag4, aval, out, red_after(Kh));
          const int ncols = nt * 8;
          // epilogue: + bias -> yhid ; row sums of squares -> sumsq[t]
          for (int i = tid; i < kRows * ncols; i += kThreads) {
            const int r = i / ncols, c = i - r * ncols;
            const int col = base * 8 + c;
            const float v = out[i] + a.b_hid[col];
            yhid[(size_t)r * d.D + col] = v;
            out[i] = v * v;
          }
          __syncthreads();
          if (tid < kRows) {
            float s = 0.f;
            for (int c = 0; c < ncols; ++c) s += out[tid * ncols + c];
            atomicAdd(a.sumsq + (size_t)t * kRows + tid, s);
          }
          __syncthreads();
        }
        tile = seg_end;
      }
    }
    MARK(2)
    bar.sync();
    MARK(3)

    // ------------------------------------------------------------------ P5
    float* deter = a.deter + (size_t)t * RD;
    {
      if (BF) prefetch_l2(blk_ph1, bytes_ph1);
      const int u0 = sp_gru.u0, u1 = sp_gru.u1;
      const int upg = d.Dg / 8;
      if (u0 < u1) {
        if (tid < kRows)
          rstd_a[tid] = rsqrtf(ldcg(a.sumsq + (size_t)t * kRows + tid) / (float)d.D + d.eps);
        __syncthreads();
      }
      for (int unit = u0; unit < u1;) {
        const int g = unit / upg;
        const int seg_end = min(u1, (g + 1) * upg);
        auto aval = [&](int r, int k) -> float {
          const int col = g * d.Dg + k;
          return silu_f(ldcg(yhid + (size_t)r * d.D + col) * (rstd_a[r] * a.s_hid[col]));
        };
        if (BF) {
          build_part(afrag, 0, yhid + g * d.Dg, d.Dg, d.D, [&](int r, int k, float v) {
            return silu_fast(v * (rstd_a[r] * a.s_hid[g * d.Dg + k])); });
          __syncthreads();
        }
        constexpr int kMaxUnits = kMaxTiles / 3;
        for (int base = unit; base < seg_end; base += kMaxUnits) {
          const int nu = min(kMaxUnits, seg_end - base);
          tile_gemm<ENG, false>(blk_gru, per_gru * 3, (base - u0) * 3, wf_gru, base * 3, nu * 3,
                                d.Dg, afrag4, aval, out, red_after(d.Dg));
          const int ncols = nu * 24;
          // epilogue: GRU gates (rssm.py:152-158)
          for (int i = tid; i < kRows * nu * 8; i += kThreads) {
            const int r = i / (nu * 8), c = i - r * (nu * 8);
            const int u = c >> 3, nn = c & 7;
            const int jj = (base - g * upg + u) * 8 + nn;           // column within the group
            const int col = g * d.Dg + jj;
            const float* o = out + r * ncols + u * 24 + nn;
            const float* bg = a.b_gru + (size_t)g * 3 * d.Dg + jj;
            const float rs = sigmoid_f(o[0] + bg[0]);
            const float cpre = o[8] + bg[d.Dg];
            const float cand = tanhf(rs * cpre);
            const float up = sigmoid_f(o[16] + bg[2 * d.Dg] - 1.0f);
            const float old = ldcg(keep + r) * ldcg(deter_prev + (size_t)r * d.D + col);
            const float nw = up * cand + (1.0f - up) * old;
            deter[(size_t)r * d.D + col] = nw;
            float* gs = a.gates + (size_t)t * 4 * RD + (size_t)r * d.D + col;
            gs[0] = rs; gs[RD] = cand; gs[2 * RD] = up; gs[3 * RD] = cpre;
            if (BF) deterA[(size_t)(t & 1) * RD + afrag_index(r, col)] = __float2bfloat16_rn(nw);
          }
          __syncthreads();
        }
        unit = seg_end;
      }
    }
    MARK(4)
    bar.sync();
    MARK(5)

    // ------------------------------------------------------------------ P1
    float* yobs = a.yobs + (size_t)t * RH;
    {
      if (BF) { prefetch_l2(blk_log, bytes_log); if (!last) prefetch_l2(blk_hid, bytes_hid); }
      const int total = (last ? d.H : 2 * d.H) / 8;
      const int u0 = min(total, cta * per_ph1), u1 = min(total, u0 + per_ph1);
      auto aval = [&](int r, int k) -> float { return ldcg(deter + (size_t)r * d.D + k); };
      const uint4* dA = reinterpret_cast<const uint4*>(deterA + (size_t)(t & 1) * RD);
      for (int base = u0; base < u1; base += kMaxTiles) {
        const int nt = min(kMaxTiles, u1 - base);
        tile_gemm<ENG, true>(blk_ph1, per_ph1, base - u0, wf_ph1, base, nt, d.D, dA, aval, out, red_after(0));
        const int ncols = nt * 8;
        for (int i = tid; i < kRows * ncols; i += kThreads) {
          const int r = i / ncols, c = i - r * ncols;
          const int col = base * 8 + c;
          if (col < d.H) {
            yobs[(size_t)r * d.H + col] = out[i] + ldcg(a.pre_tok + (size_t)t * RH + (size_t)r * d.H + col);
          } else {
            const size_t at = (size_t)(t + 1) * RH + (size_t)r * d.H + col - d.H;
            a.y0[at] = ldcg(keep_next + r) * out[i] + a.b0[col - d.H];
            a.y1[at] = a.b1[col - d.H];          // P3 adds the sampled dynin1 rows on top
          }
        }
        __syncthreads();
      }
    }
    MARK(6)
    bar.sync();
    MARK(7)

    // ------------------------------------------------------------------ P2
    float* logit = a.logit + (size_t)t * RSC;
    {
      const int u0 = min(tiles_log, cta * per_log), u1 = min(tiles_log, u0 + per_log);
      if (u0 < u1) {
        row_rstd(yobs, d.H, d.eps, rstd_a);
        __syncthreads();
        if (cta == 0 && tid < kRows) a.rstd[(size_t)t * 3 * kRows + 2 * kRows + tid] = rstd_a[tid];
        auto aval = [&](int r, int k) -> float {
          return silu_f(ldcg(yobs + (size_t)r * d.H + k) * (rstd_a[r] * a.s_obs[k]));
        };
        if (BF) {
          build_part(afrag, 0, yobs, d.H, d.H, [&](int r, int k, float v) {
            return silu_fast(v * (rstd_a[r] * a.s_obs[k])); });
          __syncthreads();
        }
        for (int base = u0; base < u1; base += kMaxTiles) {
          const int nt = min(kMaxTiles, u1 - base);
          tile_gemm<ENG, false>(blk_log, per_log, base - u0, wf_log, base, nt, d.H, afrag4, aval, out, red_after(d.H));
          const int ncols = nt * 8;
          for (int i = tid; i < kRows * ncols; i += kThreads) {
            const int r = i / ncols, c = i - r * ncols;
            const int col = base * 8 + c;
            logit[(size_t)r * d.SC + col] = out[i] + a.b_logit[col];
          }
          __syncthreads();
        }
      }
    }
    if (!last) build_x0(a.y0 + (size_t)(t + 1) * RH, t + 1);
    MARK(8)
    bar.sync();
    MARK(9)

    // ------------------------------------------------------------------ P3
    // Sampling is split over the grid: one warp per (row, latent).  The winner's
    // dynin1 row is added straight into y1[t+1] (the one-hot matmul of
    // rssm.py:143 is a row gather).
    {
      const int warp = tid >> 5, lane = tid & 31;
      const int groups = d.B * d.S;
      const float* gum = a.gumbel + (size_t)t * RSC;
      for (int grp = cta * kWarps + warp; grp < groups; grp += ncta * kWarps) {
        const int r = grp / d.S, sv = grp - r * d.S;
        const float* l = logit + (size_t)r * d.SC + (size_t)sv * d.C;
        const float* gn = gum + (size_t)r * d.SC + (size_t)sv * d.C;
        // classes c = lane, lane+32, ... (C <= 128): all loads first, then the maths
        float lv[4], gv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = lane + 32 * i;
          lv[i] = c < d.C ? ldcg(l + c) : -INFINITY;
          gv[i] = c < d.C ? ldcg(gn + c) : 0.f;
        }
        float m = fmaxf(fmaxf(lv[0], lv[1]), fmaxf(lv[2], lv[3]));
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float e[4], z = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { e[i] = lane + 32 * i < d.C ? expf(lv[i] - m) : 0.f; z += e[i]; }
#pragma unroll
        for (int o = 16; o; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
        float best = -INFINITY;
        int arg = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = lane + 32 * i;
          if (c < d.C) {
            const float pr = e[i] / z;
            a.probs[(size_t)t * RSC + (size_t)r * d.SC + (size_t)sv * d.C + c] = pr;
            const float pm = (1.0f - d.unimix) * pr + d.unimix / (float)d.C;
            const float v = logf(pm) + gv[i];
            if (v > best) { best = v; arg = c; }
          }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
          if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (lane == 0) a.index[(size_t)t * kRows * d.S + grp] = arg;
        if (!last) {
          const float kn = ldcg(keep_next + r);
          if (kn != 0.f) {
            const size_t row = (size_t)sv * d.C + arg;
            float* dst = a.y1 + (size_t)(t + 1) * RH + (size_t)r * d.H;
            // all loads of the 2 KiB row first (it may have left L2), then the adds
            for (int c0 = 0; c0 < d.H; c0 += 32 * 32) {
              float wv[32];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int c = c0 + (j * 32 + lane) * 8;
                if (c < d.H) {
                  if (BF) {
                    const uint4 q = *reinterpret_cast<const uint4*>(
                        reinterpret_cast<const __nv_bfloat16*>(a.w_in1) + row * d.H + c);
                    const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w4[i]);
                      wv[j * 8 + 2 * i] = __low2float(h2); wv[j * 8 + 2 * i + 1] = __high2float(h2);
                    }
                  } else {
                    const float* src = reinterpret_cast<const float*>(a.w_in1) + row * d.H + c;
                    const float4 q0 = *reinterpret_cast<const float4*>(src);
                    const float4 q1 = *reinterpret_cast<const float4*>(src + 4);
                    wv[j * 8] = q0.x; wv[j * 8 + 1] = q0.y; wv[j * 8 + 2] = q0.z; wv[j * 8 + 3] = q0.w;
                    wv[j * 8 + 4] = q1.x; wv[j * 8 + 5] = q1.y; wv[j * 8 + 6] = q1.z; wv[j * 8 + 7] = q1.w;
                  }
                }
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int c = c0 + (j * 32 + lane) * 8;
                if (c < d.H) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) atomicAdd(dst + c + i, kn * wv[j * 8 + i]);
                }
              }
            }
          }
        }
      }
    }
    MARK(10)
    bar.sync();
    MARK(11)
  }
#undef MARK
}

size_t fwd_smem_bytes(const emb_rssm_fwd_args& a) {
  size_t n = sizeof(float) * (kRows * kMaxTiles * 8 + 2 * kRows);
  if (a.engine == rssm::ENG_F32) return n;
  auto cdiv = [](int x, int y) { return (x + y - 1) / y; };
  auto tiles = [&](int total, int unit, int groups) {     // n8 tiles one pass of this layer handles
    const int cpg = groups > 1 ? (a.ncta / groups > 1 ? a.ncta / groups : 1) : a.ncta;
    const int per = cdiv(total / groups / unit, cpg) * unit;
    return per < kMaxTiles ? per : (kMaxTiles / unit) * unit;
  };
  auto need = [&](int K, int nt) {
    return (size_t)kRows * K * 2 + (size_t)kWarps * kRows * nt * 8 * sizeof(float);
  };
  const int Dg = a.D / a.G, Kh = Dg + 3 * a.H;
  size_t m = need(Kh, tiles(a.D / 8, 1, a.G));
  size_t v = need(Dg, tiles(3 * a.D / 8, 3, a.G)); if (v > m) m = v;
  v = need(0, tiles(2 * a.H / 8, 1, 1)); if (v > m) m = v;
  v = need(a.H, tiles(a.S * a.C / 8, 1, 1)); if (v > m) m = v;
  return n + m;
}

int g_sms = 0;

}  // namespace

namespace emb_legacy {
size_t bwd_smem(const emb_rssm_bwd_args& a);     // rssm_bwd.cu
}
namespace emb_tma {
int launch_fwd(const emb_rssm_fwd_args& a, void* stream, bool dry);   // rssm_fwd_tma.cu
int launch_bwd(const emb_rssm_bwd_args& a, void* stream, bool dry);   // rssm_bwd_tma.cu
}

extern "C" int emb_rssm_observe_fwd(const emb_rssm_fwd_args* args, void* stream) {
  const char* who = "emb_rssm_observe_fwd";
  if (!args) return emb::fail(-1, "%s: args is NULL", who);
  const emb_rssm_fwd_args& a = *args;
  if (a.B < 1 || a.B > kRows) return emb::fail(-1, "%s: B=%d outside [1,16]", who, a.B);
  if (a.T < 1) return emb::fail(-1, "%s: T=%d < 1", who, a.T);
  if (a.C > 128) return emb::fail(-1, "%s: classes=%d > 128", who, a.C);
  if (a.G < 1 || a.D % a.G || (a.D / a.G) % 16 || a.H % 16 || (a.S * a.C) % 16 || a.D % 16)
    return emb::fail(-1, "%s: D/G, H and S*C must be multiples of 16 (D=%d G=%d H=%d S=%d C=%d)",
                     who, a.D, a.G, a.H, a.S, a.C);
  if (a.engine != rssm::ENG_F32 && a.engine != rssm::ENG_TMA && a.engine != rssm::ENG_LEGACY)
    return emb::fail(-1, "%s: engine %d", who, a.engine);
  if (a.engine == rssm::ENG_TMA) return emb_tma::launch_fwd(a, stream, false);
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return emb::fail_cuda(who);
  }
  if (a.ncta < 1 || a.ncta > g_sms)
    return emb::fail(-1, "%s: ncta=%d outside [1, %d SMs] (cooperative grid)", who, a.ncta, g_sms);
  const size_t smem = fwd_smem_bytes(a);
  if (smem > 227 * 1024)
    return emb::fail(-1, "%s: needs %zu bytes of shared memory (> 227 KiB)", who, smem);
  const void* fn = a.engine != rssm::ENG_F32 ? (const void*)rssm_fwd_kernel<rssm::ENG_BF16>
                                             : (const void*)rssm_fwd_kernel<rssm::ENG_F32>;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return emb::fail_cuda(who);
  emb_rssm_fwd_args copy = a;
  void* params[] = {&copy};
  if (cudaLaunchCooperativeKernel(fn, dim3(a.ncta), dim3(kThreads), params, smem,
                                  (cudaStream_t)stream) != cudaSuccess)
    return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}

/* 1 if the register-staged bf16 engine (engine 2) of both scan kernels fits (operands +
 * per-warp partial sums within 227 KiB of shared memory), else 0. */
extern "C" int emb_rssm_legacy_fits(int32_t D, int32_t H, int32_t S, int32_t C, int32_t G, int32_t ncta) {
  if (D < 16 || H < 16 || S < 1 || C < 1 || G < 1 || D % G || (D / G) % 16 || H % 16 || (S * C) % 16 || ncta < 1)
    return emb::fail(0, "emb_rssm_legacy_fits: D=%d H=%d S=%d C=%d G=%d", D, H, S, C, G);
  emb_rssm_fwd_args f = {};
  f.D = D; f.H = H; f.S = S; f.C = C; f.G = G; f.engine = rssm::ENG_LEGACY; f.ncta = ncta;
  emb_rssm_bwd_args b = {};
  b.D = D; b.H = H; b.S = S; b.C = C; b.G = G; b.engine = rssm::ENG_LEGACY; b.ncta = ncta; b.hoist_x2 = 1;
  const size_t cap = 227 * 1024;
  if (fwd_smem_bytes(f) > cap || emb_legacy::bwd_smem(b) > cap)
    return emb::fail(0, "emb_rssm_legacy_fits: operands of D=%d H=%d exceed 227 KiB of shared memory", D, H);
  return 1;
}

/* 1 if the bf16 TMA engine (engine 1) of both scan kernels fits this model on `ncta`
 * CTAs (tiles per CTA within the consumers' accumulators, operands + a weight ring of >= 2
 * stages within 227 KiB of shared memory), else 0 with the reason in emb_last_error(). */
extern "C" int emb_rssm_tma_fits(int32_t D, int32_t H, int32_t S, int32_t C, int32_t G, int32_t ncta) {
  if (D < 16 || H < 16 || S < 1 || C < 1 || G < 1 || D % G || (D / G) % 16 || H % 16 || (S * C) % 16)
    return emb::fail(0, "emb_rssm_tma_fits: D=%d H=%d S=%d C=%d G=%d", D, H, S, C, G);
  static float dummy_f = 0.f;
  emb_rssm_fwd_args f = {};
  f.B = 1; f.T = 1; f.D = D; f.H = H; f.S = S; f.C = C; f.G = G; f.engine = rssm::ENG_TMA; f.ncta = ncta;
  f.hid_pre = &dummy_f; f.sumsq_obs = &dummy_f;
  if (emb_tma::launch_fwd(f, nullptr, true) != 0) return 0;
  emb_rssm_bwd_args b = {};
  b.B = 1; b.T = 1; b.D = D; b.H = H; b.S = S; b.C = C; b.G = G; b.engine = rssm::ENG_TMA; b.ncta = ncta;
  b.hoist_x2 = 1; b.frag_scratch = &dummy_f; b.gx_part = &dummy_f;
  if (emb_tma::launch_bwd(b, nullptr, true) != 0) return 0;
  return 1;
}
