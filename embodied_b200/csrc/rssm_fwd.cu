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
rssm_fwd_kernel(const emb_rssm_fwd_args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Dims d = {a.B, a.T, a.D, a.H, a.S, a.C, a.G, a.D / a.G, a.S * a.C, a.unimix, a.eps};
  const int Kh = d.Dg + 3 * d.H;                 // dynhid0 input width per group
  // shared memory carve-up
  float* out = reinterpret_cast<float*>(smem_raw);                   // [16][kMaxTiles*8]
  float* rstd_a = out + kRows * kMaxTiles * 8;                       // [16]
  float* rstd_b = rstd_a + kRows;                                    // [16]
  int* sidx = reinterpret_cast<int*>(rstd_b + kRows);                // [16][S]
  __nv_bfloat16* afrag = reinterpret_cast<__nv_bfloat16*>(sidx + kRows * d.S + 16);
  afrag = reinterpret_cast<__nv_bfloat16*>(((uintptr_t)afrag + 15) & ~(uintptr_t)15);
  const uint4* afrag4 = reinterpret_cast<const uint4*>(afrag);

  GridBarrier bar{a.barrier, 0};
  const int tid = threadIdx.x;
  const size_t RH = (size_t)kRows * d.H, RD = (size_t)kRows * d.D, RSC = (size_t)kRows * d.SC;

  for (int t = 0; t < d.T; ++t) {
    const float* keep = a.keep + (size_t)t * kRows;
    const float* keep_next = a.keep + (size_t)(t + 1) * kRows;
    const float* deter_prev = t == 0 ? a.deter0 : a.deter + (size_t)(t - 1) * RD;
    const float* y0 = a.y0 + (size_t)t * RH;
    const float* y1 = a.y1 + (size_t)t * RH;
    const float* x2 = a.x2 + (size_t)t * RH;
    float* yhid = a.yhid + (size_t)t * RD;

    // ------------------------------------------------------------------ P4
    {
      int u0, u1;
      cta_range(d.D / 8, u0, u1);
      const int tpg = d.Dg / 8;                   // tiles per group
      if (u0 < u1) {
        row_rstd(y0, d.H, d.eps, rstd_a);
        row_rstd(y1, d.H, d.eps, rstd_b);
        __syncthreads();
      }
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
        if (ENG == ENG_BF16) build_afrag(afrag, Kh, aval);
        const char* wg = reinterpret_cast<const char*>(a.w_hid) +
            (size_t)g * Kh * d.Dg * (ENG == ENG_BF16 ? 2 : 4);
        for (int base = tile; base < seg_end; base += kMaxTiles) {
          const int nt = min(kMaxTiles, seg_end - base);
          tile_gemm<ENG, false>(wg, tpg, base - g * tpg, nt, Kh, afrag4, aval, out);
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
    bar.sync();

    // ------------------------------------------------------------------ P5
    float* deter = a.deter + (size_t)t * RD;
    {
      int u0, u1;
      cta_range(d.D / 8, u0, u1);                 // units of 8 deter columns (3 tiles each)
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
        if (ENG == ENG_BF16) build_afrag(afrag, d.Dg, aval);
        const char* wg = reinterpret_cast<const char*>(a.w_gru) +
            (size_t)g * d.Dg * 3 * d.Dg * (ENG == ENG_BF16 ? 2 : 4);
        constexpr int kMaxUnits = kMaxTiles / 3;
        for (int base = unit; base < seg_end; base += kMaxUnits) {
          const int nu = min(kMaxUnits, seg_end - base);
          tile_gemm<ENG, false>(wg, 3 * upg, (base - g * upg) * 3, nu * 3, d.Dg, afrag4, aval, out);
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
            const float cand = tanhf(rs * (o[8] + bg[d.Dg]));
            const float up = sigmoid_f(o[16] + bg[2 * d.Dg] - 1.0f);
            const float old = ldcg(keep + r) * ldcg(deter_prev + (size_t)r * d.D + col);
            const float nw = up * cand + (1.0f - up) * old;
            deter[(size_t)r * d.D + col] = nw;
            float* gs = a.gates + (size_t)t * 3 * RD + (size_t)r * d.D + col;
            gs[0] = rs; gs[RD] = cand; gs[2 * RD] = up;
            if (ENG == ENG_BF16) {
              __nv_bfloat16* da = reinterpret_cast<__nv_bfloat16*>(a.deterA) + (size_t)(t & 1) * RD;
              da[afrag_index(r, col)] = __float2bfloat16_rn(nw);
            }
          }
          __syncthreads();
        }
        unit = seg_end;
      }
    }
    bar.sync();

    // ------------------------------------------------------------------ P1
    float* yobs = a.yobs + (size_t)t * RH;
    {
      const bool last = t + 1 == d.T;
      int u0, u1;
      cta_range((last ? d.H : 2 * d.H) / 8, u0, u1);
      auto aval = [&](int r, int k) -> float { return ldcg(deter + (size_t)r * d.D + k); };
      const uint4* dA = reinterpret_cast<const uint4*>(
          reinterpret_cast<const __nv_bfloat16*>(a.deterA) + (size_t)(t & 1) * RD);
      for (int base = u0; base < u1; base += kMaxTiles) {
        const int nt = min(kMaxTiles, u1 - base);
        tile_gemm<ENG, true>(a.w_ph1, 2 * d.H / 8, base, nt, d.D, dA, aval, out);
        const int ncols = nt * 8;
        for (int i = tid; i < kRows * ncols; i += kThreads) {
          const int r = i / ncols, c = i - r * ncols;
          const int col = base * 8 + c;
          if (col < d.H) {
            yobs[(size_t)r * d.H + col] = out[i] + ldcg(a.pre_tok + (size_t)t * RH + (size_t)r * d.H + col);
          } else {
            a.y0[(size_t)(t + 1) * RH + (size_t)r * d.H + col - d.H] =
                ldcg(keep_next + r) * out[i] + a.b0[col - d.H];
          }
        }
        __syncthreads();
      }
    }
    bar.sync();

    // ------------------------------------------------------------------ P2
    float* logit = a.logit + (size_t)t * RSC;
    {
      int u0, u1;
      cta_range(d.SC / 8, u0, u1);
      if (u0 < u1) {
        row_rstd(yobs, d.H, d.eps, rstd_a);
        __syncthreads();
        auto aval = [&](int r, int k) -> float {
          return silu_f(ldcg(yobs + (size_t)r * d.H + k) * (rstd_a[r] * a.s_obs[k]));
        };
        if (ENG == ENG_BF16) build_afrag(afrag, d.H, aval);
        for (int base = u0; base < u1; base += kMaxTiles) {
          const int nt = min(kMaxTiles, u1 - base);
          tile_gemm<ENG, false>(a.w_logit, d.SC / 8, base, nt, d.H, afrag4, aval, out);
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
    bar.sync();

    // ------------------------------------------------------------------ P3
    {
      const bool last = t + 1 == d.T;
      int u0, u1;
      cta_range(d.H / 8, u0, u1);
      const bool writer = blockIdx.x == gridDim.x - 1;     // usually idle in the gather
      if ((u0 < u1 && !last) || writer) {
        // sample every (row, latent): warp-cooperative over the C classes
        const int warp = tid >> 5, lane = tid & 31;
        const float* gum = a.gumbel + (size_t)t * RSC;
        for (int grp = warp; grp < kRows * d.S; grp += kWarps) {
          const float* l = logit + (size_t)grp * d.C;
          const float* gn = gum + (size_t)grp * d.C;
          float m = -INFINITY;
          for (int c = lane; c < d.C; c += 32) m = fmaxf(m, ldcg(l + c));
#pragma unroll
          for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          float z = 0.f;
          for (int c = lane; c < d.C; c += 32) z += expf(ldcg(l + c) - m);
#pragma unroll
          for (int o = 16; o; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
          float best = -INFINITY;
          int arg = 0x7fffffff;
          for (int c = lane; c < d.C; c += 32) {
            const float p = expf(ldcg(l + c) - m) / z;
            const float pm = (1.0f - d.unimix) * p + d.unimix / (float)d.C;
            const float v = logf(pm) + ldcg(gn + c);
            if (v > best) { best = v; arg = c; }
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
          }
          if (lane == 0) sidx[grp] = arg;
        }
        __syncthreads();
        if (writer)
          for (int i = tid; i < kRows * d.S; i += kThreads)
            a.index[(size_t)t * kRows * d.S + i] = sidx[i];
        if (!last) {
          // y1' = keep' * sum_s dynin1[s*C + idx[r][s]][cols] + b1
          const int ncols = (u1 - u0) * 8;
          for (int i = tid; i < kRows * ncols; i += kThreads) {
            const int r = i / ncols, c = i - r * ncols;
            const int col = u0 * 8 + c;
            float s = 0.f;
            for (int sv = 0; sv < d.S; ++sv) {
              const size_t row = (size_t)sv * d.C + sidx[r * d.S + sv];
              if (ENG == ENG_BF16)
                s += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.w_in1)[row * d.H + col]);
              else
                s += reinterpret_cast<const float*>(a.w_in1)[row * d.H + col];
            }
            a.y1[(size_t)(t + 1) * RH + (size_t)r * d.H + col] = ldcg(keep_next + r) * s + a.b1[col];
          }
        }
      }
    }
    bar.sync();
  }
}

size_t fwd_smem_bytes(const emb_rssm_fwd_args& a) {
  const int Kh = a.D / a.G + 3 * a.H;
  int kmax = Kh > a.H ? Kh : a.H;
  if (a.D / a.G > kmax) kmax = a.D / a.G;
  size_t n = sizeof(float) * (kRows * kMaxTiles * 8 + 2 * kRows) + sizeof(int) * (kRows * a.S + 16) + 16;
  if (a.engine == rssm::ENG_BF16) n += (size_t)kRows * kmax * 2;
  return n;
}

int g_sms = 0;

}  // namespace

extern "C" int emb_rssm_observe_fwd(const emb_rssm_fwd_args* args, void* stream) {
  const char* who = "emb_rssm_observe_fwd";
  if (!args) return emb::fail(-1, "%s: args is NULL", who);
  const emb_rssm_fwd_args& a = *args;
  if (a.B < 1 || a.B > kRows) return emb::fail(-1, "%s: B=%d outside [1,16]", who, a.B);
  if (a.T < 1) return emb::fail(-1, "%s: T=%d < 1", who, a.T);
  if (a.G < 1 || a.D % a.G || (a.D / a.G) % 16 || a.H % 16 || (a.S * a.C) % 16 || a.D % 16)
    return emb::fail(-1, "%s: D/G, H and S*C must be multiples of 16 (D=%d G=%d H=%d S=%d C=%d)",
                     who, a.D, a.G, a.H, a.S, a.C);
  if (a.engine != rssm::ENG_F32 && a.engine != rssm::ENG_BF16)
    return emb::fail(-1, "%s: engine %d", who, a.engine);
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return emb::fail_cuda(who);
  }
  const size_t smem = fwd_smem_bytes(a);
  const void* fn = a.engine == rssm::ENG_BF16 ? (const void*)rssm_fwd_kernel<rssm::ENG_BF16>
                                              : (const void*)rssm_fwd_kernel<rssm::ENG_F32>;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return emb::fail_cuda(who);
  emb_rssm_fwd_args copy = a;
  void* params[] = {&copy};
  if (cudaLaunchCooperativeKernel(fn, dim3(g_sms), dim3(kThreads), params, smem,
                                  (cudaStream_t)stream) != cudaSuccess)
    return emb::fail_cuda(who);
  emb::count_launch();
  return 0;
}
