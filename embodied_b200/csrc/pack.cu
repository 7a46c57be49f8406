// Weight packing for the RSSM scan kernels (rssm_fwd_tma.cu / rssm_bwd_tma.cu):
// the fp32 master weights of the in-scan layers (dreamerv3/rssm.py:135-159, 81-86)
// are re-laid once per update into per-CTA bf16 blocks in mma.m16n8k16 B-fragment
// order, straight from the flat parameter buffer -- one pass, 4 bytes read and 2
// written per element, instead of a cast + concat + gather + permute chain.
//
//   dst[cta][kstep][tile][lane = nn*4 + kq][reg][half] =
//       bf16(base[off[cta*per + tile] + k * kstride[..] + nn * nstride[..]])
//       with k = kstep*16 + reg*8 + kq*2 + half, zero where off < 0 (padding tiles).
//
// A slot is one n8 column tile of the logical (K, N) matrix; (off, kstride,
// nstride) locate it in the flat buffer, which expresses block-diagonal layers,
// gate-interleaved column orders, column concatenations and transposes alike.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStepsPerWarp = 4;      // k-steps of one slot a warp converts

__global__ void __launch_bounds__(kThreads)
pack_tiles_kernel(const float* __restrict__ base, const int64_t* __restrict__ slot_off,
                  const int32_t* __restrict__ slot_ks, const int32_t* __restrict__ slot_ns,
                  __nv_bfloat16* __restrict__ dst, int64_t nslots, int per, int ks_begin,
                  int ks_count, int ks_total) {
  const int lane = threadIdx.x & 31;
  const int nn = lane >> 2, kq = lane & 3;
  const int chunks = (ks_count + kStepsPerWarp - 1) / kStepsPerWarp;
  const int64_t items = nslots * chunks;
  for (int64_t it = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); it < items;
       it += (int64_t)gridDim.x * kWarps) {
    // consecutive warps take consecutive slots of one k-step chunk: their output
    // tiles are adjacent in dst
    const int64_t chunk = it / nslots, slot = it - chunk * nslots;
    const int64_t cta = slot / per, tile = slot - cta * per;
    const int64_t off = slot_off[slot];
    const int64_t ks = slot_ks[slot], ns = slot_ns[slot];
    const float* src = base + off + nn * ns + (int64_t)(kq * 2) * ks;
    const int k0 = (int)chunk * kStepsPerWarp;
#pragma unroll
    for (int j = 0; j < kStepsPerWarp; ++j) {
      const int kstep = k0 + j;
      if (kstep >= ks_count) break;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (off >= 0) {
        const float* p = src + (int64_t)kstep * 16 * ks;
        v[0] = __ldg(p); v[1] = __ldg(p + ks);
        v[2] = __ldg(p + 8 * ks); v[3] = __ldg(p + 9 * ks);
      }
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
      const __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
      uint2 w;
      w.x = *reinterpret_cast<const uint32_t*>(&lo);
      w.y = *reinterpret_cast<const uint32_t*>(&hi);
      __nv_bfloat16* out = dst + (((cta * ks_total + ks_begin + kstep) * per + tile) * 32 + lane) * 4;
      *reinterpret_cast<uint2*>(out) = w;
    }
  }
}

int g_sms = 0;

}  // namespace

extern "C" int emb_pack_tiles(const float* base, const int64_t* slot_off, const int32_t* slot_kstride,
                              const int32_t* slot_nstride, void* dst, int64_t nslots, int32_t per,
                              int32_t ks_begin, int32_t ks_count, int32_t ks_total, void* stream) {
  const char* who = "emb_pack_tiles";
  if (nslots < 0 || per <= 0 || nslots % per || ks_begin < 0 || ks_count < 0 ||
      ks_begin + ks_count > ks_total)
    return emb::fail(-1, "%s: nslots=%lld per=%d ksteps [%d, +%d) of %d", who,
                     (long long)nslots, per, ks_begin, ks_count, ks_total);
  if (nslots == 0 || ks_count == 0) return 0;
  if (!base || !slot_off || !slot_kstride || !slot_nstride || !dst)
    return emb::fail(-1, "%s: null pointer", who);
  if (!g_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int64_t items = nslots * ((ks_count + kStepsPerWarp - 1) / kStepsPerWarp);
  const int64_t want = (items + kWarps - 1) / kWarps, cap = (int64_t)g_sms * 32;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  pack_tiles_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
      base, slot_off, slot_kstride, slot_nstride, reinterpret_cast<__nv_bfloat16*>(dst), nslots,
      per, ks_begin, ks_count, ks_total);
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}
