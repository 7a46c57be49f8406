// Row engine: the one HBM-bound kernel behind Replay gather / append / update
// and the Driver's stage-obs / mask-actions (include/embodied_b200.h).
//
// A launch moves `nrows` rows of up to EMB_MAX_KEYS keys.  Keys are split by
// row size:
//   big   (row_bytes >= 256, 16-byte aligned, plain copy or u8->f32 normalise):
//         work unit = (row, key, 16 KiB slice) handled by one 256-thread CTA
//         iteration, four 16-byte loads in flight per thread before any store;
//   small (everything else: flags, rewards, stepid, consec, actions):
//         work unit = 1024 consecutive (row, vector) elements of one key.
// The grid is persistent (SMs x 8 CTAs) and strides over the unit list, so a
// default dreamerv3 batch (16x65 rows x {image 12 KiB, deter 32 KiB, stoch 8 KiB}
// + seven small keys) is ~4.2k big units + ~10 small units in ONE launch.
//
// Everything here is byte/integer exact; the only arithmetic is the typed
// multiply of Driver._mask and float(u8)/255-0.5 (IEEE fp32 divide, no FMA).
//
// Two engines for the big units:
//   rows_tma_kernel (default): one elected lane per CTA drives a 4-stage shared-memory ring with
//     TMA bulk copies -- cp.async.bulk global -> shared (mbarrier complete_tx), then shared ->
//     global (bulk_group) -- two loads and up to two stores in flight per CTA, three CTAs per SM
//     (192 KiB in flight per SM without a single register); the other seven warps convert the
//     u8 -> f32 normalised copy out of the staged tile (stage_obs) and run the small units;
//   rows_kernel: register-staged ld.global.nc / st.global (EMB_ROWS_TMA=0).

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "../../include/embodied_b200.h"
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr uint32_t kBigSlice = 16384;        // bytes per big unit
constexpr uint32_t kSmallElems = 1024;       // vector elements per small unit
constexpr uint32_t kBigMin = 256;

struct DevKey {
  const uint8_t* src;
  uint8_t* dst;
  uint8_t* dst2;
  const uint8_t* aux;
  uint64_t src_stride, dst_stride, dst2_stride, aux_stride;
  uint32_t row_bytes;
  uint32_t op;
  uint32_t dtype;
  int32_t fill;
  uint32_t unit_begin;   // big: first slice index within a row; small: first unit
  uint32_t vec;          // small: bytes per vector element; big: 16
  uint32_t vecs_per_row; // small: row_bytes / vec
  uint32_t pad;
};

struct Table {
  uint32_t nbig, nsmall;
  uint32_t big_units_per_row;
  uint32_t small_units;
  DevKey k[EMB_MAX_KEYS];   // big keys first, then small keys
};

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float norm_u8(uint32_t b) {
  // dreamerv3/rssm.py:230  x.astype(f32) / 255 - 0.5, IEEE round-to-nearest.
  return __fsub_rn(__fdiv_rn((float)b, 255.0f), 0.5f);
}

__device__ __forceinline__ void store_norm16(float* out, const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 f;
    f.x = norm_u8(w[i] & 0xff);
    f.y = norm_u8((w[i] >> 8) & 0xff);
    f.z = norm_u8((w[i] >> 16) & 0xff);
    f.w = norm_u8(w[i] >> 24);
    reinterpret_cast<float4*>(out)[i] = f;
  }
}

template <typename T>
__device__ __forceinline__ void mask_mul(uint8_t* d, const uint8_t* s, bool keep) {
  // driver.py:84-87  value * mask.astype(value.dtype)
  T v = *reinterpret_cast<const T*>(s);
  T m = keep ? T(1) : T(0);
  *reinterpret_cast<T*>(d) = v * m;
}

template <typename U>
__device__ __forceinline__ U float_bits_times_zero(U v, int ebits, int mbits) {
  const U one = 1;
  const U sign = one << (ebits + mbits);
  const U emask = ((one << ebits) - 1) << mbits;
  const U mmask = (one << mbits) - 1;
  const U quiet = one << (mbits - 1);
  if ((v & emask) == emask) {
    if (v & mmask) return v | quiet;          // NaN in -> same NaN, quieted
    return sign | emask | quiet;              // inf * 0 -> default NaN
  }
  return v & sign;                            // finite * (+0) -> +-0
}

__device__ __forceinline__ void mask_elem(uint32_t dtype, uint8_t* d,
                                          const uint8_t* s, bool keep) {
  switch (dtype) {
    case EMB_U8: case EMB_I8: mask_mul<uint8_t>(d, s, keep); break;
    case EMB_BOOL: *d = (*s && keep) ? 1 : 0; break;
    case EMB_I16: case EMB_U16: mask_mul<uint16_t>(d, s, keep); break;
    case EMB_I32: case EMB_U32: mask_mul<uint32_t>(d, s, keep); break;
    case EMB_I64: case EMB_U64: mask_mul<unsigned long long>(d, s, keep); break;
    // Floating point: value * mask.astype(dtype) with mask in {0, 1}.  Done on
    // the bit pattern so that the result is what numpy on the reference's x86
    // host produces, byte for byte: x*1 = x; finite*0 = +-0 (sign kept);
    // inf*0 = the x86 default NaN (sign set, quiet); nan*0 = the input NaN quieted.
    case EMB_F32: {
      uint32_t v = *reinterpret_cast<const uint32_t*>(s);
      *reinterpret_cast<uint32_t*>(d) =
          keep ? v : float_bits_times_zero<uint32_t>(v, 8, 23);
    } break;
    case EMB_F64: {
      unsigned long long v = *reinterpret_cast<const unsigned long long*>(s);
      *reinterpret_cast<unsigned long long*>(d) =
          keep ? v : float_bits_times_zero<unsigned long long>(v, 11, 52);
    } break;
    case EMB_F16: {
      uint16_t v = *reinterpret_cast<const uint16_t*>(s);
      *reinterpret_cast<uint16_t*>(d) =
          keep ? v : float_bits_times_zero<uint16_t>(v, 5, 10);
    } break;
    case EMB_BF16: {
      uint16_t v = *reinterpret_cast<const uint16_t*>(s);
      *reinterpret_cast<uint16_t*>(d) =
          keep ? v : float_bits_times_zero<uint16_t>(v, 8, 7);
    } break;
    default: break;
  }
}

__device__ __forceinline__ uint32_t dtype_size(uint32_t dtype) {
  switch (dtype) {
    case EMB_U8: case EMB_I8: case EMB_BOOL: return 1;
    case EMB_I16: case EMB_U16: case EMB_F16: case EMB_BF16: return 2;
    case EMB_I32: case EMB_U32: case EMB_F32: return 4;
    default: return 8;
  }
}

// ---------------- small path: 1024 (row, vector) elements of one key, by `nthreads` threads
__device__ __forceinline__ void small_unit(const Table& tab, uint32_t su, const int64_t* src_rows,
                                           const int64_t* dst_rows, int64_t nrows, int32_t window,
                                           uint32_t tid, uint32_t nthreads) {
  {
    {
      uint32_t ki = tab.nbig;
      const uint32_t kend = tab.nbig + tab.nsmall;
      while (ki + 1 < kend && tab.k[ki + 1].unit_begin <= su) ++ki;
      const DevKey& key = tab.k[ki];
      const uint64_t base = (uint64_t)(su - key.unit_begin) * kSmallElems;
      const uint64_t nelem = (uint64_t)nrows * key.vecs_per_row;
#pragma unroll 1
      for (uint32_t el = tid; el < kSmallElems; el += nthreads) {
        const uint64_t e = base + el;
        if (e >= nelem) break;
        const int64_t r = (int64_t)(e / key.vecs_per_row);
        const uint32_t c = (uint32_t)(e % key.vecs_per_row) * key.vec;
        const int64_t sr = src_rows ? src_rows[r] : r;
        const int64_t dr = dst_rows ? dst_rows[r] : r;
        if (sr < 0 || dr < 0) continue;
        uint8_t* d = key.dst ? key.dst + (uint64_t)dr * key.dst_stride + c : nullptr;
        uint8_t* d2 = key.dst2 ? key.dst2 + (uint64_t)r * key.dst2_stride + c : nullptr;
        const uint8_t* s = key.src ? key.src + (uint64_t)sr * key.src_stride + c : nullptr;
        uint8_t tmp[16];
        switch (key.op) {
          case EMB_OP_FILL32:
            *reinterpret_cast<int32_t*>(tmp) = key.fill;
            break;
          case EMB_OP_MASK: {
            const bool keep = key.aux[(uint64_t)r * key.aux_stride] == 0;
            const uint32_t es = dtype_size(key.dtype);
            for (uint32_t b = 0; b < key.vec; b += es)
              mask_elem(key.dtype, tmp + b, s + b, keep);
          } break;
          case EMB_OP_NOT:
            for (uint32_t b = 0; b < key.vec; ++b) tmp[b] = s[b] ? 0 : 1;
            break;
          default:
            switch (key.vec) {
              case 16: *reinterpret_cast<uint4*>(tmp) = *reinterpret_cast<const uint4*>(s); break;
              case 8: *reinterpret_cast<uint2*>(tmp) = *reinterpret_cast<const uint2*>(s); break;
              case 4: *reinterpret_cast<uint32_t*>(tmp) = *reinterpret_cast<const uint32_t*>(s); break;
              case 2: *reinterpret_cast<uint16_t*>(tmp) = *reinterpret_cast<const uint16_t*>(s); break;
              default: tmp[0] = *s; break;
            }
            if (c == 0 && window > 0) {
              const int32_t tpos = (int32_t)(r % window);
              if (key.op == EMB_OP_FIRST && tpos == 0) {
                tmp[0] = 1;                                   // replay.py:285-286
              } else if (key.op == EMB_OP_LAST && tpos + 1 < window) {
                const int64_t nr = src_rows ? src_rows[r + 1] : r + 1;
                if (nr >= 0 && key.aux[(uint64_t)nr * key.aux_stride])
                  tmp[0] = 1;                                 // replay.py:288-291
              }
            }
            break;
        }
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
          uint8_t* o = which ? d2 : d;
          if (!o) continue;
          switch (key.vec) {
            case 16: *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(tmp); break;
            case 8: *reinterpret_cast<uint2*>(o) = *reinterpret_cast<uint2*>(tmp); break;
            case 4: *reinterpret_cast<uint32_t*>(o) = *reinterpret_cast<uint32_t*>(tmp); break;
            case 2: *reinterpret_cast<uint16_t*>(o) = *reinterpret_cast<uint16_t*>(tmp); break;
            default: *o = tmp[0]; break;
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads)
rows_kernel(const __grid_constant__ Table tab,
            const int64_t* __restrict__ src_rows,
            const int64_t* __restrict__ dst_rows,
            int64_t nrows, int32_t window) {
  const uint64_t big_units = (uint64_t)nrows * tab.big_units_per_row;
  const uint64_t total = big_units + tab.small_units;
  for (uint64_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    if (unit < big_units) {
      // ---------------- big path: one 16 KiB slice of one row of one key
      const int64_t r = (int64_t)(unit / tab.big_units_per_row);
      const uint32_t j = (uint32_t)(unit % tab.big_units_per_row);
      uint32_t ki = 0;
      while (ki + 1 < tab.nbig && tab.k[ki + 1].unit_begin <= j) ++ki;
      const DevKey& key = tab.k[ki];
      const int64_t sr = src_rows ? src_rows[r] : r;
      const int64_t dr = dst_rows ? dst_rows[r] : r;
      if (sr < 0 || dr < 0) continue;
      const uint32_t off = (j - key.unit_begin) * kBigSlice;
      const uint32_t nvec = (min(key.row_bytes - off, kBigSlice)) >> 4;
      const uint4* s = reinterpret_cast<const uint4*>(
          key.src + (uint64_t)sr * key.src_stride + off);
      uint4* d = key.dst ? reinterpret_cast<uint4*>(
          key.dst + (uint64_t)dr * key.dst_stride + off) : nullptr;
      uint8_t* d2 = key.dst2 ? key.dst2 + (uint64_t)r * key.dst2_stride : nullptr;
      const bool norm = key.op == EMB_OP_NORM_U8_F32;
      // 4 independent 16-byte loads per thread before the stores (64 B/thread,
      // 16 KiB per CTA iteration in flight).
      uint4 v[4];
      const uint32_t t = threadIdx.x;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t i = t + u * kThreads;
        if (i < nvec) v[u] = ld_stream(s + i);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t i = t + u * kThreads;
        if (i < nvec) {
          if (d) st_stream(d + i, v[u]);
          if (d2) {
            if (norm) {
              store_norm16(reinterpret_cast<float*>(d2) + (off + i * 16), v[u]);
            } else {
              st_stream(reinterpret_cast<uint4*>(d2 + off) + i, v[u]);
            }
          }
        }
      }
    } else {
      small_unit(tab, (uint32_t)(unit - big_units), src_rows, dst_rows, nrows, window, threadIdx.x,
                 kThreads);
    }
  }
}

// ------------------------------------------------------------------ TMA engine
constexpr int kStagesT = 4;                  // 16 KiB stages per CTA
constexpr int kAhead = 2;                    // loads issued ahead of the store cursor
constexpr int kWorkers = kThreads - 32;      // warps 1..7

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
      "RW_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra RW_DONE;\n"
      "bra RW_LOOP;\n"
      "RW_DONE:\n"
      "}\n" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}

struct BigUnit {        // one 16 KiB slice of one row of one big key
  const uint8_t* src;
  uint8_t* dst;
  uint8_t* dst2;
  uint32_t bytes;
  uint32_t off;
  bool norm, live;
};

__device__ __forceinline__ BigUnit big_unit(const Table& tab, uint64_t unit, const int64_t* src_rows,
                                            const int64_t* dst_rows) {
  BigUnit u;
  const int64_t r = (int64_t)(unit / tab.big_units_per_row);
  const uint32_t j = (uint32_t)(unit % tab.big_units_per_row);
  uint32_t ki = 0;
  while (ki + 1 < tab.nbig && tab.k[ki + 1].unit_begin <= j) ++ki;
  const DevKey& key = tab.k[ki];
  const int64_t sr = src_rows ? src_rows[r] : r;
  const int64_t dr = dst_rows ? dst_rows[r] : r;
  u.live = sr >= 0 && dr >= 0;
  u.off = (j - key.unit_begin) * kBigSlice;
  u.bytes = min(key.row_bytes - u.off, kBigSlice);
  u.norm = key.op == EMB_OP_NORM_U8_F32;
  u.src = key.src + (uint64_t)(u.live ? sr : 0) * key.src_stride + u.off;
  u.dst = key.dst ? key.dst + (uint64_t)(u.live ? dr : 0) * key.dst_stride + u.off : nullptr;
  u.dst2 = key.dst2 ? key.dst2 + (uint64_t)r * key.dst2_stride : nullptr;
  return u;
}

__global__ void __launch_bounds__(kThreads)
rows_tma_kernel(const __grid_constant__ Table tab, const int64_t* __restrict__ src_rows,
                const int64_t* __restrict__ dst_rows, int64_t nrows, int32_t window, int has_norm) {
  extern __shared__ __align__(128) uint8_t ring[];          // kStagesT x 16 KiB, then the barriers
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + kStagesT * kBigSlice);
  uint64_t* wdone = full + kStagesT;                         // workers finished converting a stage
  const uint64_t big_units = (uint64_t)nrows * tab.big_units_per_row;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStagesT; ++i) { mbar_init(&full[i], 1); mbar_init(&wdone[i], kWorkers / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // this CTA's big units: blockIdx.x, blockIdx.x + gridDim.x, ...
  const uint64_t mine = big_units > blockIdx.x ? (big_units - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x == 0) {
    // ---------------------------------------------------------- the copy engine (one lane)
    uint64_t loaded = 0;
    for (uint64_t k = 0; k < mine; ++k) {
      // keep kAhead loads in front of the store cursor; a stage is reused once the bulk store
      // that read it has drained (at most kStagesT - kAhead - 1 younger stores pending) and,
      // for normalised keys, once the workers are done with it
      while (loaded < mine && loaded < k + kAhead) {
        const int st = (int)(loaded % kStagesT);
        if (loaded >= kStagesT) {
          bulk_wait_read<kStagesT - kAhead - 1>();
          if (has_norm) mbar_wait(&wdone[st], (uint32_t)((loaded / kStagesT - 1) & 1));
        }
        const BigUnit u = big_unit(tab, blockIdx.x + loaded * gridDim.x, src_rows, dst_rows);
        if (u.live) {
          mbar_expect_tx(&full[st], u.bytes);
          bulk_g2s(ring + st * kBigSlice, u.src, u.bytes, &full[st]);
        } else {
          mbar_arrive(&full[st]);                            // evicted row: nothing to move
        }
        ++loaded;
      }
      const int st = (int)(k % kStagesT);
      const BigUnit u = big_unit(tab, blockIdx.x + k * gridDim.x, src_rows, dst_rows);
      mbar_wait(&full[st], (uint32_t)((k / kStagesT) & 1));
      if (u.live) {
        if (u.dst) bulk_s2g(u.dst, ring + st * kBigSlice, u.bytes);
        if (u.dst2 && !u.norm) bulk_s2g(u.dst2 + u.off, ring + st * kBigSlice, u.bytes);
      }
      bulk_commit();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // every store has completed
  } else if (threadIdx.x >= 32) {
    const uint32_t tid = threadIdx.x - 32;
    if (has_norm) {
      // ------------------------------------------------ workers: u8 -> f32 out of the staged tile
      for (uint64_t k = 0; k < mine; ++k) {
        const int st = (int)(k % kStagesT);
        const BigUnit u = big_unit(tab, blockIdx.x + k * gridDim.x, src_rows, dst_rows);
        mbar_wait(&full[st], (uint32_t)((k / kStagesT) & 1));
        if (u.live && u.norm && u.dst2) {
          const uint4* tile = reinterpret_cast<const uint4*>(ring + st * kBigSlice);
          float* out = reinterpret_cast<float*>(u.dst2) + u.off;
          for (uint32_t i = tid; i < (u.bytes >> 4); i += kWorkers) store_norm16(out + i * 16, tile[i]);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&wdone[st]);
      }
    }
    // ---------------------------------------------------------------- small units
    for (uint64_t su = blockIdx.x; su < tab.small_units; su += gridDim.x)
      small_unit(tab, (uint32_t)su, src_rows, dst_rows, nrows, window, tid, kWorkers);
  }
}

uint32_t pow2_divisor(uint64_t x, uint32_t cap) {
  uint32_t v = 1;
  while (v < cap && (x % (v * 2)) == 0) v *= 2;
  return v;
}

int g_sm_count = 0;

int launch(const emb_key_t* keys, int nkeys, const int64_t* src_rows,
           const int64_t* dst_rows, int64_t nrows, int32_t window, void* stream,
           uint32_t allowed_ops, const char* who) {
  if (nkeys < 0 || nkeys > EMB_MAX_KEYS)
    return emb::fail(-1, "%s: nkeys=%d outside [0,%d]", who, nkeys, EMB_MAX_KEYS);
  if (nrows < 0) return emb::fail(-1, "%s: nrows=%lld < 0", who, (long long)nrows);
  if (nrows == 0 || nkeys == 0) return 0;
  if (!keys) return emb::fail(-1, "%s: keys is NULL", who);
  Table tab;
  memset(&tab, 0, sizeof(tab));
  DevKey big[EMB_MAX_KEYS], small[EMB_MAX_KEYS];
  uint32_t nbig = 0, nsmall = 0, big_units = 0;
  uint64_t small_units = 0;
  for (int i = 0; i < nkeys; ++i) {
    const emb_key_t& k = keys[i];
    if (!((allowed_ops >> k.op) & 1u))
      return emb::fail(-2, "%s: key %d has op %u not valid for this entry point",
                       who, i, k.op);
    if (k.row_bytes == 0) continue;
    if (!k.dst && !k.dst2)
      return emb::fail(-2, "%s: key %d has no destination", who, i);
    if (!k.src && k.op != EMB_OP_FILL32)
      return emb::fail(-2, "%s: key %d has no source", who, i);
    if ((k.op == EMB_OP_LAST || k.op == EMB_OP_MASK) && !k.aux)
      return emb::fail(-2, "%s: key %d (op %u) needs aux", who, i, k.op);
    if (k.op == EMB_OP_NORM_U8_F32 && !k.dst2)
      return emb::fail(-2, "%s: key %d NORM needs dst2", who, i);
    if (k.op == EMB_OP_FILL32 && k.row_bytes != 4)
      return emb::fail(-2, "%s: key %d FILL32 needs row_bytes==4", who, i);
    DevKey d;
    memset(&d, 0, sizeof(d));
    d.src = (const uint8_t*)k.src; d.dst = (uint8_t*)k.dst;
    d.dst2 = (uint8_t*)k.dst2; d.aux = (const uint8_t*)k.aux;
    d.src_stride = k.src_stride; d.dst_stride = k.dst_stride;
    d.dst2_stride = k.dst2_stride; d.aux_stride = k.aux_stride;
    d.row_bytes = k.row_bytes; d.op = k.op; d.dtype = k.dtype; d.fill = k.fill;
    // alignment every pointer/stride of this key shares
    uint64_t mix = k.row_bytes | 16;
    if (k.src) mix |= (uint64_t)k.src | k.src_stride;
    if (k.dst) mix |= (uint64_t)k.dst | k.dst_stride;
    const bool norm = k.op == EMB_OP_NORM_U8_F32;
    if (k.dst2) mix |= (uint64_t)k.dst2 | (norm ? 16 : k.dst2_stride);
    uint32_t vec = pow2_divisor(mix, 16);
    if (norm && (k.dst2_stride % 16 || ((uint64_t)k.dst2 % 16)))
      return emb::fail(-2, "%s: key %d NORM dst2 must be 16-byte aligned", who, i);
    const bool plain = k.op == EMB_OP_COPY || norm;
    if (norm && vec != 16)
      return emb::fail(-2, "%s: key %d NORM needs 16-byte aligned rows", who, i);
    if (plain && vec == 16 && (k.row_bytes >= kBigMin || norm)) {
      d.vec = 16;
      d.unit_begin = big_units;
      big_units += (k.row_bytes + kBigSlice - 1) / kBigSlice;
      big[nbig++] = d;
    } else {
      if (k.op == EMB_OP_MASK) {
        uint32_t es = 8;
        switch (k.dtype) {
          case EMB_U8: case EMB_I8: case EMB_BOOL: es = 1; break;
          case EMB_I16: case EMB_U16: case EMB_F16: case EMB_BF16: es = 2; break;
          case EMB_I32: case EMB_U32: case EMB_F32: es = 4; break;
          default: es = 8; break;
        }
        if (k.row_bytes % es)
          return emb::fail(-2, "%s: key %d row_bytes %% elem size", who, i);
        if (vec < es) vec = es;   // element-aligned by construction of the dtype
      }
      if (k.op == EMB_OP_FILL32) vec = 4;
      d.vec = vec;
      d.vecs_per_row = k.row_bytes / vec;
      d.unit_begin = (uint32_t)small_units;
      small_units += ((uint64_t)nrows * d.vecs_per_row + kSmallElems - 1) / kSmallElems;
      small[nsmall++] = d;
    }
  }
  if (small_units > 0xffffffffull)
    return emb::fail(-3, "%s: too many small units", who);
  tab.nbig = nbig; tab.nsmall = nsmall;
  tab.big_units_per_row = big_units; tab.small_units = (uint32_t)small_units;
  for (uint32_t i = 0; i < nbig; ++i) tab.k[i] = big[i];
  for (uint32_t i = 0; i < nsmall; ++i) tab.k[nbig + i] = small[i];
  const uint64_t total = (uint64_t)nrows * big_units + small_units;
  if (total == 0) return 0;
  if (g_sm_count == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return emb::fail_cuda(who);
    g_sm_count = sms;
  }
  static int use_tma = -1;
  if (use_tma < 0) {
    const char* e = getenv("EMB_ROWS_TMA");
    use_tma = (e && e[0] == '0') ? 0 : 1;
  }
  if (use_tma && nbig > 0) {
    // TMA engine: three CTAs per SM (64 KiB ring each); every big slice is a 16-byte multiple
    // (vec == 16 was required to classify the key as big)
    int has_norm = 0;
    for (uint32_t i = 0; i < nbig; ++i) has_norm |= big[i].op == EMB_OP_NORM_U8_F32;
    const size_t smem = (size_t)kStagesT * kBigSlice + 2 * kStagesT * sizeof(uint64_t);
    static bool attr = false;
    if (!attr) {
      if (cudaFuncSetAttribute((const void*)rows_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem) != cudaSuccess)
        return emb::fail_cuda(who);
      attr = true;
    }
    uint64_t grid = (uint64_t)g_sm_count * 3;
    const uint64_t big_total = (uint64_t)nrows * big_units;
    const uint64_t want = big_total > small_units ? big_total : small_units;
    if (want < grid) grid = want;
    rows_tma_kernel<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(
        tab, src_rows, dst_rows, nrows, window, has_norm);
    emb::count_launch();
    if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
    return 0;
  }
  // persistent grid: a multiple of the SM count (8 x 256 threads = full SM)
  uint64_t grid = (uint64_t)g_sm_count * 8;
  if (total < grid) grid = total;
  rows_kernel<<<(unsigned)grid, kThreads, 0, (cudaStream_t)stream>>>(
      tab, src_rows, dst_rows, nrows, window);
  emb::count_launch();
  if (cudaPeekAtLastError() != cudaSuccess) return emb::fail_cuda(who);
  return 0;
}

constexpr uint32_t bit(uint32_t op) { return 1u << op; }

}  // namespace

extern "C" {

int emb_rows_copy(const emb_key_t* keys, int nkeys, const int64_t* src_rows,
                  const int64_t* dst_rows, int64_t nrows, int32_t window,
                  void* stream) {
  return launch(keys, nkeys, src_rows, dst_rows, nrows, window, stream,
                0xffffffffu, "emb_rows_copy");
}

int emb_replay_gather(const emb_key_t* keys, int nkeys, const int64_t* src_rows,
                      int64_t nrows, int32_t window, void* stream) {
  if (!src_rows) return emb::fail(-1, "emb_replay_gather: src_rows is NULL");
  if (window <= 0 || nrows % window)
    return emb::fail(-1, "emb_replay_gather: nrows=%lld not a multiple of window=%d",
                     (long long)nrows, window);
  return launch(keys, nkeys, src_rows, nullptr, nrows, window, stream,
                bit(EMB_OP_COPY) | bit(EMB_OP_FIRST) | bit(EMB_OP_LAST) |
                bit(EMB_OP_FILL32), "emb_replay_gather");
}

int emb_replay_append_rows(const emb_key_t* keys, int nkeys,
                           const int64_t* dst_rows, int64_t nrows, void* stream) {
  if (!dst_rows) return emb::fail(-1, "emb_replay_append_rows: dst_rows is NULL");
  return launch(keys, nkeys, nullptr, dst_rows, nrows, 0, stream,
                bit(EMB_OP_COPY), "emb_replay_append_rows");
}

int emb_replay_scatter_update(const emb_key_t* keys, int nkeys,
                              const int64_t* dst_rows, int64_t nrows, void* stream) {
  if (!dst_rows) return emb::fail(-1, "emb_replay_scatter_update: dst_rows is NULL");
  return launch(keys, nkeys, nullptr, dst_rows, nrows, 0, stream,
                bit(EMB_OP_COPY), "emb_replay_scatter_update");
}

int emb_driver_stage_obs(const emb_key_t* keys, int nkeys, const int64_t* dst_rows,
                         int64_t nrows, void* stream) {
  return launch(keys, nkeys, nullptr, dst_rows, nrows, 0, stream,
                bit(EMB_OP_COPY) | bit(EMB_OP_NORM_U8_F32),
                "emb_driver_stage_obs");
}

int emb_driver_scatter_mask_actions(const emb_key_t* keys, int nkeys,
                                    const int64_t* dst_rows, int64_t nrows,
                                    void* stream) {
  return launch(keys, nkeys, nullptr, dst_rows, nrows, 0, stream,
                bit(EMB_OP_COPY) | bit(EMB_OP_MASK) | bit(EMB_OP_NOT),
                "emb_driver_scatter_mask_actions");
}

}  // extern "C"

// Chunk.save / load: a slab's rows are consecutive rows of every table (embodied/core/chunk.py:64-99)
static int chunk_copy(const emb_key_t* keys, int nkeys, int64_t row0, int64_t nrows, void* stream,
                      bool to_table, const char* who) {
  if (nkeys < 0 || nkeys > EMB_MAX_KEYS)
    return emb::fail(-1, "%s: nkeys=%d outside [0,%d]", who, nkeys, EMB_MAX_KEYS);
  if (row0 < 0 || nrows < 0) return emb::fail(-1, "%s: row0=%lld nrows=%lld", who, (long long)row0, (long long)nrows);
  if (nrows == 0 || nkeys == 0) return 0;
  if (!keys) return emb::fail(-1, "%s: keys is NULL", who);
  for (int i = 0; i < nkeys; ++i) {
    const emb_key_t& k = keys[i];
    if (k.row_bytes == 0) continue;
    if (!k.src || !k.dst) return emb::fail(-2, "%s: key %d needs src and dst", who, i);
    const size_t sp = k.src_stride ? k.src_stride : k.row_bytes, dp = k.dst_stride ? k.dst_stride : k.row_bytes;
    if (sp < k.row_bytes || dp < k.row_bytes) return emb::fail(-2, "%s: key %d has a pitch below its row size", who, i);
    const unsigned char* src = (const unsigned char*)k.src + (to_table ? 0 : (size_t)row0 * sp);
    unsigned char* dst = (unsigned char*)k.dst + (to_table ? (size_t)row0 * dp : 0);
    if (cudaMemcpy2DAsync(dst, dp, src, sp, k.row_bytes, (size_t)nrows, cudaMemcpyDefault,
                          (cudaStream_t)stream) != cudaSuccess)
      return emb::fail_cuda(who);
  }
  return 0;
}

extern "C" int emb_replay_export_chunk(const emb_key_t* keys, int nkeys, int64_t row0, int64_t nrows,
                                       void* stream) {
  return chunk_copy(keys, nkeys, row0, nrows, stream, false, "emb_replay_export_chunk");
}

extern "C" int emb_replay_import_chunk(const emb_key_t* keys, int nkeys, int64_t row0, int64_t nrows,
                                       void* stream) {
  return chunk_copy(keys, nkeys, row0, nrows, stream, true, "emb_replay_import_chunk");
}
