// lgmi_small.cuh -- small units (S <= 60 sites, R <= 256 reads) with their
// contingency counts on the tensor cores.
//
// The popcount form of the small-unit path spends about half of its time on
// AND / carry-save / POPC instructions (k_count + the counts phase of
// k_pairs_fast).  Here the counts of a whole unit are ONE small Gram matrix:
//
//   k_small_gram   per unit (persistent CTAs, two per SM):
//       land     the unit's planes by cp.async (next unit prefetched)
//       expand   bits -> 0/1 bytes straight into shared memory in the K-major
//                128-byte-swizzled layout tcgen05 reads (no X in HBM, no TMA);
//                row of (site s, label a) = 32*(s/10) + 3*(s%10) + a, so that the
//                three label rows of a site sit in three adjacent TMEM lanes
//       mma      G = X * X^T: tcgen05.mma.cta_group::1.kind::i8, M = 128,
//                N = 32*ceil(S/10) (<= 192), K = 32 x (4 or 8), s32 accumulators in
//                TMEM; sites 40.. are a second M block against rows 128..
//       readout  tcgen05.ld 32 lanes x 32 columns = 10 x 10 site pairs per warp; the
//                three lanes of a site swap their label rows with three shuffles
//                per partner, one of them packs the pair's nine cells into the
//                64-bit form k_pairs_fast's epilogue already consumes, applies the
//                min-common filter and stores it in pair order
//     -> val[unit][pair] (8 bytes per candidate pair) + emitted pairs per unit
//   (scan of the counts, as before)
//   k_pairs_pre    the MI / ordered emit / per-site mean phases of k_pairs_fast over
//                  the precounted values (cp.async double-buffered), no popcounts.
//
// Reference semantics: /root/reference/src/giremi/mutual_information.py:6-60; the
// counts are the same integers as the popcount path's (tests compare the two).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lgmi.h"
#include "lgmi_dense.cuh"
#include "lgmi_fast_kernel.cuh"

namespace lgmi {

constexpr int kSgThreads = 256;
constexpr int kSgSitesPerGroup = 10;                 // 30 of the 32 rows (TMEM lanes) of a group
constexpr int kSgMaxGroups = 6;
constexpr int kSgMaxS = kSgSitesPerGroup * kSgMaxGroups;  // 60
constexpr uint32_t kSgRows = 256;                    // rows of X kept addressable: two M blocks of 128
constexpr uint32_t kSgKbBytes = kSgRows * 128u;      // one k-block (128 reads) of X
constexpr uint32_t kSgTmemCols = 256;                // D0 at column 0 (<= 192), D1 behind it (<= 64)
constexpr unsigned long long kValEmit = 1ull << 52;  // the pair is written as a record (min-common and het filter passed)

struct SgParams {
  const FastItem* items;  // FastItem::pad = first slot of the unit in `val`
  uint32_t n_items;
  const uint32_t* planes;
  const uint8_t* site_flags;
  int min_common;
  uint32_t mode;
  unsigned long long* val;
  unsigned long long* item_cnt;
  uint8_t* item_dense;
  uint32_t* n_generic;
  uint32_t* error;
};

struct SgSmem {
  uint8_t X[2][kSgKbBytes];                 // 64 KB, 1024-byte aligned
  uint32_t planes[2][kSgMaxS * 24];         // landed rows [M 8 | m 8 | C 8], double-buffered
  uint8_t flags[2][64];
  uint32_t het[2];                          // het_snp sites of the current unit, bit s
  uint64_t mma_bar;
  uint32_t tmem_slot;
  uint32_t n_emit, overflow;
};

// request one unit's plane rows: 6 x 16 B per site
__device__ __forceinline__ void sg_prefetch(uint32_t* __restrict__ dst, const FastItem& it,
                                            const uint32_t* __restrict__ planes) {
  const uint32_t W = it.W, W4 = W >> 2;
  const uint32_t* __restrict__ src = planes + it.plane_off;
  for (uint32_t e = threadIdx.x; e < (uint32_t)it.S * 6u; e += kSgThreads) {
    const uint32_t s = e / 6u, q = e - s * 6u;
    const uint32_t plane = q >> 1, half = q & 1u;
    const bool have = half < W4;
    cp_async16(dst + s * 24u + plane * 8u + half * 4u, src + (size_t)s * 3u * W + plane * W + (have ? half * 4u : 0u), have);
  }
}

// 16 bits -> 16 bytes of 0/1
__device__ __forceinline__ uint4 spread16(uint32_t bits) {
  uint4 o;
  o.x = spread4(bits & 15u);
  o.y = spread4((bits >> 4) & 15u);
  o.z = spread4((bits >> 8) & 15u);
  o.w = spread4((bits >> 12) & 15u);
  return o;
}

__device__ __forceinline__ uint32_t sg_row(uint32_t s, uint32_t a) {
  return 32u * (s / kSgSitesPerGroup) + 3u * (s % kSgSitesPerGroup) + a;
}

__device__ __forceinline__ void sg_store(uint8_t* __restrict__ Xkb, uint32_t row, uint32_t chunk, uint4 v) {
  // K-major SWIZZLE_128B: 8-row groups of 1024 B, 16-byte chunk index XOR row-in-group
  *reinterpret_cast<uint4*>(Xkb + (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4)) = v;
}

__global__ void __launch_bounds__(kSgThreads, 2) k_small_gram(const SgParams P) {
  extern __shared__ uint8_t sg_raw[];
  SgSmem& sm = *reinterpret_cast<SgSmem*>((reinterpret_cast<uintptr_t>(sg_raw) + 1023u) & ~uintptr_t(1023));
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const bool het_only = (P.mode & LGMI_MODE_HET_ONLY) != 0u;
  const bool skip_nonhet = het_only && (P.mode & LGMI_MODE_SKIP_NONHET) != 0u;

  if (tid == 0) {
    mbar_init(&sm.mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_slot)),
                 "r"(kSgTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_slot;

  uint32_t idx = blockIdx.x, buf = 0, phase = 0;
  FastItem it;
  if (idx < P.n_items) {
    it = P.items[idx];
    sg_prefetch(sm.planes[0], it, P.planes);
    if (tid < it.S) sm.flags[0][tid] = P.site_flags[it.site_off + tid];
  }
  cp_async_commit();

  while (idx < P.n_items) {
    const uint32_t idx_next = idx + gridDim.x;
    const uint32_t S = it.S;
    const uint32_t ng = (S + kSgSitesPerGroup - 1u) / kSgSitesPerGroup;  // row groups of 32
    const uint32_t nkb = it.W >> 2;                                        // k-blocks of 128 reads: 1 or 2
    if (tid == 0) {
      sm.n_emit = 0u;
      sm.overflow = 0u;
    }
    cp_async_wait<0>();
    __syncthreads();  // planes + flags of this unit landed; the previous unit's readout is over (X, TMEM free)

    if (warp < 2u) {
      const uint32_t s = warp * 32u + lane;
      const bool h = s < S && (sm.flags[buf][s] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP;
      const uint32_t m = __ballot_sync(0xffffffffu, h);
      if (lane == 0) sm.het[warp] = m;
    }
    // ---- expand: one thread per (site, 16-read chunk), the three label rows at once
    {
      const uint32_t* __restrict__ pl = sm.planes[buf];
      const uint32_t n_chunks = nkb * 8u;
      for (uint32_t e = tid; e < S * n_chunks; e += kSgThreads) {
        const uint32_t s = e / n_chunks, q = e - s * n_chunks;
        const uint32_t sh = (q & 1u) * 16u;
        const uint32_t* row = pl + s * 24u + (q >> 1);
        const uint32_t M = (row[0] >> sh) & 0xffffu, m = (row[8] >> sh) & 0xffffu, C = (row[16] >> sh) & 0xffffu;
        const uint32_t L2 = M & C, L1 = m & C & ~M, L0 = C & ~M & ~m;
        uint8_t* Xkb = sm.X[q >> 3];
        const uint32_t chunk = q & 7u, r0 = sg_row(s, 0u);
        sg_store(Xkb, r0, chunk, spread16(L0));
        sg_store(Xkb, r0 + 1u, chunk, spread16(L1));
        sg_store(Xkb, r0 + 2u, chunk, spread16(L2));
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
    __syncthreads();

    // ---- mma (one thread) while the others request the next unit
    FastItem it_next = it;
    if (idx_next < P.n_items) it_next = P.items[idx_next];
    if (tid == 0) {
      tc_fence_after();
      const uint32_t n0 = 32u * ng;
      const uint32_t idesc0 = umma_idesc_u8(128u, n0);
      for (uint32_t kb = 0; kb < nkb; ++kb) {
        const uint64_t d = umma_desc_sw128(smem_u32(sm.X[kb]));
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) umma_i8(tmem_base, d + 2ull * k, d + 2ull * k, idesc0, (kb | k) != 0u);
      }
      if (ng > 4u) {  // sites 40..: rows 128.. against rows 128.. (only pairs i < j are needed)
        const uint32_t idesc1 = umma_idesc_u8(128u, n0 - 128u);
        for (uint32_t kb = 0; kb < nkb; ++kb) {
          const uint64_t d = umma_desc_sw128(smem_u32(sm.X[kb]) + 128u * 128u);
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k) umma_i8(tmem_base + n0, d + 2ull * k, d + 2ull * k, idesc1, (kb | k) != 0u);
        }
      }
      umma_commit(&sm.mma_bar);
    }
    if (idx_next < P.n_items) {
      sg_prefetch(sm.planes[buf ^ 1u], it_next, P.planes);
      if (tid < it_next.S) sm.flags[buf ^ 1u][tid] = P.site_flags[it_next.site_off + tid];
    }
    cp_async_commit();

    mbar_wait(&sm.mma_bar, phase, P.error);
    phase ^= 1u;
    tc_fence_after();

    // ---- readout: warp -> TMEM lane quarter lq = warp & 3; the two warps of a quarter alternate column groups
    {
      const uint32_t lq = warp & 3u, half = warp >> 2;
      const uint32_t si = lane / 3u, a = lane - 3u * si;  // lanes 30, 31 carry the padding rows
      const unsigned long long het_mask = (unsigned long long)sm.het[0] | ((unsigned long long)sm.het[1] << 32);
      unsigned long long* __restrict__ vout = P.val + it.pad;
      uint32_t n_emit = 0u;
      bool overflow = false;
      for (uint32_t mb = 0; mb < 2u; ++mb) {
        const uint32_t g = 4u * mb + lq;
        if (g >= ng) continue;
        const uint32_t i = kSgSitesPerGroup * g + si;
        const bool row_ok = lane < 30u && i < S;
        const bool het_i = ((het_mask >> (i & 63u)) & 1ull) != 0ull;
        unsigned long long* __restrict__ vrow = vout + ((long long)((i * (2u * S - i - 1u)) / 2u) - (long long)i - 1ll);  // + j
        for (uint32_t q = g + half; q < ng; q += 2u) {
          uint32_t r[32];
          const uint32_t col = mb ? (32u * ng + 32u * (q - 4u)) : 32u * q;
          tmem_load_32x32(tmem_base + ((lq * 32u) << 16) + col, r);
          uint32_t k0[4], k1[4], k2[4];  // the three label rows of site i for "my" partners sj = a, a+3, a+6, a+9
#pragma unroll
          for (int sj = 0; sj < kSgSitesPerGroup; ++sj) {
            const uint32_t w = r[3 * sj] | (r[3 * sj + 1] << 10) | (r[3 * sj + 2] << 20);
            const uint32_t w0 = __shfl_sync(0xffffffffu, w, 3u * si);
            const uint32_t w1 = __shfl_sync(0xffffffffu, w, 3u * si + 1u);
            const uint32_t w2 = __shfl_sync(0xffffffffu, w, 3u * si + 2u);
            if ((uint32_t)(sj % 3) == a) {
              k0[sj / 3] = w0;
              k1[sj / 3] = w1;
              k2[sj / 3] = w2;
            }
          }
          const uint32_t j0 = kSgSitesPerGroup * q + a;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t j = j0 + 3u * t;
            const bool ok = row_ok && 3u * t + a < (uint32_t)kSgSitesPerGroup && j < S && i < j;
            bool emit = false;
            if (ok) {
              unsigned long long v = kNoMi;
              const uint32_t T11 = (k1[t] >> 10) & 1023u, T12 = k1[t] >> 20;
              const uint32_t T21 = (k2[t] >> 10) & 1023u, T22 = k2[t] >> 20;
              const uint32_t nMP = T21 + T22, nPM = T12 + T22, nPP = T11 + T12 + nMP;
              const uint32_t oth = k0[t] | ((k1[t] | k2[t]) & 1023u);  // any cell with an "other" label
              uint32_t n_common = nPP, o = 0u;
              if (oth) {
                const uint32_t T00 = k0[t] & 1023u, T01 = (k0[t] >> 10) & 1023u, T02 = k0[t] >> 20;
                const uint32_t T10 = k1[t] & 1023u, T20 = k2[t] & 1023u;
                n_common += T00 + T01 + T02 + T10 + T20;
                overflow |= ((T00 | T01 | T02 | T10 | T20) > 7u);
                o = (T00 & 7u) | ((T01 & 7u) << 3) | ((T02 & 7u) << 6) | ((T10 & 7u) << 9) | ((T20 & 7u) << 12);
              }
              const bool het = het_i || ((het_mask >> j) & 1ull) != 0ull;
              if ((het || !skip_nonhet) && (int)n_common >= P.min_common) {  // strict '<' drops (mutual_information.py:19)
                emit = het || !het_only;
                v = (unsigned long long)(nPP | (nMP << 9) | (nPM << 18)) | ((unsigned long long)T22 << 27) |
                    ((unsigned long long)(o | (emit ? 1u << 16 : 0u)) << 36);
              }
              vrow[j] = v;
            }
            n_emit += __popc(__ballot_sync(0xffffffffu, emit));
          }
        }
      }
      if (lane == 0 && n_emit) atomicAdd(&sm.n_emit, n_emit);
      if (__any_sync(0xffffffffu, overflow) && lane == 0) sm.overflow = 1u;
    }
    tc_fence_before();
    __syncthreads();  // readout finished: counters complete, TMEM and X reusable
    if (tid == 0) {
      P.item_cnt[it.item] = sm.n_emit;
      if (sm.overflow) {  // a cell with more than 7 "other" reads: the unit goes to the generic kernel
        P.item_dense[it.item] = 1;
        atomicAdd(P.n_generic, 1u);
      }
    }
    it = it_next;
    idx = idx_next;
    buf ^= 1u;
  }
  cp_async_wait<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kSgTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------
// k_pairs_pre: the MI / emit / mean phases of k_pairs_fast over precounted values
struct PreSmem {
  FastTab tab;
  unsigned long long val[2][kFastMaxPairs];  // landing buffers of the packed counts; MI bits in place
  uint16_t list[kFastMaxPairs];
  uint32_t emit_mask[kFastChunks];
  uint32_t chunk_off[kFastChunks];
  uint8_t flags[kFastMaxS];
  uint8_t info[kFastMaxS];
  uint8_t het_list[kFastMaxS];
  uint8_t nonhet_list[kFastMaxS];
  unsigned long long het_mask;
  uint32_t n_list2, n_list3, next_chunk, total;
  __device__ __forceinline__ uint32_t oth_count(uint32_t) const { return 0u; }
};

struct PreParams {
  FastParams F;
  const unsigned long long* val;
};

__device__ __forceinline__ void pre_prefetch(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src,
                                             uint32_t n_pairs) {
  // the unit's first slot is a multiple of 2 (16-byte aligned): whole 16-byte pieces, the odd tail zero-filled
  const uint32_t n16 = (n_pairs + 1u) >> 1;
  for (uint32_t e = threadIdx.x; e < n16; e += kFastThreads) cp_async16(dst + 2u * e, src + 2u * e, true);
}

__global__ void __launch_bounds__(kFastThreads, 4) k_pairs_pre(const PreParams Q) {
  extern __shared__ __align__(16) unsigned char pre_raw[];
  PreSmem& sm = *reinterpret_cast<PreSmem*>(pre_raw);
  const FastParams& P = Q.F;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;

  for (uint32_t k = tid; k <= (uint32_t)kFastMaxR; k += kFastThreads) {
    const double2 v = (k < P.ln_cap) ? __ldg(reinterpret_cast<const double2*>(P.lntab) + k) : make_double2(0.0, 0.0);
    sm.tab.ln[k].hi = v.x;
    sm.tab.ln[k].lo = v.y;
    sm.tab.inv[k] = k ? __drcp_rn((double)k) : 0.0;
  }

  uint32_t idx = blockIdx.x, buf = 0;
  FastItem it;
  if (idx < P.n_items) {
    it = P.items[idx];
    pre_prefetch(sm.val[0], Q.val + it.pad, (uint32_t)it.S * (it.S - 1u) / 2u);
  }
  cp_async_commit();

  while (idx < P.n_items) {
    const uint32_t idx_next = idx + gridDim.x;
    FastItem it_next = it;
    if (idx_next < P.n_items) {
      it_next = P.items[idx_next];
      pre_prefetch(sm.val[buf ^ 1u], Q.val + it_next.pad, (uint32_t)it_next.S * (it_next.S - 1u) / 2u);
    }
    cp_async_commit();

    const uint32_t S = it.S;
    const uint32_t n_pairs = S * (S - 1u) / 2u;
    const uint32_t n_chunks = (n_pairs + 31u) >> 5;
    const bool dense = P.item_dense[it.item] != 0u;  // handled by the generic kernel
    const unsigned long long base = P.item_off[it.item];
    if (!dense) {
      unsigned long long* val = sm.val[buf];
      const uint16_t* __restrict__ ijt = P.ij_tab + lg_ij_tab_off(S);
      if (tid == 0) {
        sm.n_list2 = 0u;
        sm.n_list3 = 0u;
      }
      if (warp < 2u) {  // site flags + het mask (sites ascending)
        const uint32_t s = warp * 32u + lane;
        uint32_t f = 0u;
        if (s < S) {
          f = P.site_flags[it.site_off + s];
          sm.flags[s] = (uint8_t)f;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, s < S && (f & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP);
        if (lane == 0) reinterpret_cast<uint32_t*>(&sm.het_mask)[warp] = m;
      }
      cp_async_wait<1>();
      __syncthreads();
      fast_site_lists(sm, S);
      // emit masks and the two lists straight from the precounted values
      const uint32_t n_slots = (n_pairs + 31u) & ~31u;
      for (uint32_t p = tid; p < n_slots; p += kFastThreads) {
        uint32_t cls = 0u;
        bool emit = false;
        if (p < n_pairs) {
          const unsigned long long v = val[p];
          if (v != kNoMi) {
            cls = ((v >> 36) & 0x7fffull) ? 3u : 2u;
            emit = (v & kValEmit) != 0ull;
          }
        }
        const uint32_t me = __ballot_sync(0xffffffffu, emit);
        const uint32_t m2 = __ballot_sync(0xffffffffu, cls == 2u);
        const uint32_t m3 = __ballot_sync(0xffffffffu, cls == 3u);
        uint32_t b2 = 0u, b3 = 0u;
        if (lane == 0) {
          sm.emit_mask[p >> 5] = me;
          if (m2) b2 = atomicAdd(&sm.n_list2, (uint32_t)__popc(m2));
          if (m3) b3 = atomicAdd(&sm.n_list3, (uint32_t)__popc(m3));
        }
        b2 = __shfl_sync(0xffffffffu, b2, 0);
        b3 = __shfl_sync(0xffffffffu, b3, 0);
        if (cls == 2u) sm.list[b2 + __popc(m2 & lt)] = (uint16_t)p;
        if (cls == 3u) sm.list[kFastMaxPairs - 1u - (b3 + __popc(m3 & lt))] = (uint16_t)p;
      }
      __syncthreads();
      if (tid < 32u) fast_chunk_prefix(sm, n_chunks);
      if (P.mode & LGMI_MODE_EMIT_COUNTS) {
        __syncthreads();
        fast_emit_counts(P, sm, val, base, n_chunks);
      }
      fast_mi(sm, val);
      __syncthreads();
      if (tid == 0) P.unit_rec_off[it.unit] = base;
      const uint32_t mean_warps = (S + 31u) >> 5;
      if (warp < mean_warps) fast_means(P, sm, val, it);
      else fast_emit(P, sm, val, it, ijt, base, n_chunks, mean_warps);
    } else {
      cp_async_wait<1>();
    }
    __syncthreads();
    it = it_next;
    idx = idx_next;
    buf ^= 1u;
  }
  cp_async_wait<0>();
}

}  // namespace lgmi
