// lgmi_tilegram.cuh -- the mid-depth path: units too large for k_pairs_fast (more than
// 64 sites or 256 reads) and too shallow for k_gram_i8's one-GEMM-per-unit form, counted
// on the tensor cores in ONE batched launch over all of them.
//
//   k_tile_gram   work item = (unit, block I of 128 sites, block J of 48 sites) with some pair
//                 i < j.  Per 128-read k-block the CTA expands the unit's bit-planes IN THE
//                 KERNEL (no indicator matrix in HBM, no TMA) straight into the K-major
//                 128-byte-swizzled shared-memory layout tcgen05 reads:
//                     A_a  128 rows (site i, label a), a = other / minor / major   3 x 16 KB
//                     B    144 rows (label b major: row = 48 b + j)                    18 KB
//                 and one thread issues, per K = 32 step, three
//                 tcgen05.mma.cta_group::1.kind::i8 M=128 N=144 -- one per label a, each into
//                 its own 144 TMEM columns.  A TMEM lane is then a SITE: the thread that owns
//                 lane i reads the nine cells of pair (i, j) from its own lane
//                 (D_a[i][48 b + j]) without a shuffle.  Two smem stages: the expansion of
//                 k-block k+1 overlaps the MMAs of k-block k (mbarrier per stage, bounded waits).
//                 Readout: 8 warps = 4 lane quarters x 2 column halves, tcgen05.ld 32x32b.x8,
//                 nine u16 counts + flag bits per pair (24 B) into a scratch in pair order.
//   k_tile_finish one thread per pair: min-common filter + fp64 MI from the nine counts ->
//                 the unit's dense MI scratch (NaN: dropped / not evaluated) and the number of
//                 emitted pairs per work item (what k_count does for the other paths).
// k_pairs_generic<true> / k_site_mean_dense then order, emit and average as for k_tile_mi,
// which stays as the popcount form of this path (lgmi_set_tile_path(ctx, 0); units deeper
// than 65 535 reads, whose counts do not fit 16 bits, always take it).
//
// Reference semantics: the counts of /root/reference/src/giremi/mutual_information.py:15-40
// (labels over the common reads, strict '<' drop at :19), bit-exact; MI as lgmi_math.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// (included at the end of lgmi_kernels.cuh: DevUnit, Item, TileItem, RunParams, LnGlobal are defined there)

namespace lgmi {

constexpr int kTgSitesI = 128;                        // sites per row block: one per TMEM lane
constexpr int kTgSitesJ = 48;                         // partner sites per column block
constexpr int kTgN = 3 * kTgSitesJ;                   // 144 columns per accumulator (label-major)
constexpr int kTgThreads = 256;
constexpr uint32_t kTgATile = 128u * 128u;            // one label tile of A: 128 rows x 128 reads, 16 KB
constexpr uint32_t kTgABytes = 3u * kTgATile;         // 48 KB
constexpr uint32_t kTgBBytes = (uint32_t)kTgN * 128u; // 18 KB
constexpr uint32_t kTgStageBytes = kTgABytes + kTgBBytes;
constexpr int kTgStages = 2;
constexpr uint32_t kTgSmemBytes = kTgStages * kTgStageBytes + 1024u /*align*/ + 256u /*barriers*/;
constexpr uint32_t kTgTmemCols = 512;                 // 3 x 144 used
constexpr uint32_t kTgMaxReads = 65535;               // counts are stored as u16

// flag bits of a pair's 24-byte count record
constexpr uint32_t kTgNotEvaluated = 1u;  // SKIP_NONHET and no het_snp partner
constexpr uint32_t kTgHetPair = 2u;       // one of the two sites is a het_snp

__device__ __forceinline__ void tmem_load_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// tcgen05.wait::ld with the loaded registers as in/out operands: nothing that reads them can be
// scheduled above the wait
__device__ __forceinline__ void tmem_wait_x8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}

// one k-block (128 reads) of the tile's operands: thread item = (site, 32-read word) -> the three
// label rows' 2 x 16 bytes each.  Rows of absent sites (>= S) are left as they are: their products
// land in accumulator rows / columns nobody reads.
__device__ __forceinline__ void tg_expand(uint8_t* __restrict__ stage, const uint32_t* __restrict__ unit_planes, uint32_t W,
                                          uint32_t S, uint32_t i0, uint32_t j0, uint32_t kb) {
  constexpr uint32_t kItemsA = (uint32_t)kTgSitesI * 4u, kItems = kItemsA + (uint32_t)kTgSitesJ * 4u;
#pragma unroll 3
  for (uint32_t e = threadIdx.x; e < kItems; e += kTgThreads) {
    const bool is_a = e < kItemsA;
    const uint32_t local = is_a ? e : e - kItemsA;
    const uint32_t sl = local >> 2, w = local & 3u;
    const uint32_t s = (is_a ? i0 : j0) + sl;
    if (s >= S) continue;
    const uint32_t* __restrict__ src = unit_planes + (size_t)s * 3u * W + kb * 4u + w;
    const uint32_t M = __ldg(src), m = __ldg(src + W), C = __ldg(src + 2u * W);
    const uint32_t L2 = M & C, L1 = m & C & ~M, L0 = C & ~M & ~m;
    uint8_t* t0 = is_a ? stage : stage + kTgABytes;                 // label "other"
    uint8_t* t1 = is_a ? stage + kTgATile : t0;                     // minor
    uint8_t* t2 = is_a ? stage + 2u * kTgATile : t0;                // major
    const uint32_t r0 = sl, r1 = is_a ? sl : (uint32_t)kTgSitesJ + sl, r2 = is_a ? sl : 2u * (uint32_t)kTgSitesJ + sl;
    sg_store(t0, r0, 2u * w, spread16(L0 & 0xffffu));
    sg_store(t0, r0, 2u * w + 1u, spread16(L0 >> 16));
    sg_store(t1, r1, 2u * w, spread16(L1 & 0xffffu));
    sg_store(t1, r1, 2u * w + 1u, spread16(L1 >> 16));
    sg_store(t2, r2, 2u * w, spread16(L2 & 0xffffu));
    sg_store(t2, r2, 2u * w + 1u, spread16(L2 >> 16));
  }
}

struct TileGramParams {
  const DevUnit* units;
  const uint32_t* planes;
  const uint8_t* site_flags;
  uint32_t mode;
  const TileItem* tiles;   // I: block of 128 sites, J: block of 48 sites
  uint32_t n_tiles;
  uint2* cnt;              // three 8-byte words per pair slot (DevUnit::dense_off + pair index)
  uint32_t* next;          // work counter (zeroed by k_run_init)
  uint32_t* error;
};

__global__ void __launch_bounds__(kTgThreads, 1) k_tile_gram(const TileGramParams P) {
  extern __shared__ uint8_t tg_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tg_raw) + 1023u) & ~uintptr_t(1023));
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem + kTgStages * kTgStageBytes);  // [stages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + kTgStages);
  uint32_t* s_next = tmem_slot + 1;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < kTgStages; ++s) mbar_init(&mma_done[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *s_next = atomicAdd(P.next, 1u);
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTgTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = umma_idesc_u8(128u, (uint32_t)kTgN);
  const bool skip_nonhet = (P.mode & LGMI_MODE_HET_ONLY) && (P.mode & LGMI_MODE_SKIP_NONHET);

  // every commit on a stage's barrier is waited for exactly once, in order (parity per stage)
  // (bit s of each word: no runtime-indexed arrays, which would live in local memory)
  uint32_t parity = 0u, pending = 0u, stage = 0u;
  auto wait_stage = [&](uint32_t s) {
    if ((pending >> s) & 1u) {
      mbar_wait(&mma_done[s], (parity >> s) & 1u, P.error);
      parity ^= 1u << s;
      pending &= ~(1u << s);
    }
  };

  uint32_t t = *s_next;
  while (t < P.n_tiles) {
    const TileItem tile = P.tiles[t];
    const DevUnit u = P.units[tile.unit];
    const uint32_t* __restrict__ unit_planes = P.planes + u.plane_off;
    const uint32_t i0 = (uint32_t)tile.I * (uint32_t)kTgSitesI, j0 = (uint32_t)tile.J * (uint32_t)kTgSitesJ;
    const uint32_t nkb = u.W >> 2;  // whole k-blocks of 128 reads (W is a multiple of 4; pad bits are zero)

    for (uint32_t kb = 0; kb < nkb; ++kb) {
      wait_stage(stage);  // the MMAs that read this stage two k-blocks ago
      uint8_t* st = smem + stage * kTgStageBytes;
      tg_expand(st, unit_planes, u.W, u.S, i0, j0, kb);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t sa = smem_u32(st), sb = sa + kTgABytes;
        const uint64_t db = umma_desc_sw128(sb);
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {  // K = 32 reads per instruction: +32 B on both operands
#pragma unroll
          for (uint32_t a = 0; a < 3; ++a)
            umma_i8(tmem_base + a * (uint32_t)kTgN, umma_desc_sw128(sa + a * kTgATile) + 2ull * k, db + 2ull * k, idesc,
                    (kb | k) != 0u);
        }
        umma_commit(&mma_done[stage]);
      }
      pending |= 1u << stage;
      stage ^= 1u;
    }
    if (tid == 0) *s_next = atomicAdd(P.next, 1u);  // (read after the barrier that ends the readout)
    // all MMAs of the tile: the commits complete in issue order
    wait_stage(stage);  // older commit first
    wait_stage(stage ^ 1u);
    tc_fence_after();

    // ---- readout: lane = site i; this warp's lane quarter and column half
    {
      const uint32_t lq = warp & 3u, half = warp >> 2;
      const uint32_t i = i0 + lq * 32u + lane;
      const uint8_t* __restrict__ flags = P.site_flags + u.site_off;
      const bool i_ok = i < u.S;
      const bool het_i = i_ok && (flags[i] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP;
      const uint64_t row = u.dense_off + lg_row_off(i_ok ? i : 0u, u.S);
      // nothing of this warp's 32 rows pairs with a column of this half: skip the loads
      const uint32_t j_hi = min(j0 + (half + 1u) * 24u, u.S);  // exclusive
      if (i0 + lq * 32u + 1u < j_hi) {
#pragma unroll 1
        for (uint32_t c = 0; c < 3u; ++c) {
          const uint32_t jl = half * 24u + c * 8u;  // first of eight partner columns
          if (j0 + jl >= u.S || i0 + lq * 32u >= j0 + jl + 7u) continue;  // (warp-uniform) no pair i < j here
          uint32_t r[3][3][8];
#pragma unroll
          for (uint32_t a = 0; a < 3; ++a)
#pragma unroll
            for (uint32_t b = 0; b < 3; ++b)
              tmem_load_x8(tmem_base + ((lq * 32u) << 16) + a * (uint32_t)kTgN + b * (uint32_t)kTgSitesJ + jl, r[a][b]);
#pragma unroll
          for (uint32_t a = 0; a < 3; ++a)
#pragma unroll
            for (uint32_t b = 0; b < 3; ++b) tmem_wait_x8(r[a][b]);
#pragma unroll
          for (uint32_t q = 0; q < 8; ++q) {
            const uint32_t j = j0 + jl + q;
            if (!i_ok || j >= u.S || i >= j) continue;
            const bool het = het_i || (flags[j] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP;
            const uint32_t fl = (het ? kTgHetPair : 0u) | ((skip_nonhet && !het) ? kTgNotEvaluated : 0u);
            uint2 w0, w1, w2;
            w0.x = r[0][0][q] | (r[0][1][q] << 16);
            w0.y = r[0][2][q] | (r[1][0][q] << 16);
            w1.x = r[1][1][q] | (r[1][2][q] << 16);
            w1.y = r[2][0][q] | (r[2][1][q] << 16);
            w2.x = r[2][2][q] | (fl << 16);
            w2.y = 0u;
            uint2* out = P.cnt + (row + (j - i - 1u)) * 3ull;
            out[0] = w0;
            out[1] = w1;
            out[2] = w2;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // accumulators read: TMEM may be overwritten; s_next is the next tile
    tc_fence_after();
    t = *s_next;
    __syncthreads();  // everybody has read s_next before thread 0 replaces it
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTgTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------
// MI of the pairs k_tile_gram counted: one CTA per work item (<= 2048 consecutive pairs of a unit),
// one thread per pair.  Writes the unit's dense MI scratch (what k_tile_mi writes) and the item's
// number of emitted pairs (what k_count computes for the other paths).
__global__ void __launch_bounds__(kThreads) k_tile_finish(const RunParams P, const uint2* __restrict__ cnt) {
  __shared__ uint32_t s_warp[kThreads / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const LnGlobal ln{P.lntab};
  const bool het_only = (P.mode & LGMI_MODE_HET_ONLY) != 0u;
  for (uint32_t item_idx = blockIdx.x; item_idx < P.n_items; item_idx += gridDim.x) {
    const Item it = P.items[item_idx];
    if (!(it.flags & ITEM_TILED_GRAM)) continue;
    const DevUnit u = P.units[it.unit];
    uint32_t mine = 0;
    for (uint32_t pl = tid; pl < it.pair_cnt; pl += kThreads) {
      const uint64_t slot = u.dense_off + it.pair_begin + pl;
      const uint2 w0 = __ldg(cnt + slot * 3ull), w1 = __ldg(cnt + slot * 3ull + 1), w2 = __ldg(cnt + slot * 3ull + 2);
      uint32_t T[9];
      T[0] = w0.x & 0xffffu; T[1] = w0.x >> 16; T[2] = w0.y & 0xffffu; T[3] = w0.y >> 16;
      T[4] = w1.x & 0xffffu; T[5] = w1.x >> 16; T[6] = w1.y & 0xffffu; T[7] = w1.y >> 16;
      T[8] = w2.x & 0xffffu;
      const uint32_t fl = w2.x >> 16;
      uint32_t N = 0;
#pragma unroll
      for (int k = 0; k < 9; ++k) N += T[k];
      double v = lg_nan();
      if (!(fl & kTgNotEvaluated) && (int)N >= P.min_common) {  // strict '<' drops (mutual_information.py:19)
        if ((T[0] | T[1] | T[2] | T[3] | T[6]) == 0u) v = lg_mi_from_2x2(T[4], T[5], T[7], T[8], ln);
        else v = lg_mi_from_table(T, ln);
        mine += ((fl & kTgHetPair) || !het_only) ? 1u : 0u;
      }
      P.dense[slot] = v;
      if (P.mode & LGMI_MODE_EMIT_COUNTS) {
#pragma unroll
        for (int k = 0; k < 9; ++k) P.tile_counts[slot * 9ull + k] = T[k];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    __syncthreads();  // s_warp of the previous item has been read
    if (lane == 0) s_warp[warp] = mine;
    __syncthreads();
    if (tid == 0) {
      uint32_t tot = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) tot += s_warp[w];
      P.item_cnt[item_idx] = tot;
    }
  }
}

}  // namespace lgmi
