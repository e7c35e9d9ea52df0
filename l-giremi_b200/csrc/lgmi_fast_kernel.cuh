// lgmi_fast_kernel.cuh -- k_pairs_fast: the persistent small-unit pair kernel.
// See lgmi_fast.cuh for the phase description and the arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lgmi.h"
#include "lgmi_fast.cuh"

namespace lgmi {

struct FastParams {
  const FastItem* items;
  uint32_t n_items;
  const uint32_t* planes;
  const uint8_t* site_flags;
  const lg_dd* lntab;
  uint32_t ln_cap;
  const uint16_t* ij_tab;              // triangular pair tables, see lg_ij_tab_off
  int min_common;
  uint32_t mode;
  const unsigned long long* item_off;  // exclusive scan of the per-item emit counts
  const uint8_t* fast_empty;           // per entry of `items`: nothing to emit (written by k_count_fast)
  uint8_t* item_dense;                 // 1: some site has more than kOthCap "other" reads -> generic kernel
  uint32_t* n_generic;                 // items the generic kernel has to take (incremented with item_dense)
  lgmi_pair_rec* records;
  uint32_t* counts;                    // EMIT_COUNTS: 9 per record
  double* site_mean;
  uint32_t* site_cnt;
  unsigned long long* unit_rec_off;
  uint32_t unit_base;                  // added to the unit field of every record
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool copy) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = copy ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// issue the loads of one unit's plane rows into a landing buffer: 6 x 16 B per site
// (t, nt: index and size of the thread group that shares the work -- the whole CTA by default)
__device__ __forceinline__ void fast_prefetch(uint32_t* __restrict__ rows, const FastItem& it,
                                              const uint32_t* __restrict__ planes, uint32_t t = threadIdx.x,
                                              uint32_t nt = kFastThreads) {
  const uint32_t W = it.W, W4 = W >> 2;
  const uint32_t* __restrict__ src = planes + it.plane_off;
  for (uint32_t e = t; e < (uint32_t)it.S * 6u; e += nt) {
    const uint32_t s = e / 6u, q = e - s * 6u;
    const uint32_t plane = q >> 1, half = q & 1u;
    const bool have = half < W4;
    cp_async16(rows + s * kRowStride + plane * 8u + half * 4u,
               src + (size_t)s * 3u * W + plane * W + (have ? half * 4u : 0u), have);
  }
}

// in-place transform of the landed rows + "other" lists + site flags
__device__ __forceinline__ void fast_land(FastSmem& sm, uint32_t* __restrict__ rows, const FastItem& it,
                                          const uint8_t* __restrict__ flags, uint32_t tid = threadIdx.x,
                                          uint32_t nt = kFastThreads) {
  const uint32_t lane = tid & 31u, warp = tid >> 5;
  const uint32_t S = it.S;
  for (uint32_t e = tid; e < S * 8u; e += nt) {
    const uint32_t s = e >> 3, k = e & 7u;
    uint32_t* row = rows + s * kRowStride;
    const uint32_t C = row[16 + k];
    const uint32_t M = row[k] & C;
    const uint32_t Pw = (M | row[8 + k]) & C;
    row[k] = M;
    row[8 + k] = Pw;
    const uint32_t O = C & ~Pw;
    if (O) {  // a word with reads labelled "other": listed as (site, word); at most kOthCap such words per site count
      const uint32_t idx = atomicAdd(&sm.n_oth[s], (uint32_t)__popc(O));
      if (idx < (uint32_t)kOthCap) sm.oth_flat[atomicAdd(&sm.n_oth_total, 1u)] = (uint16_t)((s << 8) | k);
    }
  }
  // het mask (sites ascending) by the first two warps
  if (warp < 2u) {
    const uint32_t s = warp * 32u + lane;
    uint32_t f = 0u;
    if (s < S) {
      f = flags[s];
      sm.flags[s] = (uint8_t)f;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, s < S && (f & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP);
    if (lane == 0) reinterpret_cast<uint32_t*>(&sm.het_mask)[warp] = m;
  }
}

// ordered lists of the het sites and of the other sites (for the mean phase), and the
// per-site byte the counts phase reads: bit 0 het_snp, bits 1-4 number of "other" reads
template <class SM>
__device__ __forceinline__ void fast_site_lists(SM& sm, uint32_t S, uint32_t s = threadIdx.x) {
  if (s < S) {
    sm.info[s] = (uint8_t)(((sm.flags[s] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP ? 1u : 0u) | (sm.oth_count(s) << 1));
    const unsigned long long het = sm.het_mask;
    const uint32_t rank = __popcll(het & ((1ull << s) - 1ull));
    if ((het >> s) & 1ull) sm.het_list[rank] = (uint8_t)s;
    else sm.nonhet_list[s - rank] = (uint8_t)s;
  }
}

// The table cells that involve an "other" label, scattered from the words that hold such reads: for every
// listed (site s, word k) and every partner site t, the reads of the word that are "other" at s and covered at t
// are counted by t's label -- T[0][0], T[0][1], T[0][2] when s is the pair's first site, T[1][0], T[2][0] when it
// is the second (a read "other" at both sites is counted from the first site only) -- and the packed 3-bit
// increments go into the pair's field with ONE shared-memory atomic.  One warp per entry, lanes over the
// partners.  A site holds at most kOthCap such reads (else the unit is k_pairs_generic's), so no field overflows.
// `need` = the sites some partner of which is needed whatever the entry's site is: the het sites when only pairs
// next to a het SNP are evaluated (SKIP_NONHET).
template <bool kEveryPair>  // kEveryPair: `need` is all ones (the all-pairs kernel): no test
__device__ __forceinline__ void fast_other_cells(FastSmem& sm, const uint32_t* __restrict__ rows, uint32_t S,
                                                 unsigned long long need, uint32_t warp = threadIdx.x >> 5,
                                                 uint32_t n_warps = kFastWarps) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t E = sm.n_oth_total;
  for (uint32_t e = warp; e < E; e += n_warps) {
    const uint32_t v = sm.oth_flat[e];
    const uint32_t s = v >> 8, k = v & 255u;
    const uint32_t* rs = rows + s * kRowStride;
    const uint32_t Os = rs[16u + k] & ~rs[8u + k];  // reads of the word labelled "other" at s
    const bool s_needed = kEveryPair || ((need >> s) & 1ull) != 0ull;
    for (uint32_t t = lane; t < S; t += 32u) {
      if (t == s) continue;
      if (!kEveryPair && !s_needed && !((need >> t) & 1ull)) continue;  // this pair is not evaluated
      const uint32_t* rt = rows + t * kRowStride;
      const uint32_t x = Os & rt[16u + k];  // ... and covered at t
      if (!x) continue;
      const uint32_t Pt = rt[8u + k], Mt = rt[k];
      const uint32_t n_minor = (uint32_t)__popc(x & Pt & ~Mt), n_major = (uint32_t)__popc(x & Mt);
      uint32_t inc, p;
      if (s < t) {
        inc = (uint32_t)__popc(x & ~Pt) | (n_minor << 3) | (n_major << 6);
        p = ((s * (2u * S - s - 1u)) >> 1) + (t - s - 1u);
      } else {
        inc = (n_minor << 9) | (n_major << 12);
        if (!inc) continue;
        p = ((t * (2u * S - t - 1u)) >> 1) + (s - t - 1u);
      }
      atomicAdd(&sm.ocell[p >> 1], inc << (16u * (p & 1u)));
    }
  }
}

// the four AND+popcount sets of every pair, "other" cells, min-common filter, emit masks, lists
template <int NW>
__device__ __forceinline__ void fast_counts(const FastParams& P, FastSmem& sm, const uint32_t* __restrict__ rows,
                                            const uint16_t* __restrict__ ijt, uint32_t n_pairs) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const bool het_only = (P.mode & LGMI_MODE_HET_ONLY) != 0u;
  const bool skip_nonhet = het_only && (P.mode & LGMI_MODE_SKIP_NONHET) != 0u;
  const uint32_t n_slots = (n_pairs + 31u) & ~31u;
  for (uint32_t p = tid; p < n_slots; p += kFastThreads) {
    uint32_t cls = 0u;  // 0 no MI, 2 -> 2x2 list, 3 -> 3x3 list
    bool emit = false;
    if (p < n_pairs) {
      const uint32_t ij = __ldg(ijt + p);
      const uint32_t i = ij >> 6, j = ij & 63u;
      const uint32_t fi = sm.info[i], fj = sm.info[j];
      const bool het = ((fi | fj) & 1u) != 0u;
      unsigned long long v = kNoMi;
      if (het || !skip_nonhet) {
        const uint32_t o = (sm.ocell[p >> 1] >> (16u * (p & 1u))) & 0x7fffu;
        uint32_t n_other = 0u;
        if (o) n_other = (o & 7u) + ((o >> 3) & 7u) + ((o >> 6) & 7u) + ((o >> 9) & 7u) + (o >> 12);
        unsigned long long cnt;
        if (pair_counts<NW>(rows + i * kRowStride, rows + j * kRowStride, n_other, P.min_common, cnt)) {
          v = cnt | ((unsigned long long)o << 36);
          cls = o ? 3u : 2u;
          emit = het || !het_only;
        }
      }
      sm.val[p] = v;
    }
    // this warp's 32 pairs are one lexicographic chunk: three ballots are all the bookkeeping here; the lists
    // are built from the masks afterwards (fast_build_lists), without atomics
    const uint32_t me = __ballot_sync(0xffffffffu, emit);
    const uint32_t m2 = __ballot_sync(0xffffffffu, cls == 2u);
    const uint32_t m3 = __ballot_sync(0xffffffffu, cls == 3u);
    if (lane == 0) {
      sm.emit_mask[p >> 5] = me;
      sm.m2_mask[p >> 5] = m2;
      sm.m3_mask[p >> 5] = m3;
    }
  }
}

// The two lists of fast_mi from the chunks' class masks: every warp scans the (at most 64) chunk counts itself --
// 2x2 and 3x3 counts packed in one word -- and lists the pairs of its own chunks; no atomics, no barrier between
// scan and fill.  2x2 pairs from the front of sm.list, 3x3 pairs from the back; the lists are in pair order.
__device__ __forceinline__ void fast_build_lists(FastSmem& sm, uint32_t n_chunks) {
  static_assert(kFastChunks == 64, "two chunks per lane");
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t lo = 0u, hi = 0u;  // chunks `lane` and `lane + 32`: 2x2 count | 3x3 count << 16
  if (lane < n_chunks) lo = (uint32_t)__popc(sm.m2_mask[lane]) | ((uint32_t)__popc(sm.m3_mask[lane]) << 16);
  if (lane + 32u < n_chunks) hi = (uint32_t)__popc(sm.m2_mask[lane + 32u]) | ((uint32_t)__popc(sm.m3_mask[lane + 32u]) << 16);
  uint32_t ilo = lo, ihi = hi;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t a = __shfl_up_sync(0xffffffffu, ilo, o), b = __shfl_up_sync(0xffffffffu, ihi, o);
    if ((int)lane >= o) {
      ilo += a;
      ihi += b;
    }
  }
  const uint32_t tot_lo = __shfl_sync(0xffffffffu, ilo, 31);
  ihi += tot_lo;
  const uint32_t elo = ilo - lo, ehi = ihi - hi;  // exclusive prefixes
  if (threadIdx.x == 31u) {  // (lane 31 of warp 0 holds the totals)
    sm.n_list2 = ihi & 0xffffu;
    sm.n_list3 = ihi >> 16;
  }
  for (uint32_t c = warp; c < n_chunks; c += kFastWarps) {
    const uint32_t off = c < 32u ? __shfl_sync(0xffffffffu, elo, c) : __shfl_sync(0xffffffffu, ehi, c - 32u);
    const uint32_t m2 = sm.m2_mask[c], m3 = sm.m3_mask[c];
    const uint32_t p = c * 32u + lane;
    if ((m2 >> lane) & 1u) sm.list[(off & 0xffffu) + __popc(m2 & lt)] = (uint16_t)p;
    if ((m3 >> lane) & 1u) sm.list[kFastMaxPairs - 1u - ((off >> 16) + __popc(m3 & lt))] = (uint16_t)p;
  }
}

// SKIP_NONHET: the same for the pairs next to a het SNP only, enumerated from the het sites instead
// of filtered out of all pairs (a tenth of the sites are het: a filter would leave most lanes idle).
// One warp per (het site h, block of 32 partners t); a pair of two het sites belongs to the smaller
// one.  sm.val / sm.emit_mask were pre-filled with "no MI" / 0 for every pair of the unit.
template <int NW>
__device__ __forceinline__ void fast_counts_het(const FastParams& P, FastSmem& sm, const uint32_t* __restrict__ rows,
                                                uint32_t S) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned long long het = sm.het_mask;
  const uint32_t n_het = __popcll(het);
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t two = S > 32u ? 1u : 0u;  // one or two 32-partner blocks per het site (S <= 64)
  for (uint32_t item = warp; item < (n_het << two); item += kFastWarps) {
    const uint32_t h = sm.het_list[item >> two];
    {
      const uint32_t t = ((item & two) << 5) + lane;
      uint32_t cls = 0u, p = 0u;
      if (t < S && t != h && !(((het >> t) & 1ull) && t < h)) {
        const uint32_t i = t < h ? t : h, j = t < h ? h : t;
        p = ((i * (2u * S - i - 1u)) >> 1) + (j - i - 1u);
        const uint32_t o = (sm.ocell[p >> 1] >> (16u * (p & 1u))) & 0x7fffu;
        uint32_t n_other = 0u;
        if (o) n_other = (o & 7u) + ((o >> 3) & 7u) + ((o >> 6) & 7u) + ((o >> 9) & 7u) + (o >> 12);
        unsigned long long cnt;
        if (pair_counts<NW>(rows + i * kRowStride, rows + j * kRowStride, n_other, P.min_common, cnt)) {
          sm.val[p] = cnt | ((unsigned long long)o << 36);
          cls = o ? 3u : 2u;
          atomicOr(&sm.emit_mask[p >> 5], 1u << (p & 31u));
        }
      }
      const uint32_t m2 = __ballot_sync(0xffffffffu, cls == 2u);
      const uint32_t m3 = __ballot_sync(0xffffffffu, cls == 3u);
      uint32_t b2 = 0u, b3 = 0u;
      if (lane == 0) {
        if (m2) b2 = atomicAdd(&sm.n_list2, (uint32_t)__popc(m2));
        if (m3) b3 = atomicAdd(&sm.n_list3, (uint32_t)__popc(m3));
      }
      b2 = __shfl_sync(0xffffffffu, b2, 0);
      b3 = __shfl_sync(0xffffffffu, b3, 0);
      if (cls == 2u) sm.list[b2 + __popc(m2 & lt)] = (uint16_t)p;
      if (cls == 3u) sm.list[kFastMaxPairs - 1u - (b3 + __popc(m3 & lt))] = (uint16_t)p;
    }
  }
}

__device__ __forceinline__ void unpack_counts(unsigned long long v, uint32_t& n_mm, uint32_t& n_mM, uint32_t& n_Mm,
                                              uint32_t& n_MM) {
  const uint32_t lo = (uint32_t)v;
  const uint32_t nPP = lo & 511u, nMP = (lo >> 9) & 511u, nPM = (lo >> 18) & 511u;
  n_MM = (uint32_t)(v >> 27) & 511u;
  n_Mm = nMP - n_MM;  // i major, j minor
  n_mM = nPM - n_MM;  // i minor, j major
  n_mm = nPP - nMP - nPM + n_MM;
}

// MI of the listed pairs, in place of their packed counts; warp-sized chunks of both lists
template <class SM>
__device__ __forceinline__ void fast_mi(SM& sm, unsigned long long* __restrict__ val) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t n2 = sm.n_list2, n3 = sm.n_list3;
  const uint32_t nc2 = (n2 + 31u) >> 5, nc3 = (n3 + 31u) >> 5;
  for (uint32_t c = warp; c < nc2 + nc3; c += kFastWarps) {
    if (c < nc2) {
      const uint32_t q = c * 32u + lane;
      if (q < n2) {
        const uint32_t p = sm.list[q];
        uint32_t n_mm, n_mM, n_Mm, n_MM;
        unpack_counts(val[p], n_mm, n_mM, n_Mm, n_MM);
        const double mi = mi_2x2(sm.tab, n_mm, n_mM, n_Mm, n_MM);
        val[p] = (unsigned long long)__double_as_longlong(mi);
      }
    } else {
      const uint32_t q = (c - nc2) * 32u + lane;
      if (q < n3) {
        const uint32_t p = sm.list[kFastMaxPairs - 1u - q];
        const unsigned long long v = val[p];
        const uint32_t o = (uint32_t)(v >> 36);
        uint32_t T[9];
        unpack_counts(v, T[4], T[5], T[7], T[8]);
        T[0] = o & 7u;
        T[1] = (o >> 3) & 7u;
        T[2] = (o >> 6) & 7u;
        T[3] = (o >> 9) & 7u;
        T[6] = (o >> 12) & 7u;
        const double mi = mi_3x3(sm.tab, T);
        val[p] = (unsigned long long)__double_as_longlong(mi);
      }
    }
  }
}

// exclusive prefix of the chunk emit counts (warp 0)
template <class SM>
__device__ __forceinline__ void fast_chunk_prefix(SM& sm, uint32_t n_chunks) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t carry = 0u;
  for (uint32_t c0 = 0; c0 < n_chunks; c0 += 32u) {
    const uint32_t c = c0 + lane;
    const uint32_t v = (c < n_chunks) ? (uint32_t)__popc(sm.emit_mask[c]) : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int)lane >= o) inc += t;
    }
    if (c < n_chunks) sm.chunk_off[c] = carry + inc - v;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) {
    sm.total = carry;
    sm.next_chunk = 0u;
  }
}

// ordered write of the surviving pairs: one 16-byte record {unit, i | j << 16, mi} each.
// Warps [first_warp, 8) share the chunks; the first two warps are busy with the mean phase.
template <class SM>
__device__ __forceinline__ void fast_emit(const FastParams& P, SM& sm, const unsigned long long* __restrict__ val,
                                          const FastItem& it, const uint16_t* __restrict__ ijt, unsigned long long base,
                                          uint32_t n_chunks, uint32_t first_warp) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  for (uint32_t c = warp - first_warp; c < n_chunks; c += kFastWarps - first_warp) {
    const uint32_t mask = sm.emit_mask[c];
    if ((mask >> lane) & 1u) {
      const uint32_t p = c * 32u + lane;
      const unsigned long long slot = base + sm.chunk_off[c] + __popc(mask & lt);
      const uint32_t ij = __ldg(ijt + p);
      const unsigned long long bits = val[p];
      uint4 rec;
      rec.x = it.unit + P.unit_base;
      rec.y = (ij >> 6) | ((ij & 63u) << 16);
      rec.z = (uint32_t)bits;
      rec.w = (uint32_t)(bits >> 32);
      reinterpret_cast<uint4*>(P.records)[slot] = rec;
    }
  }
}

// EMIT_COUNTS (test / audit mode): the 3x3 tables of the emitted pairs, from the
// packed counts and "other" cells the MI epilogue is about to consume (so what is
// checked bit-for-bit against the oracle is exactly what MI is computed from).
// Runs between the counts phase and fast_mi.
template <class SM>
__device__ __forceinline__ void fast_emit_counts(const FastParams& P, SM& sm, const unsigned long long* __restrict__ val,
                                                 unsigned long long base, uint32_t n_chunks) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  for (uint32_t c = warp; c < n_chunks; c += kFastWarps) {
    const uint32_t mask = sm.emit_mask[c];
    if ((mask >> lane) & 1u) {
      const uint32_t p = c * 32u + lane;
      const unsigned long long slot = base + sm.chunk_off[c] + __popc(mask & lt);
      const unsigned long long v = val[p];
      const uint32_t o = (uint32_t)(v >> 36);
      uint32_t* out = P.counts + slot * 9ull;
      uint32_t n_mm, n_mM, n_Mm, n_MM;
      unpack_counts(v, n_mm, n_mM, n_Mm, n_MM);
      out[0] = o & 7u;
      out[1] = (o >> 3) & 7u;
      out[2] = (o >> 6) & 7u;
      out[3] = (o >> 9) & 7u;
      out[4] = n_mm;
      out[5] = n_mM;
      out[6] = (o >> 12) & 7u;
      out[7] = n_Mm;
      out[8] = n_MM;
    }
  }
}

// per-site mean over the het-kept pairs (mutual_information.py:48-60): partners in
// ascending order, CPython's compensated float sum (all MI values are >= 0, so
// |s| >= |x| is s >= x).  Het sites sum over every partner, the others over the
// het sites only.  Threads [0, n_het) take the het sites (the long chains share
// warps), the next S - n_het threads the other sites.
// The rounding error of t = RN(s + x), exactly (Knuth's two-sum: six additions, no comparison).  CPython's
// compensated sum adds (s - t) + x when |s| >= |x| and (x - t) + s otherwise (Dekker's fast two-sum with the
// operands ordered): either way that IS the exact error of the addition, which is a double, so the three forms
// give the same bits -- and this one needs no select on the serial path.
__device__ __forceinline__ double two_sum_err(double s, double x, double t) {
  const double bp = __dsub_rn(t, s);
  return __dadd_rn(__dsub_rn(s, __dsub_rn(t, bp)), __dsub_rn(x, bp));
}

struct MeanAcc {
  double s, c;
  uint32_t n;
  // branch-free so that the loop unrolls and the loads run ahead of the serial sum: a pair
  // without MI (NaN pattern) adds +0.0, which leaves s and c as they are (s, c >= +0.0)
  __device__ __forceinline__ void add(const double* __restrict__ s_mi, uint32_t p) {
    const double v = s_mi[p];
    const bool have = __double2hiint(v) < 0x7ff00000;
    const double x = have ? v : 0.0;
    const double t = __dadd_rn(s, x);
    c = __dadd_rn(c, two_sum_err(s, x, t));
    s = t;
    n += have ? 1u : 0u;
  }
};

template <class SM>
__device__ __forceinline__ void fast_means(const FastParams& P, SM& sm, const unsigned long long* __restrict__ val,
                                           const FastItem& it) {
  const uint32_t tid = threadIdx.x, S = it.S;
  if (tid >= S) return;
  if (sm.n_list2 + sm.n_list3 == 0u) {  // no pair of the unit reached min_common: nothing to average
    P.site_mean[it.site_off + tid] = mi_nan();
    P.site_cnt[it.site_off + tid] = 0u;
    return;
  }
  const uint32_t n_het = __popcll(sm.het_mask);
  const double* s_mi = reinterpret_cast<const double*>(val);
  MeanAcc acc{0.0, 0.0, 0u};
  uint32_t s;
  if (tid < n_het) {
    s = sm.het_list[tid];
    uint32_t p = s - 1u;  // pair (t, s) for t = 0; the next one is S - t - 2 further
#pragma unroll 4
    for (uint32_t t = 0; t < s; ++t) {
      acc.add(s_mi, p);
      p += S - t - 2u;
    }
    p = (s * (2u * S - s - 1u)) >> 1;  // pair (s, s + 1)
#pragma unroll 4
    for (uint32_t t = s + 1u; t < S; ++t, ++p) acc.add(s_mi, p);
  } else {
    s = sm.nonhet_list[tid - n_het];
    const uint32_t row_s = (s * (2u * S - s - 1u)) >> 1;
#pragma unroll 4
    for (uint32_t q = 0; q < n_het; ++q) {
      const uint32_t t = sm.het_list[q];
      acc.add(s_mi, (t < s) ? (((t * (2u * S - t - 1u)) >> 1) + (s - t - 1u)) : (row_s + (t - s - 1u)));
    }
  }
  double mean = mi_nan();
  if (acc.n) {
    double tot = acc.s;
    if (acc.c != 0.0) tot = __dadd_rn(tot, acc.c);  // every term is finite, so is the compensation
    mean = __ddiv_rn(tot, (double)acc.n);
  }
  P.site_mean[it.site_off + s] = mean;
  P.site_cnt[it.site_off + s] = acc.n;
}



// kHetPairsOnly: the HET_ONLY | SKIP_NONHET instantiation (k_pairs_fast_het): counts by fast_counts_het
template <bool kHetPairsOnly>
__device__ __forceinline__ void pairs_fast_body(const FastParams& P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FastSmem& sm = *reinterpret_cast<FastSmem*>(smem_raw);
  const uint32_t tid = threadIdx.x;

  // per-CTA table: ln k (hi, lo) and RN(1/k) for k <= 256
  for (uint32_t k = tid; k <= (uint32_t)kFastMaxR; k += kFastThreads) {
    const double2 v = (k < P.ln_cap) ? __ldg(reinterpret_cast<const double2*>(P.lntab) + k) : make_double2(0.0, 0.0);
    sm.tab.ln[k].hi = v.x;
    sm.tab.ln[k].lo = v.y;
    sm.tab.inv[k] = k ? __drcp_rn((double)k) : 0.0;
  }

  constexpr bool het_pairs_only = kHetPairsOnly;
  uint32_t idx = blockIdx.x;
  uint32_t buf = 0;
  FastItem it;
  // an item the pre-pass counted no emitted pair for (no pair with enough common reads, or none next to a
  // het SNP in HET_ONLY mode) has no rows and only NaN means: it is neither fetched nor processed
  if (idx < P.n_items) {
    it = P.items[idx];
    if (!P.fast_empty[idx]) fast_prefetch(sm.rows[0], it, P.planes);
  }
  cp_async_commit();

  while (idx < P.n_items) {
    // descriptor + planes of the next item are requested before this one is touched
    const uint32_t idx_next = idx + gridDim.x;
    FastItem it_next;
    if (idx_next < P.n_items) {
      const bool empty_next = P.fast_empty[idx_next] != 0;  // (independent of the descriptor load: no added latency)
      it_next = P.items[idx_next];
      if (!empty_next) fast_prefetch(sm.rows[buf ^ 1u], it_next, P.planes);
    }
    cp_async_commit();
    const uint32_t S = it.S;
    const unsigned long long base = P.item_off[it.item];
    if (P.fast_empty[idx]) {  // touches no shared memory: no barrier
      if (tid < S) {
        P.site_mean[it.site_off + tid] = mi_nan();
        P.site_cnt[it.site_off + tid] = 0u;
      }
      if (tid == 0) P.unit_rec_off[it.unit] = base;
    } else {
      const uint32_t n_pairs = S * (S - 1u) / 2u;
      const uint32_t n_chunks = (n_pairs + 31u) >> 5;
      uint32_t* rows = sm.rows[buf];
      const uint16_t* __restrict__ ijt = P.ij_tab + lg_ij_tab_off(S);
      if (tid < (uint32_t)kFastMaxS) sm.n_oth[tid] = 0u;
      for (uint32_t q = tid; q < (n_pairs + 1u) >> 1; q += kFastThreads) sm.ocell[q] = 0u;
      if (tid == 0) {
        sm.n_list2 = 0u;
        sm.n_list3 = 0u;
        sm.n_oth_total = 0u;
      }
      if (het_pairs_only) {  // fast_counts_het only touches the pairs next to a het site
        for (uint32_t q = tid; q < n_pairs; q += kFastThreads) sm.val[q] = kNoMi;
        if (tid < n_chunks) sm.emit_mask[tid] = 0u;
      }
      cp_async_wait<1>();  // this item's rows have landed (the next item's may still be in flight)
      __syncthreads();
      fast_land(sm, rows, it, P.site_flags + it.site_off);
      __syncthreads();
      // a site with more than kOthCap "other" reads does not fit the sparse lists: the generic kernel takes the unit
      const bool over = tid < S && sm.n_oth[tid] > (uint32_t)kOthCap;
      fast_site_lists(sm, S);
      fast_other_cells<!het_pairs_only>(sm, rows, S, het_pairs_only ? sm.het_mask : ~0ull);
      bool dense;
      dense = __syncthreads_or(over) != 0;  // (also: sm.info is read by every thread of the counts phase)
      if (dense) {
        if (tid == 0) {
          P.item_dense[it.item] = 1;
          atomicAdd(P.n_generic, 1u);
        }
      } else {
        const uint32_t nw = ((uint32_t)it.R + 31u) >> 5;
        if (het_pairs_only) {
          if (nw <= 4u) fast_counts_het<4>(P, sm, rows, S);
          else fast_counts_het<8>(P, sm, rows, S);
        } else if (nw <= 2u) fast_counts<2>(P, sm, rows, ijt, n_pairs);
        else if (nw <= 4u) fast_counts<4>(P, sm, rows, ijt, n_pairs);
        else if (nw <= 7u) fast_counts<7>(P, sm, rows, ijt, n_pairs);
        else fast_counts<8>(P, sm, rows, ijt, n_pairs);
        __syncthreads();
        if (!het_pairs_only) fast_build_lists(sm, n_chunks);  // (fast_counts_het appends to the lists itself)
        if (tid < 32u) fast_chunk_prefix(sm, n_chunks);
        if (!het_pairs_only || (P.mode & LGMI_MODE_EMIT_COUNTS)) __syncthreads();
        if (P.mode & LGMI_MODE_EMIT_COUNTS) {
          fast_emit_counts(P, sm, sm.val, base, n_chunks);
          __syncthreads();  // the packed counts have been written out: fast_mi may replace them by MI values
        }
        fast_mi(sm, sm.val);
        __syncthreads();
        if (tid == 0) P.unit_rec_off[it.unit] = base;
        // the serial per-site sums occupy the first one or two warps; the others write the records
        const uint32_t mean_warps = (S + 31u) >> 5;
        if ((tid >> 5) < mean_warps) fast_means(P, sm, sm.val, it);
        else fast_emit(P, sm, sm.val, it, ijt, base, n_chunks, mean_warps);
      }
      __syncthreads();  // everything of this item consumed before its buffers are reused
    }
    it = it_next;
    idx = idx_next;
    buf ^= 1u;
  }
  cp_async_wait<0>();
}


__global__ void __launch_bounds__(kFastThreads, 4) k_pairs_fast(const FastParams P) { pairs_fast_body<false>(P); }
// launched instead of k_pairs_fast when the mode is HET_ONLY | SKIP_NONHET
__global__ void __launch_bounds__(kFastThreads, 4) k_pairs_fast_het(const FastParams P) { pairs_fast_body<true>(P); }

}  // namespace lgmi
