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
  int min_common;
  uint32_t mode;
  const unsigned long long* item_off;  // exclusive scan of the per-item emit counts
  const uint8_t* item_dense;           // 1: some site has more than kOthCap "other" reads -> generic kernel
  lgmi_pair_rec* records;
  uint32_t* counts;                    // EMIT_COUNTS: 9 per record
  double* site_mean;
  uint32_t* site_cnt;
  unsigned long long* unit_rec_off;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool copy) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = copy ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// issue the loads of one unit's plane rows into a landing buffer: 6 x 16 B per site
__device__ __forceinline__ void fast_prefetch(uint32_t* __restrict__ rows, const FastItem& it,
                                              const uint32_t* __restrict__ planes) {
  const uint32_t W = it.W, W4 = W >> 2;
  const uint32_t* __restrict__ src = planes + it.plane_off;
  for (uint32_t e = threadIdx.x; e < (uint32_t)it.S * 6u; e += kFastThreads) {
    const uint32_t s = e / 6u, q = e - s * 6u;
    const uint32_t plane = q >> 1, half = q & 1u;
    const bool have = half < W4;
    cp_async16(rows + s * kRowStride + plane * 8u + half * 4u,
               src + (size_t)s * 3u * W + plane * W + (have ? half * 4u : 0u), have);
  }
}

// in-place transform of the landed rows + "other" lists + site flags
__device__ __forceinline__ void fast_land(FastSmem& sm, uint32_t* __restrict__ rows, const FastItem& it,
                                          const uint8_t* __restrict__ flags) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t S = it.S;
  for (uint32_t e = tid; e < S * 8u; e += kFastThreads) {
    const uint32_t s = e >> 3, k = e & 7u;
    uint32_t* row = rows + s * kRowStride;
    const uint32_t C = row[16 + k];
    const uint32_t M = row[k] & C;
    const uint32_t Pw = (M | row[8 + k]) & C;
    row[k] = M;
    row[8 + k] = Pw;
    uint32_t O = C & ~Pw;
    while (O) {
      const uint32_t b = __ffs((int)O) - 1u;
      O &= O - 1u;
      const uint32_t idx = atomicAdd(&sm.n_oth[s], 1u);
      if (idx < (uint32_t)kOthCap) sm.oth_list[s * 8u + idx] = (uint16_t)(k * 32u + b);
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

// list of the sites that have "other" reads (any order)
__device__ __forceinline__ void fast_other_sites(FastSmem& sm, uint32_t S) {
  const uint32_t tid = threadIdx.x;
  if (tid < S && sm.n_oth[tid] != 0u) sm.oth_sites[atomicAdd(&sm.n_oth_sites, 1u)] = (uint8_t)tid;
}

// the cells that involve an "other" label, from the sparse lists.
// field f of sm.oth[p] (3 bits each): 0 T[0][0], 1 T[0][1], 2 T[0][2], 3 T[1][0], 4 T[2][0]
__device__ __forceinline__ void fast_fixup(FastSmem& sm, const uint32_t* __restrict__ rows, uint32_t S) {
  const uint32_t n_items = sm.n_oth_sites * 64u;
  for (uint32_t e = threadIdx.x; e < n_items; e += kFastThreads) {
    const uint32_t a = sm.oth_sites[e >> 6], b = e & 63u;
    if (b >= S || a == b) continue;
    const uint32_t na = sm.n_oth[a];
    uint32_t add = 0u;
    for (uint32_t q = 0; q < na; ++q) {
      const uint32_t r = sm.oth_list[a * 8u + q];
      const uint32_t w = r >> 5, bit = r & 31u;
      if ((rows[b * kRowStride + 8u + w] >> bit) & 1u) {
        const uint32_t label_b = 1u + ((rows[b * kRowStride + w] >> bit) & 1u);  // 1 minor, 2 major
        add += 1u << (3u * ((a < b) ? label_b : (2u + label_b)));
      } else if (a < b && ((rows[b * kRowStride + 16u + w] >> bit) & 1u)) {
        add += 1u;  // "other" at both sites: counted once, from the lower site
      }
    }
    if (add) {
      const uint32_t i = min(a, b), j = max(a, b);
      const uint32_t p = ((i * (2u * S - i - 1u)) >> 1) + (j - i - 1u);
      // 16-bit cells packed two per word: add into the right half
      atomicAdd(reinterpret_cast<uint32_t*>(sm.oth) + (p >> 1), add << ((p & 1u) * 16u));
    }
  }
}

// the four AND+popcount sets of every pair, min-common filter, emit masks, lists
template <int NW>
__device__ __forceinline__ void fast_counts(const FastParams& P, FastSmem& sm, const uint32_t* __restrict__ rows,
                                            uint32_t S, uint32_t n_pairs) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const bool het_only = (P.mode & LGMI_MODE_HET_ONLY) != 0u;
  const bool skip_nonhet = het_only && (P.mode & LGMI_MODE_SKIP_NONHET) != 0u;
  const uint32_t n_slots = (n_pairs + 31u) & ~31u;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t i = 0, k = tid, len = S - 1u;  // lexicographic walk: p -> (i, j = i + 1 + k)
  for (uint32_t p = tid; p < n_slots; p += kFastThreads) {
    uint32_t cls = 0u;  // 0 dropped, 2 -> 2x2 list, 3 -> 3x3 list
    bool emit = false;
    if (p < n_pairs) {
      while (k >= len) {
        k -= len;
        ++i;
        --len;
      }
      const uint32_t j = i + 1u + k;
      sm.ij[p] = (uint16_t)(i * 64u + j);
      const bool het = ((sm.flags[i] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP) ||
                       ((sm.flags[j] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP);
      unsigned long long v = 0x7ff8000000000000ull;  // NaN: no MI for this candidate
      if (het || !skip_nonhet) {
        const unsigned long long cnt = pair_counts<NW>(rows + i * kRowStride, rows + j * kRowStride);
        const uint32_t o = sm.oth[p];
        const uint32_t n_oth = (o & 7u) + ((o >> 3) & 7u) + ((o >> 6) & 7u) + ((o >> 9) & 7u) + ((o >> 12) & 7u);
        const uint32_t n_common = (uint32_t)(cnt & 0xffffu) + n_oth;
        if ((int)n_common >= P.min_common) {  // strict '<' drops (mutual_information.py:19)
          v = cnt;
          cls = o ? 3u : 2u;
          emit = het || !het_only;
        }
      }
      sm.val[p] = v;
    }
    k += kFastThreads;
    // warp-aggregated bookkeeping: this warp's 32 pairs are one lexicographic chunk
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
}

// MI of the listed pairs, in place of their packed counts; warp-sized chunks of both lists
__device__ __forceinline__ void fast_mi(FastSmem& sm) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t n2 = sm.n_list2, n3 = sm.n_list3;
  const uint32_t nc2 = (n2 + 31u) >> 5, nc3 = (n3 + 31u) >> 5;
  for (uint32_t c = warp; c < nc2 + nc3; c += kFastWarps) {
    if (c < nc2) {
      const uint32_t q = c * 32u + lane;
      if (q < n2) {
        const uint32_t p = sm.list[q];
        const unsigned long long cnt = sm.val[p];
        const uint32_t nPP = (uint32_t)(cnt & 0xffffu), nMP = (uint32_t)((cnt >> 16) & 0xffffu);
        const uint32_t nPM = (uint32_t)((cnt >> 32) & 0xffffu), nMM = (uint32_t)(cnt >> 48);
        // i major & j minor = nMP - nMM ; i minor & j major = nPM - nMM
        const double mi = mi_2x2(sm.tab, nPP - nMP - nPM + nMM, nPM - nMM, nMP - nMM, nMM);
        sm.val[p] = (unsigned long long)__double_as_longlong(mi);
      }
    } else {
      const uint32_t q = (c - nc2) * 32u + lane;
      if (q < n3) {
        const uint32_t p = sm.list[kFastMaxPairs - 1u - q];
        const unsigned long long cnt = sm.val[p];
        const uint32_t o = sm.oth[p];
        const uint32_t nPP = (uint32_t)(cnt & 0xffffu), nMP = (uint32_t)((cnt >> 16) & 0xffffu);
        const uint32_t nPM = (uint32_t)((cnt >> 32) & 0xffffu), nMM = (uint32_t)(cnt >> 48);
        uint32_t T[9];
        T[0] = o & 7u;
        T[1] = (o >> 3) & 7u;
        T[2] = (o >> 6) & 7u;
        T[3] = (o >> 9) & 7u;
        T[6] = (o >> 12) & 7u;
        T[4] = nPP - nMP - nPM + nMM;
        T[5] = nPM - nMM;
        T[7] = nMP - nMM;
        T[8] = nMM;
        const double mi = mi_3x3(sm.tab, T);
        sm.val[p] = (unsigned long long)__double_as_longlong(mi);
      }
    }
  }
}

// exclusive prefix of the chunk emit counts (warp 0)
__device__ __forceinline__ void fast_chunk_prefix(FastSmem& sm, uint32_t n_chunks) {
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
  if (lane == 0) sm.total = carry;
}

// ordered write of the surviving pairs: one 16-byte record {unit, i | j << 16, mi} each
__device__ __forceinline__ void fast_emit(const FastParams& P, FastSmem& sm, const FastItem& it,
                                          unsigned long long base, uint32_t n_chunks) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  for (uint32_t c = warp; c < n_chunks; c += kFastWarps) {
    const uint32_t mask = sm.emit_mask[c];
    if ((mask >> lane) & 1u) {
      const uint32_t p = c * 32u + lane;
      const unsigned long long slot = base + sm.chunk_off[c] + __popc(mask & lt);
      const uint32_t ij = sm.ij[p];
      const unsigned long long bits = sm.val[p];
      uint4 rec;
      rec.x = it.unit;
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
__device__ __forceinline__ void fast_emit_counts(const FastParams& P, FastSmem& sm, unsigned long long base,
                                                 uint32_t n_chunks) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  for (uint32_t c = warp; c < n_chunks; c += kFastWarps) {
    const uint32_t mask = sm.emit_mask[c];
    if ((mask >> lane) & 1u) {
      const uint32_t p = c * 32u + lane;
      const unsigned long long slot = base + sm.chunk_off[c] + __popc(mask & lt);
      const unsigned long long cnt = sm.val[p];
      const uint32_t o = sm.oth[p];
      const uint32_t nPP = (uint32_t)(cnt & 0xffffu), nMP = (uint32_t)((cnt >> 16) & 0xffffu);
      const uint32_t nPM = (uint32_t)((cnt >> 32) & 0xffffu), nMM = (uint32_t)(cnt >> 48);
      uint32_t* out = P.counts + slot * 9ull;
      out[0] = o & 7u;
      out[1] = (o >> 3) & 7u;
      out[2] = (o >> 6) & 7u;
      out[3] = (o >> 9) & 7u;
      out[4] = nPP - nMP - nPM + nMM;
      out[5] = nPM - nMM;
      out[6] = (o >> 12) & 7u;
      out[7] = nMP - nMM;
      out[8] = nMM;
    }
  }
}

// per-site mean over the het-kept pairs (mutual_information.py:48-60): partners in
// ascending order, CPython's compensated float sum.  Het sites sum over every
// partner, the others over the het sites only.
__device__ __forceinline__ uint32_t nth_set_bit(unsigned long long m, uint32_t n) {
  const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
  const uint32_t c = __popc(lo);
  return (n < c) ? __fns(lo, 0, n + 1) : 32u + __fns(hi, 0, n - c + 1);
}

__device__ __forceinline__ void fast_means(const FastParams& P, FastSmem& sm, const FastItem& it) {
  const uint32_t tid = threadIdx.x, S = it.S;
  const unsigned long long het = sm.het_mask;
  const uint32_t n_het = __popcll(het);
  const unsigned long long all = (S == 64u) ? ~0ull : ((1ull << S) - 1ull);
  // threads [0, n_het): het sites (long chains, packed into the first warps); then the other sites
  uint32_t s;
  unsigned long long partners;
  if (tid < n_het) {
    s = nth_set_bit(het, tid);
    partners = all & ~(1ull << s);
  } else if (tid < S) {
    s = nth_set_bit(all & ~het, tid - n_het);
    partners = het;
  } else {
    return;
  }
  const double* s_mi = reinterpret_cast<const double*>(sm.val);
  const uint32_t row_s = (s * (2u * S - s - 1u)) >> 1;
  lg_neumaier acc;
  lg_neumaier_init(acc);
  while (partners) {
    const uint32_t t = __ffsll((long long)partners) - 1u;
    partners &= partners - 1ull;
    const uint32_t p = (t < s) ? (((t * (2u * S - t - 1u)) >> 1) + (s - t - 1u)) : (row_s + (t - s - 1u));
    const double v = s_mi[p];
    if (__double2hiint(v) != 0x7ff80000) lg_neumaier_add(acc, v);  // NaN pattern: no MI for this pair
  }
  P.site_mean[it.site_off + s] = lg_neumaier_mean(acc);
  P.site_cnt[it.site_off + s] = (uint32_t)acc.n;
}

__global__ void __launch_bounds__(kFastThreads, 4) k_pairs_fast(const FastParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FastSmem& sm = *reinterpret_cast<FastSmem*>(smem_raw);
  const uint32_t tid = threadIdx.x;

  // per-CTA table: ln k (hi, lo), (double)k, RN(1/k) for k <= 256
  for (uint32_t k = tid; k <= (uint32_t)kFastMaxR; k += kFastThreads) {
    FastTabEntry e;
    const double2 v = (k < P.ln_cap) ? __ldg(reinterpret_cast<const double2*>(P.lntab) + k) : make_double2(0.0, 0.0);
    e.ln_hi = v.x;
    e.ln_lo = v.y;
    e.dk = (double)k;
    e.inv = k ? __drcp_rn((double)k) : 0.0;
    sm.tab[k] = e;
  }

  uint32_t idx = blockIdx.x;
  uint32_t buf = 0;
  FastItem it;
  if (idx < P.n_items) {
    it = P.items[idx];
    fast_prefetch(sm.rows[0], it, P.planes);
  }
  cp_async_commit();

  while (idx < P.n_items) {
    // descriptor + planes of the next item are requested before this one is touched
    const uint32_t idx_next = idx + gridDim.x;
    FastItem it_next;
    if (idx_next < P.n_items) {
      it_next = P.items[idx_next];
      fast_prefetch(sm.rows[buf ^ 1u], it_next, P.planes);
    }
    cp_async_commit();

    const uint32_t S = it.S;
    const uint32_t n_pairs = S * (S - 1u) / 2u;
    const uint32_t n_chunks = (n_pairs + 31u) >> 5;
    const bool dense = P.item_dense[it.item] != 0u;  // handled by the generic kernel
    if (!dense) {
      uint32_t* rows = sm.rows[buf];
      for (uint32_t p = tid; p < (n_chunks * 32u + 1u) / 2u; p += kFastThreads) reinterpret_cast<uint32_t*>(sm.oth)[p] = 0u;
      if (tid < (uint32_t)kFastMaxS) sm.n_oth[tid] = 0u;
      if (tid == 0) {
        sm.n_list2 = 0u;
        sm.n_list3 = 0u;
        sm.n_oth_sites = 0u;
      }
      cp_async_wait<1>();  // this item's rows have landed (the next item's may still be in flight)
      __syncthreads();
      fast_land(sm, rows, it, P.site_flags + it.site_off);
      __syncthreads();
      fast_other_sites(sm, S);
      __syncthreads();
      fast_fixup(sm, rows, S);
      __syncthreads();
      const uint32_t nw = ((uint32_t)it.R + 31u) >> 5;
      if (nw <= 2u) fast_counts<2>(P, sm, rows, S, n_pairs);
      else if (nw <= 4u) fast_counts<4>(P, sm, rows, S, n_pairs);
      else if (nw <= 7u) fast_counts<7>(P, sm, rows, S, n_pairs);
      else fast_counts<8>(P, sm, rows, S, n_pairs);
      __syncthreads();
      const unsigned long long base = P.item_off[it.item];
      if (tid < 32u) fast_chunk_prefix(sm, n_chunks);
      if (P.mode & LGMI_MODE_EMIT_COUNTS) {
        __syncthreads();
        fast_emit_counts(P, sm, base, n_chunks);
      }
      fast_mi(sm);
      __syncthreads();
      if (tid == 0) P.unit_rec_off[it.unit] = base;
      fast_emit(P, sm, it, base, n_chunks);
      fast_means(P, sm, it);
    } else {
      cp_async_wait<1>();
    }
    __syncthreads();  // everything of this item consumed before its buffers are reused
    it = it_next;
    idx = idx_next;
    buf ^= 1u;
  }
  cp_async_wait<0>();
}

}  // namespace lgmi
