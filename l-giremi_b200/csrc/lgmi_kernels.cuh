// lgmi_kernels.cuh -- sm_100a kernels of the MI step.
//
//   k_pairs            K1+K2: per-pair AND+popcount contingency counts over the
//                      common reads, fused fp64 MI epilogue, min-common filter,
//                      het filter, ORDERED compaction of the surviving pairs
//                      (single pass, decoupled look-back across CTAs) and, for
//                      units that fit one CTA, the per-site mean MI.
//   k_site_mean_dense  per-site mean MI for units spanning several CTAs.
//   k_site_mean_csr    mean of caller-supplied rows (drop-in for
//                      mean_mismatch_pair_mutual_info).
//   k_ecdf_*           K4: het-SNP mean collection, searchsorted-left, mip, call.
//
// Reference semantics: /root/reference/src/giremi/mutual_information.py:6-60,
// mismatch.py:393-396, stat.py:7-29, script/giremi.py:97-114,415-429.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lgmi.h"
#include "lgmi_math.cuh"

namespace lgmi {

constexpr int kThreads = 256;   // CTA size of k_pairs
constexpr int kPairsMax = 2048; // pairs per work item (>= 64*63/2: a 64-site unit is one item)

struct DevUnit {
  uint64_t plane_off;  // words
  uint64_t dense_off;  // first slot in the dense MI scratch (multi-item units), else ~0
  uint32_t S, R, W, site_off;
  uint32_t first_item, n_items;
};

enum : uint32_t { ITEM_FIRST = 1u, ITEM_SINGLE = 2u };

struct Item {
  uint32_t unit;
  uint32_t pair_begin;
  uint32_t pair_cnt;
  uint32_t flags;
};

struct Header {
  unsigned long long n_records;
  unsigned int ticket;
  unsigned int pad;
};

struct RunParams {
  const DevUnit* units;
  const Item* items;
  uint32_t n_items;
  uint32_t n_units;
  const uint32_t* planes;
  const uint8_t* site_flags;
  const lg_dd* lntab;
  int min_common;
  uint32_t mode;
  unsigned long long* status;  // per item: look-back word
  Header* header;
  lgmi_pair_rec* records;
  uint32_t* counts;
  double* site_mean;
  uint32_t* site_cnt;
  double* dense;
  unsigned long long* unit_rec_off;
};

struct LnGlobal {
  const lg_dd* tab;
  __device__ __forceinline__ lg_dd operator()(uint32_t k) const {
    const double2 v = __ldg(reinterpret_cast<const double2*>(tab) + k);
    lg_dd r;
    r.hi = v.x;
    r.lo = v.y;
    return r;
  }
};

__device__ __forceinline__ uint32_t popc4(uint4 a, uint4 b) {
  return __popc(a.x & b.x) + __popc(a.y & b.y) + __popc(a.z & b.z) + __popc(a.w & b.w);
}

__device__ __forceinline__ double lg_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

// Common reads of sites i and j: |C_i & C_j|  (mutual_information.py:17-19).
__device__ __forceinline__ uint32_t pair_common(const uint4* __restrict__ ri,
                                                const uint4* __restrict__ rj, uint32_t W4) {
  uint32_t n = 0;
  for (uint32_t k = 0; k < W4; ++k) n += popc4(__ldg(ri + 2 * W4 + k), __ldg(rj + 2 * W4 + k));
  return n;
}

// Fills the 3x3 table (label order other/minor/major) of a pair whose common
// read count is already known.  Returns true when no "other" label occurs
// among the common reads (the table is then 2x2 in the minor/major block).
__device__ __forceinline__ bool pair_table(const uint4* __restrict__ ri, const uint4* __restrict__ rj,
                                           uint32_t W4, uint32_t n_common, uint32_t T[9]) {
  uint32_t MM = 0, Mm = 0, mM = 0, mm = 0;
  for (uint32_t k = 0; k < W4; ++k) {
    const uint4 Mi = __ldg(ri + k), mi = __ldg(ri + W4 + k);
    const uint4 Mj = __ldg(rj + k), mj = __ldg(rj + W4 + k);
    MM += popc4(Mi, Mj);
    Mm += popc4(Mi, mj);
    mM += popc4(mi, Mj);
    mm += popc4(mi, mj);
  }
  T[8] = MM;
  T[7] = Mm;
  T[5] = mM;
  T[4] = mm;
  if (MM + Mm + mM + mm == n_common) {
    T[0] = T[1] = T[2] = T[3] = T[6] = 0;
    return true;
  }
  uint32_t aM = 0, am = 0, bM = 0, bm = 0;  // marginals against the partner's coverage
  for (uint32_t k = 0; k < W4; ++k) {
    const uint4 Mi = __ldg(ri + k), mi = __ldg(ri + W4 + k), Ci = __ldg(ri + 2 * W4 + k);
    const uint4 Mj = __ldg(rj + k), mj = __ldg(rj + W4 + k), Cj = __ldg(rj + 2 * W4 + k);
    aM += popc4(Mi, Cj);
    am += popc4(mi, Cj);
    bM += popc4(Ci, Mj);
    bm += popc4(Ci, mj);
  }
  T[6] = aM - MM - Mm;  // site1 major, site2 other
  T[3] = am - mM - mm;  // site1 minor, site2 other
  T[2] = bM - MM - mM;  // site1 other, site2 major
  T[1] = bm - Mm - mm;  // site1 other, site2 minor
  T[0] = n_common - (MM + Mm + mM + mm) - T[6] - T[3] - T[2] - T[1];
  return false;
}

// ---------------------------------------------------------------------------
// look-back status word: [63:62] flag, [61:0] value
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValMask = (1ull << 62) - 1ull;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exclusive prefix of `count` over items in ticket order.  Called by warp 0.
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long* status,
                                                                 uint32_t item,
                                                                 unsigned long long count) {
  const uint32_t lane = threadIdx.x & 31u;
  if (item == 0) {
    if (lane == 0) st_status(status, kFlagPrefix | count);
    return 0ull;
  }
  if (lane == 0) st_status(status + item, kFlagAgg | count);
  unsigned long long excl = 0ull;
  int64_t base = (int64_t)item - 1;
  while (true) {
    const int64_t idx = base - (int64_t)lane;
    unsigned long long st = kFlagPrefix;  // virtual zero prefix before item 0
    if (idx >= 0) {
      st = ld_status(status + idx);
      while ((st >> 62) == 0ull) {
        __nanosleep(20);
        st = ld_status(status + idx);
      }
    }
    const uint32_t pm = __ballot_sync(0xffffffffu, (st >> 62) == 2ull);
    const uint32_t first = pm ? (uint32_t)(__ffs((int)pm) - 1) : 32u;
    unsigned long long v = (lane <= first) ? (st & kValMask) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    if (pm) break;
    base -= 32;
  }
  if (lane == 0) st_status(status + item, kFlagPrefix | (excl + count));
  return excl;
}

// ---------------------------------------------------------------------------
// K1 + K2.  One CTA per work item, items taken in ticket order.
__global__ void __launch_bounds__(kThreads) k_pairs(const RunParams P) {
  __shared__ double s_mi[kPairsMax];     // MI of each pair of the item, NaN = dropped
  __shared__ uint32_t s_ij[kPairsMax];   // (i << 16) | j
  __shared__ uint32_t s_warp[kThreads / 32];
  __shared__ uint32_t s_item;
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_total;

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) s_item = atomicAdd(&P.header->ticket, 1u);
  __syncthreads();
  const uint32_t item_idx = s_item;
  if (item_idx >= P.n_items) return;
  const Item it = P.items[item_idx];
  const DevUnit u = P.units[it.unit];
  const uint32_t W4 = u.W >> 2;
  const uint4* __restrict__ base = reinterpret_cast<const uint4*>(P.planes + u.plane_off);
  const uint8_t* __restrict__ flags = P.site_flags + u.site_off;
  const LnGlobal ln{P.lntab};
  const bool het_only = (P.mode & LGMI_MODE_HET_ONLY) != 0u;
  const bool skip_nonhet = het_only && (P.mode & LGMI_MODE_SKIP_NONHET) != 0u;

  // ---- phase 1: counts + MI for every candidate pair of the item
  for (uint32_t pl = tid; pl < it.pair_cnt; pl += kThreads) {
    uint32_t i, j;
    lg_pair_ij(it.pair_begin + pl, u.S, i, j);
    s_ij[pl] = (i << 16) | j;
    double mi = lg_nan();
    bool evaluate = true;
    if (skip_nonhet) {
      evaluate = ((flags[i] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP) ||
                 ((flags[j] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP);
    }
    if (evaluate) {
      const uint4* ri = base + (size_t)i * 3u * W4;
      const uint4* rj = base + (size_t)j * 3u * W4;
      const uint32_t n_common = pair_common(ri, rj, W4);
      if ((int)n_common >= P.min_common) {  // strict '<' drops (mutual_information.py:19)
        uint32_t T[9];
        if (pair_table(ri, rj, W4, n_common, T))
          mi = lg_mi_from_2x2(T[4], T[5], T[7], T[8], ln);
        else
          mi = lg_mi_from_table(T, ln);
      }
    }
    s_mi[pl] = mi;
  }
  __syncthreads();

  // ---- phase 2: how many pairs does this item emit?
  uint32_t my_emit = 0;
  for (uint32_t pl = tid; pl < it.pair_cnt; pl += kThreads) {
    bool e = !isnan(s_mi[pl]);
    if (e && het_only) {
      const uint32_t ij = s_ij[pl];
      e = ((flags[ij >> 16] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP) ||
          ((flags[ij & 0xffffu] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP);
    }
    my_emit += e ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_emit += __shfl_xor_sync(0xffffffffu, my_emit, o);
  if (lane == 0) s_warp[warp] = my_emit;
  __syncthreads();
  if (warp == 0) {
    uint32_t tot = (lane < kThreads / 32) ? s_warp[lane] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    const unsigned long long excl = lookback_exclusive(P.status, item_idx, tot);
    if (lane == 0) {
      s_base = excl;
      s_total = tot;
      if (it.flags & ITEM_FIRST) P.unit_rec_off[it.unit] = excl;
      if (item_idx == P.n_items - 1) {
        P.header->n_records = excl + tot;
        P.unit_rec_off[P.n_units] = excl + tot;
      }
    }
  }
  __syncthreads();

  // ---- phase 3: ordered write of the emitted pairs
  unsigned long long out = s_base;
  if (s_total != 0u) {
    for (uint32_t c0 = 0; c0 < it.pair_cnt; c0 += kThreads) {
      const uint32_t pl = c0 + tid;
      bool e = false;
      double mi = 0.0;
      uint32_t ij = 0;
      if (pl < it.pair_cnt) {
        mi = s_mi[pl];
        ij = s_ij[pl];
        e = !isnan(mi);
        if (e && het_only)
          e = ((flags[ij >> 16] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP) ||
              ((flags[ij & 0xffffu] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP);
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, e);
      if (lane == 0) s_warp[warp] = __popc(bal);
      __syncthreads();
      uint32_t before = 0, chunk_total = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) {
        const uint32_t c = s_warp[w];
        before += (w < (int)warp) ? c : 0u;
        chunk_total += c;
      }
      if (e) {
        const unsigned long long slot = out + before + __popc(bal & ((1u << lane) - 1u));
        lgmi_pair_rec r;
        r.unit = it.unit;
        r.i = (uint16_t)(ij >> 16);
        r.j = (uint16_t)(ij & 0xffffu);
        r.mi = mi;
        P.records[slot] = r;
        if (P.mode & LGMI_MODE_EMIT_COUNTS) {
          const uint4* ri = base + (size_t)(ij >> 16) * 3u * W4;
          const uint4* rj = base + (size_t)(ij & 0xffffu) * 3u * W4;
          uint32_t T[9];
          pair_table(ri, rj, W4, pair_common(ri, rj, W4), T);
#pragma unroll
          for (int k = 0; k < 9; ++k) P.counts[slot * 9ull + k] = T[k];
        }
      }
      out += chunk_total;
      __syncthreads();
    }
  }

  // ---- phase 4: per-site mean over the het-kept pairs (mutual_information.py:48-60)
  if (it.flags & ITEM_SINGLE) {
    for (uint32_t s = tid; s < u.S; s += kThreads) {
      const bool s_het = (flags[s] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP;
      lg_neumaier acc;
      lg_neumaier_init(acc);
      for (uint32_t t = 0; t < u.S; ++t) {
        if (t == s) continue;
        if (!s_het && (flags[t] & LGMI_SITE_TYPE_MASK) != LGMI_SITE_HET_SNP) continue;
        const uint32_t p = (t < s) ? (uint32_t)lg_row_off(t, u.S) + (s - t - 1u)
                                   : (uint32_t)lg_row_off(s, u.S) + (t - s - 1u);
        const double v = s_mi[p];
        if (!isnan(v)) lg_neumaier_add(acc, v);
      }
      P.site_mean[u.site_off + s] = lg_neumaier_mean(acc);
      P.site_cnt[u.site_off + s] = (uint32_t)acc.n;
    }
  } else {
    for (uint32_t pl = tid; pl < it.pair_cnt; pl += kThreads)
      P.dense[u.dense_off + it.pair_begin + pl] = s_mi[pl];
  }
}

// Per-site mean for units whose pairs span several work items; reads the dense
// per-unit MI scratch written by k_pairs.  One thread per site.
struct MeanItem {
  uint32_t unit;
  uint32_t site_begin;
};
__global__ void __launch_bounds__(128) k_site_mean_dense(const DevUnit* __restrict__ units,
                                                         const MeanItem* __restrict__ items,
                                                         const uint8_t* __restrict__ site_flags,
                                                         const double* __restrict__ dense,
                                                         double* __restrict__ site_mean,
                                                         uint32_t* __restrict__ site_cnt) {
  const MeanItem mi = items[blockIdx.x];
  const DevUnit u = units[mi.unit];
  const uint32_t s = mi.site_begin + threadIdx.x;
  if (s >= u.S) return;
  const uint8_t* flags = site_flags + u.site_off;
  const double* d = dense + u.dense_off;
  const bool s_het = (flags[s] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP;
  lg_neumaier acc;
  lg_neumaier_init(acc);
  for (uint32_t t = 0; t < u.S; ++t) {
    if (t == s) continue;
    if (!s_het && (flags[t] & LGMI_SITE_TYPE_MASK) != LGMI_SITE_HET_SNP) continue;
    const uint64_t p = (t < s) ? lg_row_off(t, u.S) + (s - t - 1u) : lg_row_off(s, u.S) + (t - s - 1u);
    const double v = d[p];
    if (!isnan(v)) lg_neumaier_add(acc, v);
  }
  site_mean[u.site_off + s] = lg_neumaier_mean(acc);
  site_cnt[u.site_off + s] = (uint32_t)acc.n;
}

// sites of units without any pair (S < 2): mean = NaN, cnt = 0
__global__ void k_fill_nan(double* __restrict__ mean, uint32_t* __restrict__ cnt, uint64_t n) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    mean[k] = lg_nan();
    cnt[k] = 0u;
  }
}

// mean_mismatch_pair_mutual_info on caller-supplied rows (CSR by site)
__global__ void k_site_mean_csr(const unsigned long long* __restrict__ off,
                                const double* __restrict__ val, uint64_t n, double* __restrict__ out) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  lg_neumaier acc;
  lg_neumaier_init(acc);
  for (unsigned long long k = off[s]; k < off[s + 1]; ++k) lg_neumaier_add(acc, val[k]);
  out[s] = lg_neumaier_mean(acc);
}

// ---------------------------------------------------------------------------
// K4: ecdf / mip / call
__global__ void k_ecdf_keys(const double* __restrict__ mean, const uint8_t* __restrict__ flags,
                            uint64_t n, double* __restrict__ keys, unsigned long long* __restrict__ n_het) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool het = false;
  if (k < n) {
    const double v = mean[k];
    het = !isnan(v) && ((flags[k] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_HET_SNP);
    keys[k] = het ? v : __longlong_as_double(0x7ff0000000000000LL);  // +inf sorts last
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, het);
  if ((threadIdx.x & 31u) == 0 && bal) atomicAdd(n_het, (unsigned long long)__popc(bal));
}

__device__ __forceinline__ uint64_t lower_bound(const double* __restrict__ x, uint64_t n, double v) {
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (x[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void k_ecdf_mip(const double* __restrict__ mean, const uint8_t* __restrict__ flags,
                           uint64_t n, const double* __restrict__ sorted_keys,
                           const unsigned long long* __restrict__ n_het_p, double threshold,
                           double* __restrict__ mip, uint8_t* __restrict__ call) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint64_t n_het = *n_het_p;
  const double v = mean[k];
  double p = lg_nan();
  uint8_t c = 0;
  if (!isnan(v) && n_het > 0) {
    p = lg_ecdf_y(lower_bound(sorted_keys, n_het, v), n_het);
    const bool is_mm = (flags[k] & LGMI_SITE_TYPE_MASK) == LGMI_SITE_MISMATCH;
    if (p <= threshold && is_mm) c = 1;
    else if (p > threshold && !is_mm) c = 2;
  }
  mip[k] = p;
  if (call) call[k] = c;
}

// ecdf(x)(samples) with x already sorted on the device
__global__ void k_ecdf_eval(const double* __restrict__ sorted_x, uint64_t n,
                            const double* __restrict__ samples, uint64_t m, double* __restrict__ out) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const double v = samples[k];
  // np.searchsorted places NaN after everything (NaN sorts last)
  const uint64_t idx = isnan(v) ? n : lower_bound(sorted_x, n, v);
  out[k] = lg_ecdf_y(idx, n);
}

}  // namespace lgmi
