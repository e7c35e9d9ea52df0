// Host build of the pure device functions of l-giremi_b200/csrc/lgmi_fast.cuh
// (carry-save popcount, branch-free 2x2 / 3x3 MI epilogues) for the CPU unit
// tests.  The CUDA rounding intrinsics are shimmed with the IEEE operations
// they denote (compiled with -ffp-contract=off).  Test infrastructure only.
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline double __longlong_as_double(long long v) { double d; __builtin_memcpy(&d, &v, 8); return d; }
static inline double __hiloint2double(int hi, int lo) {
  const unsigned long long v = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
  double d; __builtin_memcpy(&d, &v, 8); return d;
}

static inline double __drcp_rn(double a) { return 1.0 / a; }   // IEEE division is correctly rounded
static inline double2 __ldg(const double2* p) { return *p; }

#include "../../l-giremi_b200/csrc/lgmi_fast.cuh"

using namespace lgmi;

// the kernel's table covers k <= 256 (kFastMaxR); `n` is ignored beyond that
static FastTab make_tab(const double* lntab, uint32_t n) {
  FastTab t;
  for (uint32_t k = 0; k <= (uint32_t)kFastMaxR; ++k) {
    t.ln[k].hi = k < n ? lntab[2 * k] : 0.0;
    t.ln[k].lo = k < n ? lntab[2 * k + 1] : 0.0;
    t.inv[k] = k ? 1.0 / (double)k : 0.0;
  }
  return t;
}

extern "C" {
double f_mi_2x2(uint32_t mm, uint32_t mM, uint32_t Mm, uint32_t MM, const double* lntab, uint32_t n) {
  auto t = make_tab(lntab, n);
  return mi_2x2(t, mm, mM, Mm, MM);
}
double f_mi_3x3(const uint32_t* T, const double* lntab, uint32_t n) {
  auto t = make_tab(lntab, n);
  return mi_3x3(t, T);
}
// batch versions (one table build)
void f_mi_2x2_many(const uint32_t* cells, int64_t m, const double* lntab, uint32_t n, double* out) {
  auto t = make_tab(lntab, n);
  for (int64_t k = 0; k < m; ++k) out[k] = mi_2x2(t, cells[4 * k], cells[4 * k + 1], cells[4 * k + 2], cells[4 * k + 3]);
}
void f_mi_3x3_many(const uint32_t* T, int64_t m, const double* lntab, uint32_t n, double* out) {
  auto t = make_tab(lntab, n);
  for (int64_t k = 0; k < m; ++k) out[k] = mi_3x3(t, T + 9 * k);
}
// the same epilogues over the context's full ln table and the computed reciprocal (k_tile_finish: counts of any size)
void g_mi_2x2_many(const uint32_t* cells, int64_t m, const double* lntab, double* out) {
  const GlobalTab t{reinterpret_cast<const lg_dd*>(lntab)};
  for (int64_t k = 0; k < m; ++k) out[k] = mi_2x2(t, cells[4 * k], cells[4 * k + 1], cells[4 * k + 2], cells[4 * k + 3]);
}
void g_mi_3x3_many(const uint32_t* T, int64_t m, const double* lntab, double* out) {
  const GlobalTab t{reinterpret_cast<const lg_dd*>(lntab)};
  for (int64_t k = 0; k < m; ++k) out[k] = mi_3x3(t, T + 9 * k);
}
// the same for `samples` random (n, N), N < 2^31 (deep units: counts far beyond the exhaustive range)
int64_t f_markstein_large_mismatches(int64_t samples, uint64_t seed) {
  int64_t bad = 0;
  uint64_t x = seed * 0x9e3779b97f4a7c15ull + 1;
  for (int64_t k = 0; k < samples; ++k) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    const uint32_t N = (uint32_t)(x >> 33) | 1u;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    const uint32_t n = (uint32_t)((x >> 11) % ((uint64_t)N + 1));
    const double dN = (double)N, y = 1.0 / dN, dn = (double)n;
    const double q0 = dn * y;
    const double r = std::fma(-q0, dN, dn);
    const double q = std::fma(r, y, q0);
    bad += (q != dn / dN);
  }
  return bad;
}
// number of (n, N) with 0 <= n <= N <= n_max where the Markstein quotient differs from n / N
int64_t f_markstein_mismatches(uint32_t n_max) {
  int64_t bad = 0;
  for (uint32_t N = 1; N <= n_max; ++N) {
    const double dN = (double)N, y = 1.0 / dN;
    for (uint32_t n = 0; n <= N; ++n) {
      const double dn = (double)n;
      const double q0 = dn * y;
      const double r = std::fma(-q0, dN, dn);
      const double q = std::fma(r, y, q0);
      bad += (q != dn / dN);
    }
  }
  return bad;
}
uint32_t f_and_popc(int nw, const uint32_t* x, const uint32_t* y) {
  switch (nw) {
    case 2: return and_popc<2>(x, y);
    case 4: return and_popc<4>(x, y);
    case 7: return and_popc<7>(x, y);
    default: return and_popc<8>(x, y);
  }
}
double f_u32_to_double(uint32_t n) { return u32_to_double(n); }
// packed counts of a pair, or ~0 when |Pi&Pj| + n_other < min_common (the early exit)
uint64_t f_pair_counts(int nw, const uint32_t* ri, const uint32_t* rj, uint32_t n_other, int min_common) {
  unsigned long long v = 0;
  bool ok;
  switch (nw) {
    case 2: ok = pair_counts<2>(ri, rj, n_other, min_common, v); break;
    case 4: ok = pair_counts<4>(ri, rj, n_other, min_common, v); break;
    case 7: ok = pair_counts<7>(ri, rj, n_other, min_common, v); break;
    default: ok = pair_counts<8>(ri, rj, n_other, min_common, v); break;
  }
  return ok ? (uint64_t)v : ~(uint64_t)0;
}
}
