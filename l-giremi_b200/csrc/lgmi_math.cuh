// lgmi_math.cuh -- the fp64 arithmetic of the MI step, shared by the CUDA
// kernels and (compiled for the host) by the CPU unit tests of that arithmetic.
//
// Every function here is a deterministic function of small integers, written
// with explicitly rounded operations (no FMA contraction) so that device and
// host produce the same bits.  The term structure follows scikit-learn 1.9.0
// mutual_info_score (sklearn/metrics/cluster/_supervised.py:920-935), which is
// what the reference calls at /root/reference/src/giremi/mutual_information.py:41.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define LG_HD __host__ __device__ __forceinline__
#else
#define LG_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define LG_ADD(a, b) __dadd_rn((a), (b))
#define LG_SUB(a, b) __dsub_rn((a), (b))
#define LG_MUL(a, b) __dmul_rn((a), (b))
#define LG_DIV(a, b) __ddiv_rn((a), (b))
#else
// host build: compiled with -ffp-contract=off, so these stay separate roundings
#define LG_ADD(a, b) ((a) + (b))
#define LG_SUB(a, b) ((a) - (b))
#define LG_MUL(a, b) ((a) * (b))
#define LG_DIV(a, b) ((a) / (b))
#endif

// ln(k) as an unevaluated sum hi+lo with hi = RN(ln k) (correctly rounded) and
// |lo| <= ulp(hi)/2.  The table is built on the host in binary128
// (libquadmath logq) and indexed by k = 0..k_max (entry 0 is unused).
struct lg_dd {
  double hi, lo;
};

#define LG_EPS 2.220446049250313e-16 /* np.finfo(float64).eps, _supervised.py:934 */

// RN(ln(a*b)) from the double-double logs of a and b.  np.log(float(a*b)) at
// _supervised.py:929 is the correctly rounded log for all but ~1e-5 of
// arguments (measured); this returns RN(ln a + ln b) exactly unless
// the true value lies within 2^-100 of a rounding boundary.
LG_HD double lg_ln_product(lg_dd la, lg_dd lb) {
  double s = LG_ADD(la.hi, lb.hi);
  double bb = LG_SUB(s, la.hi);
  double e = LG_ADD(LG_SUB(la.hi, LG_SUB(s, bb)), LG_SUB(lb.hi, bb));  // two_sum error
  e = LG_ADD(e, LG_ADD(la.lo, lb.lo));
  return LG_ADD(s, e);
}

// One cell's contribution (_supervised.py:923-934):
//   q = n/N ; t = q*(ln n - ln N) + q*((-ln(a*b) + ln N) + ln N) ; |t|<eps -> 0
LG_HD double lg_mi_term(double n, double total, double ln_n, double ln_total,
                        double ln_outer_ab) {
  double q = LG_DIV(n, total);
  double lo = LG_ADD(LG_ADD(-ln_outer_ab, ln_total), ln_total);
  double t = LG_ADD(LG_MUL(q, LG_SUB(ln_n, ln_total)), LG_MUL(q, lo));
  return (fabs(t) < LG_EPS) ? 0.0 : t;
}

// MI of a 3x3 table T[a*3+b], a = label of site 1, b = label of site 2,
// labels 0 other / 1 minor / 2 major (np.unique order == sp.find row-major
// order over the classes present).  `ln` is any callable k -> lg_dd.
template <class LnTab>
LG_HD double lg_mi_from_table(const uint32_t T[9], LnTab ln) {
  uint32_t r[3], c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) r[a] = T[a * 3] + T[a * 3 + 1] + T[a * 3 + 2];
#pragma unroll
  for (int b = 0; b < 3; ++b) c[b] = T[b] + T[3 + b] + T[6 + b];
  const uint32_t total = r[0] + r[1] + r[2];
  const int nrow = (r[0] != 0) + (r[1] != 0) + (r[2] != 0);
  const int ncol = (c[0] != 0) + (c[1] != 0) + (c[2] != 0);
  if (nrow <= 1 || ncol <= 1) return 0.0;  // :920 single-class shortcut
  const double dtotal = (double)total;
  const double ln_total = ln(total).hi;
  lg_dd lr[3], lc[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) lr[a] = ln(r[a] ? r[a] : 1u);
#pragma unroll
  for (int b = 0; b < 3; ++b) lc[b] = ln(c[b] ? c[b] : 1u);
  double t[9];
  int nnz = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const uint32_t n = T[a * 3 + b];
      double v = 0.0;
      if (n) {
        v = lg_mi_term((double)n, dtotal, ln(n).hi, ln_total, lg_ln_product(lr[a], lc[b]));
        ++nnz;
      }
      t[a * 3 + b] = v;
    }
  }
  double s;
  if (nnz < 8) {
    // ndarray.sum() below 8 elements is a plain left-to-right loop; adding the
    // +0.0 of an absent cell does not change any partial sum.
    s = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) s = LG_ADD(s, t[k]);
  } else {
    // 8 or 9 terms: numpy pairwise_sum uses 8 accumulators seeded with the
    // first 8 terms, a balanced combine, then the tail.
    double u[9];
    int m = 0;
    for (int k = 0; k < 9; ++k)
      if (T[k]) u[m++] = t[k];
    s = LG_ADD(LG_ADD(LG_ADD(u[0], u[1]), LG_ADD(u[2], u[3])),
               LG_ADD(LG_ADD(u[4], u[5]), LG_ADD(u[6], u[7])));
    if (m == 9) s = LG_ADD(s, u[8]);
  }
  return (s > 0.0) ? s : 0.0;  // :935 clip(lower=0); NaN cannot occur
}

// 2x2 specialisation (both sites bi-allelic over the common reads): the four
// cells in sp.find order are mm, mM, Mm, MM (minor label 1 sorts before
// major label 2).  Same bits as lg_mi_from_table on the embedded table.
template <class LnTab>
LG_HD double lg_mi_from_2x2(uint32_t n_mm, uint32_t n_mM, uint32_t n_Mm, uint32_t n_MM,
                            LnTab ln) {
  const uint32_t r_m = n_mm + n_mM, r_M = n_Mm + n_MM;
  const uint32_t c_m = n_mm + n_Mm, c_M = n_mM + n_MM;
  if (r_m == 0 || r_M == 0 || c_m == 0 || c_M == 0) return 0.0;
  const uint32_t total = r_m + r_M;
  const double dtotal = (double)total;
  const double ln_total = ln(total).hi;
  const lg_dd lrm = ln(r_m), lrM = ln(r_M), lcm = ln(c_m), lcM = ln(c_M);
  double s = 0.0;
  if (n_mm) s = LG_ADD(s, lg_mi_term((double)n_mm, dtotal, ln(n_mm).hi, ln_total, lg_ln_product(lrm, lcm)));
  if (n_mM) s = LG_ADD(s, lg_mi_term((double)n_mM, dtotal, ln(n_mM).hi, ln_total, lg_ln_product(lrm, lcM)));
  if (n_Mm) s = LG_ADD(s, lg_mi_term((double)n_Mm, dtotal, ln(n_Mm).hi, ln_total, lg_ln_product(lrM, lcm)));
  if (n_MM) s = LG_ADD(s, lg_mi_term((double)n_MM, dtotal, ln(n_MM).hi, ln_total, lg_ln_product(lrM, lcM)));
  return (s > 0.0) ? s : 0.0;
}

// CPython >= 3.12 float sum (Neumaier compensation), one step.  The reference
// averages with sum(...)/len(...) at mutual_information.py:56-58.
struct lg_neumaier {
  double s, c;
  int n;
};
LG_HD void lg_neumaier_init(lg_neumaier& a) {
  a.s = 0.0;
  a.c = 0.0;
  a.n = 0;
}
LG_HD void lg_neumaier_add(lg_neumaier& a, double x) {
  if (a.n == 0) {
    a.s = LG_ADD(0.0, x);  // int 0 + first float
  } else {
    double t = LG_ADD(a.s, x);
    if (fabs(a.s) >= fabs(x))
      a.c = LG_ADD(a.c, LG_ADD(LG_SUB(a.s, t), x));
    else
      a.c = LG_ADD(a.c, LG_ADD(LG_SUB(x, t), a.s));
    a.s = t;
  }
  ++a.n;
}
LG_HD double lg_neumaier_mean(const lg_neumaier& a) {
  if (a.n == 0) return nan("");
  double s = a.s;
  if (a.c != 0.0 && isfinite(a.c)) s = LG_ADD(s, a.c);
  return LG_DIV(s, (double)a.n);
}

// Triangular pair index: pairs (i<j) of S sites in lexicographic order.
LG_HD uint64_t lg_row_off(uint32_t i, uint32_t S) {
  return ((uint64_t)i * (2ull * S - i - 1ull)) >> 1;
}
// (i, j) of pair p + step from the (i, j) of pair p, without the square root
LG_HD void lg_pair_advance(uint32_t& i, uint32_t& j, uint32_t S, uint32_t step) {
  j += step;
  while (j >= S) {  // past the end of row i by j - S pairs: row i + 1 starts at column i + 2
    ++i;
    j = j - S + i + 1u;
  }
}
LG_HD void lg_pair_ij(uint32_t p, uint32_t S, uint32_t& i, uint32_t& j) {
  const double b = 2.0 * (double)S - 1.0;
  double disc = b * b - 8.0 * (double)p;
  if (disc < 0.0) disc = 0.0;
  int64_t ii = (int64_t)((b - sqrt(disc)) * 0.5);
  if (ii < 0) ii = 0;
  if (ii > (int64_t)S - 2) ii = (int64_t)S - 2;
  while (ii + 1 <= (int64_t)S - 2 && lg_row_off((uint32_t)ii + 1u, S) <= p) ++ii;
  while (ii > 0 && lg_row_off((uint32_t)ii, S) > p) --ii;
  i = (uint32_t)ii;
  j = (uint32_t)(p - lg_row_off(i, S)) + i + 1u;
}

// numpy.linspace(1/n, 1, n)[k] as stat.py:19 builds the ECDF ordinate;
// y[0] = 0, y[idx] = linspace[idx-1].
LG_HD double lg_ecdf_y(uint64_t idx, uint64_t n) {
  if (idx == 0) return 0.0;
  const uint64_t k = idx - 1;
  if (n == 1 || k == n - 1) return 1.0;  // endpoint is assigned exactly
  const double start = LG_DIV(1.0, (double)n);
  const double delta = LG_SUB(1.0, start);
  const double div = (double)(n - 1);
  const double step = LG_DIV(delta, div);
  if (step == 0.0) return LG_ADD(LG_MUL(LG_DIV((double)k, div), delta), start);
  return LG_ADD(LG_MUL((double)k, step), start);
}
