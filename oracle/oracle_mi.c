/* oracle_mi.c -- CPU oracle of the L-GIREMI MI step in plain C.
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg; never by the product (l-giremi_b200/).
 *
 * A from-scratch restatement (no reference source is copied) of
 *   /root/reference/src/giremi/mutual_information.py:6-45   pair MI
 *   /root/reference/src/giremi/mutual_information.py:48-60  per-site mean
 *   /root/reference/src/giremi/mismatch.py:393-396          het filter
 *   /root/reference/src/giremi/stat.py:7-29 + script/giremi.py:415-429, 97-114
 * and of the third-party arithmetic the reference calls but does not vendor:
 *   scikit-learn 1.9.0 mutual_info_score,
 *   sklearn/metrics/cluster/_supervised.py:822-935 (contingency :96-182).
 *
 * Parity is PINNED: tests/test_oracle.py checks this file against the golden
 * vectors in tests/golden/ that tests/golden/make_golden.py produced by
 * executing the real reference and the installed scikit-learn.
 *
 * Input is the reference's `mismatches` dict in integer-coded form, keeping
 * everything the algorithm's result depends on (orders included):
 *   site s owns  nt entries  ent_allele[e], ent_read[e]  for e in
 *   [ent_off[s], ent_off[s+1])  in ('nt' key order, list order), and depth
 *   entries dep_allele[d], dep_count[d] for d in [dep_off[s], dep_off[s+1])
 *   in `depth` dict order.  Reads are ids 0..n_reads-1 (one id per name).
 *   Sites are already sorted by position.
 * Logs are libm `log` (what math.log / numpy's scalar path resolve to for
 * these arguments; see DESIGN.md on correctly-rounded agreement).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_EPS 2.220446049250313e-16

/* ---- sklearn mutual_info_score on a 3x3 table, index = label ------------- */
static double sum_like_numpy(const double* v, int n) {
  if (n < 8) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc += v[k];
    return acc;
  }
  double r[8];
  for (int k = 0; k < 8; ++k) r[k] = v[k];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int k = 0; k < 8; ++k) r[k] += v[i + k];
  double acc = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) acc += v[i];
  return acc;
}

double oracle_mi_from_table(const int64_t t[9]) {
  int64_t row[3], col[3], total = 0;
  int nrow = 0, ncol = 0;
  for (int a = 0; a < 3; ++a) {
    row[a] = t[3 * a] + t[3 * a + 1] + t[3 * a + 2];
    col[a] = t[a] + t[3 + a] + t[6 + a];
    total += row[a];
  }
  for (int a = 0; a < 3; ++a) {
    nrow += row[a] > 0;
    ncol += col[a] > 0;
  }
  if (nrow <= 1 || ncol <= 1) return 0.0;
  const double ln_total = log((double)total);
  double terms[9];
  int n = 0;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      const int64_t c = t[3 * a + b];
      if (c == 0) continue;
      const double q = (double)c / (double)total;
      const double ln_c = log((double)c);
      const double ln_outer = -log((double)(row[a] * col[b])) + ln_total + ln_total;
      double term = q * (ln_c - ln_total) + q * ln_outer;
      if (fabs(term) < ORACLE_EPS) term = 0.0;
      terms[n++] = term;
    }
  const double s = sum_like_numpy(terms, n);
  return s > 0.0 ? s : 0.0;
}

/* ---- CPython >= 3.12 float sum ------------------------------------------- */
double oracle_python_sum(const double* v, int64_t n) {
  if (n == 0) return 0.0;
  double s = 0.0 + v[0], c = 0.0;
  for (int64_t k = 1; k < n; ++k) {
    const double x = v[k], t = s + x;
    if (fabs(s) >= fabs(x)) c += (s - t) + x; else c += (x - t) + s;
    s = t;
  }
  if (c != 0.0 && isfinite(c)) s += c;
  return s;
}

/* ---- one unit ------------------------------------------------------------ */
/* labels[s*n_reads + r] = 2 major / 1 minor / 0 other / -1 not covered.
 * bad[s] = 1 when `depth` has fewer than two alleles (IndexError in the
 * reference once a surviving pair touches the site). */
void oracle_site_labels(int32_t n_sites, int32_t n_reads, const int64_t* ent_off,
                        const int32_t* ent_allele, const int32_t* ent_read, const int64_t* dep_off,
                        const int32_t* dep_allele, const int64_t* dep_count, int8_t* labels,
                        uint8_t* bad) {
  int32_t* allele_of = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_reads > 0 ? n_reads : 1));
  for (int32_t s = 0; s < n_sites; ++s) {
    for (int32_t r = 0; r < n_reads; ++r) allele_of[r] = -1;
    for (int64_t e = ent_off[s]; e < ent_off[s + 1]; ++e) allele_of[ent_read[e]] = ent_allele[e];
    /* stable descending order by depth: pick the best, then the best of the rest */
    int64_t best = -1, second = -1;
    for (int64_t d = dep_off[s]; d < dep_off[s + 1]; ++d)
      if (best < 0 || dep_count[d] > dep_count[best]) best = d;
    for (int64_t d = dep_off[s]; d < dep_off[s + 1]; ++d)
      if (d != best && (second < 0 || dep_count[d] > dep_count[second])) second = d;
    /* stability: among equal depths the earlier dict entry ranks first; the
       strict '>' scans keep the earliest maximum, first overall and then
       among the remaining entries */
    const int32_t major = best >= 0 ? dep_allele[best] : -2;
    const int32_t minor = second >= 0 ? dep_allele[second] : -2;
    bad[s] = (uint8_t)(second < 0);
    for (int32_t r = 0; r < n_reads; ++r) {
      const int32_t a = allele_of[r];
      labels[(size_t)s * n_reads + r] = (int8_t)(a < 0 ? -1 : (a == major ? 2 : (a == minor ? 1 : 0)));
    }
  }
  free(allele_of);
}

/* Evaluates every pair of one unit.  Outputs (capacity n_sites*(n_sites-1)/2):
 *   out_i/out_j/out_mi/out_table(9 per row) for the surviving pairs in
 *   combinations order; returns their number, or -1 if a surviving pair
 *   touches a bad site (the reference raises IndexError). */
int64_t oracle_unit_pairs(int32_t n_sites, int32_t n_reads, const int8_t* labels, const uint8_t* bad,
                          int32_t min_common, int32_t* out_i, int32_t* out_j, double* out_mi,
                          int64_t* out_table) {
  int64_t n_out = 0;
  for (int32_t i = 0; i < n_sites; ++i)
    for (int32_t j = i + 1; j < n_sites; ++j) {
      const int8_t* li = labels + (size_t)i * n_reads;
      const int8_t* lj = labels + (size_t)j * n_reads;
      int64_t t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      int64_t common = 0;
      for (int32_t r = 0; r < n_reads; ++r)
        if (li[r] >= 0 && lj[r] >= 0) {
          ++common;
          ++t[3 * li[r] + lj[r]];
        }
      if (common < min_common) continue;
      if (bad[i] || bad[j]) return -1;
      out_i[n_out] = i;
      out_j[n_out] = j;
      out_mi[n_out] = oracle_mi_from_table(t);
      if (out_table) memcpy(out_table + 9 * n_out, t, sizeof t);
      ++n_out;
    }
  return n_out;
}

/* het filter + per-site mean; mean[s] = NaN when the site is in no kept pair.
 * is_het[s] != 0 marks 'het_snp'. */
void oracle_site_means(int32_t n_sites, const uint8_t* is_het, int64_t n_rows, const int32_t* row_i,
                       const int32_t* row_j, const double* row_mi, double* mean, int32_t* cnt) {
  double* buf = (double*)malloc(sizeof(double) * (size_t)(n_rows > 0 ? n_rows : 1));
  for (int32_t s = 0; s < n_sites; ++s) {
    int64_t n = 0;
    for (int64_t k = 0; k < n_rows; ++k) {
      if (!(is_het[row_i[k]] || is_het[row_j[k]])) continue;
      if (row_i[k] == s || row_j[k] == s) buf[n++] = row_mi[k];
    }
    cnt[s] = (int32_t)n;
    mean[s] = n ? oracle_python_sum(buf, n) / (double)n : NAN;
  }
  free(buf);
}

/* ---- global pass ---------------------------------------------------------- */
static int cmp_double(const void* a, const void* b) {
  const double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

static double linspace_at(uint64_t k, uint64_t n) { /* numpy.linspace(1/n, 1, n)[k] */
  if (n == 1 || k == n - 1) return 1.0;
  const double start = 1.0 / (double)n, delta = 1.0 - start, div = (double)(n - 1);
  const double step = delta / div;
  if (step == 0.0) return ((double)k / div) * delta + start;
  return (double)k * step + start;
}

/* type codes: 0 mismatch, 1 snp, 2 het_snp.  call: 1 positive, 2 negative. */
void oracle_mip_calls(int64_t n, const double* mean, const uint8_t* type, double threshold, double* mip,
                      uint8_t* call) {
  double* x = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  uint64_t nh = 0;
  for (int64_t k = 0; k < n; ++k)
    if (!isnan(mean[k]) && type[k] == 2) x[nh++] = mean[k];
  qsort(x, nh, sizeof(double), cmp_double);
  for (int64_t k = 0; k < n; ++k) {
    mip[k] = NAN;
    call[k] = 0;
    if (isnan(mean[k]) || nh == 0) continue;
    uint64_t lo = 0, hi = nh;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) / 2;
      if (x[mid] < mean[k]) lo = mid + 1; else hi = mid;
    }
    mip[k] = lo == 0 ? 0.0 : linspace_at(lo - 1, nh);
    if (mip[k] <= threshold && type[k] == 0) call[k] = 1;
    else if (mip[k] > threshold && type[k] != 0) call[k] = 2;
  }
  free(x);
}
