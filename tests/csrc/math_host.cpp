// Host build of l-giremi_b200/csrc/lgmi_math.cuh for the CPU unit tests: the
// same source the kernels compile, with LG_* resolving to plain IEEE ops
// (-ffp-contract=off).  Test infrastructure; not part of the product library.
#include <cstdint>
#include <cstring>
#include "../../l-giremi_b200/csrc/lgmi_math.cuh"

extern "C" void lgmi_build_lntab(double* hi_lo_pairs, uint64_t k_begin, uint64_t k_end);

struct LnHost {
  const lg_dd* tab;
  lg_dd operator()(uint32_t k) const { return tab[k]; }
};

extern "C" {
double t_mi_from_table(const uint32_t* T, const double* lntab) {
  return lg_mi_from_table(T, LnHost{reinterpret_cast<const lg_dd*>(lntab)});
}
double t_mi_from_2x2(uint32_t mm, uint32_t mM, uint32_t Mm, uint32_t MM, const double* lntab) {
  return lg_mi_from_2x2(mm, mM, Mm, MM, LnHost{reinterpret_cast<const lg_dd*>(lntab)});
}
double t_ln_product(uint32_t a, uint32_t b, const double* lntab) {
  const lg_dd* t = reinterpret_cast<const lg_dd*>(lntab);
  return lg_ln_product(t[a], t[b]);
}
double t_neumaier_mean(const double* v, int64_t n) {
  lg_neumaier acc;
  lg_neumaier_init(acc);
  for (int64_t k = 0; k < n; ++k) lg_neumaier_add(acc, v[k]);
  return lg_neumaier_mean(acc);
}
void t_pair_ij(uint32_t p, uint32_t S, uint32_t* i, uint32_t* j) { lg_pair_ij(p, S, *i, *j); }
void t_pair_advance(uint32_t* i, uint32_t* j, uint32_t S, uint32_t step) { lg_pair_advance(*i, *j, S, step); }
uint64_t t_row_off(uint32_t i, uint32_t S) { return lg_row_off(i, S); }
double t_ecdf_y(uint64_t idx, uint64_t n) { return lg_ecdf_y(idx, n); }
void t_build_lntab(double* out, uint64_t k0, uint64_t k1) { lgmi_build_lntab(out, k0, k1); }
}
