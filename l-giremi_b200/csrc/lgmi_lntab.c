/* lgmi_lntab.c -- host-side builder of the double-double ln(k) table.
 *
 * Compiled with gcc (not nvcc) because it uses binary128 (__float128 / logq
 * from libquadmath, linked statically).  hi = RN(ln k), lo = RN(ln k - hi):
 * 113-bit logs split into two doubles, so that ln(a*b) = ln a + ln b can be
 * rounded correctly on the device (lgmi_math.cuh: lg_ln_product).
 */
#include <quadmath.h>
#include <stddef.h>
#include <stdint.h>

void lgmi_build_lntab(double* hi_lo_pairs, uint64_t k_begin, uint64_t k_end) {
  for (uint64_t k = k_begin; k < k_end; ++k) {
    double hi = 0.0, lo = 0.0;
    if (k >= 1) {
      __float128 q = logq((__float128)k);
      hi = (double)q;
      lo = (double)(q - (__float128)hi);
    }
    hi_lo_pairs[2 * (k - k_begin)] = hi;
    hi_lo_pairs[2 * (k - k_begin) + 1] = lo;
  }
}
