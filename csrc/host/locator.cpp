// Host twin of the cyclic-code locator (N1 replacement).  Same core as the device kernel
// (csrc/common/locator_core.h); exposed with a C ABI for ctypes.
//
// Reference counterpart: `c_coding.solve_poly_a(n, s, R)` (src/c_coding.cpp:15-84, pybind11 + Eigen) plus the
// numpy/scipy tail of `CyclicMaster._decoding` (src/master/cyclic_master.py:159-170).
#include "../common/locator_core.h"

#include <stdint.h>

extern "C" {

// E: [T][n][2] doubles (re, im).  v_out: [T][n][2] doubles.  healthy_out: [T] bitmasks.  flagged_out: [T].
int drc_host_locate(const double* E, int T, int n, int s, double rel_tol, double* v_out, uint32_t* healthy_out,
                    int* flagged_out) {
  if (n > DRC_LOC_MAX_N || s > DRC_LOC_MAX_S || n < 2 * s + 1) return -1;
  for (int t = 0; t < T; ++t) {
    cplx e[DRC_LOC_MAX_N], v[DRC_LOC_MAX_N];
    for (int i = 0; i < n; ++i) e[i] = c_make(E[(t * (long long)n + i) * 2], E[(t * (long long)n + i) * 2 + 1]);
    unsigned int mask = 0;
    int fl = locate_and_recombine(e, n, s, rel_tol, v, &mask);
    for (int i = 0; i < n; ++i) {
      v_out[(t * (long long)n + i) * 2] = v[i].re;
      v_out[(t * (long long)n + i) * 2 + 1] = v[i].im;
    }
    if (healthy_out) healthy_out[t] = mask;
    if (flagged_out) flagged_out[t] = fl;
  }
  return 0;
}

// Drop-in for the reference's `solve_poly_a`: given the projected column E (n complex), return alpha (s complex).
int drc_host_solve_poly_a(const double* E, int n, int s, double* alpha_out) {
  if (n > DRC_LOC_MAX_N || s > DRC_LOC_MAX_S || s < 1 || n < 2 * s + 1) return -1;
  const double PI2 = 6.283185307179586476925286766559;
  const int k = n - 2 * s;
  const double inv_sqrt_n = 1.0 / sqrt((double)n);
  cplx synd[2 * DRC_LOC_MAX_S];
  for (int j = 0; j < 2 * s; ++j) {
    cplx acc = c_make(0.0, 0.0);
    for (int i = 0; i < n; ++i) {
      long long e = ((long long)i * (k + j)) % n;
      acc = c_add(acc, c_mul(c_polar(PI2 * (double)e / n), c_make(E[2 * i], E[2 * i + 1])));
    }
    synd[j] = c_make(acc.re * inv_sqrt_n, acc.im * inv_sqrt_n);
  }
  cplx A[DRC_LOC_MAX_S * DRC_LOC_MAX_S], b[DRC_LOC_MAX_S], alpha[DRC_LOC_MAX_S];
  for (int i = 0; i < s; ++i) {
    for (int j = 0; j < s; ++j) A[i * s + j] = synd[s - i - 1 + j];
    b[i] = synd[2 * s - i - 1];
    alpha[i] = c_make(0.0, 0.0);
  }
  c_solve_pivoted(A, b, alpha, s, 1e-10);
  for (int i = 0; i < s; ++i) { alpha_out[2 * i] = alpha[i].re; alpha_out[2 * i + 1] = alpha[i].im; }
  return 0;
}

}  // extern "C"
