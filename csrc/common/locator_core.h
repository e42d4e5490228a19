// Error-locator core of the cyclic (Fourier) code, shared by the host library (csrc/host/locator.cpp) and
// the device decode kernel (csrc/cuda/fourier.cu).  All arithmetic is complex fp64.
//
// Replaces the reference's native piece `c_coding.solve_poly_a` (src/c_coding.cpp:15-84: rebuild the n x n DFT,
// syndrome = C2^H * R, s x s Hankel system, Eigen JacobiSVD least squares) *and* the Python that follows it
// (src/master/cyclic_master.py:159-170: evaluate the locator at the n-th roots of unity, keep n-2s healthy
// rows, solve C1[h]^T v = e1 with scipy lsq_linear).  Differences by design: the DFT is never materialised
// (twiddles are generated on the fly), the rank-deficient Hankel case (fewer than s actual liars) is handled
// by rank-revealing elimination with complete pivoting instead of an SVD, and the healthy-row threshold is
// relative because the payload is complex64, not complex128.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define DRC_HD __host__ __device__ __forceinline__
#else
#define DRC_HD inline
#endif

#define DRC_LOC_MAX_N 32
#define DRC_LOC_MAX_S 8

struct cplx {
  double re, im;
};
DRC_HD cplx c_make(double r, double i) { cplx z; z.re = r; z.im = i; return z; }
DRC_HD cplx c_add(cplx a, cplx b) { return c_make(a.re + b.re, a.im + b.im); }
DRC_HD cplx c_sub(cplx a, cplx b) { return c_make(a.re - b.re, a.im - b.im); }
DRC_HD cplx c_mul(cplx a, cplx b) { return c_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
DRC_HD double c_abs2(cplx a) { return a.re * a.re + a.im * a.im; }
DRC_HD cplx c_div(cplx a, cplx b) {
  double d = c_abs2(b);
  return c_make((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}
DRC_HD cplx c_polar(double ang) { return c_make(cos(ang), sin(ang)); }

// Solve the m x m complex system A x = b (row-major A, overwritten) by Gaussian elimination with complete
// pivoting.  Pivots below rel_tol * |largest entry of A| are treated as zero (rank deficiency): the
// corresponding unknowns are set to 0, which yields a valid particular solution of a consistent system.
// Returns the detected rank.
DRC_HD int c_solve_pivoted(cplx* A, cplx* b, cplx* x, int m, double rel_tol) {
  int colperm[DRC_LOC_MAX_N];
  for (int i = 0; i < m; ++i) colperm[i] = i;
  double amax = 0.0;
  for (int i = 0; i < m * m; ++i) { double v = c_abs2(A[i]); if (v > amax) amax = v; }
  const double thresh2 = amax * rel_tol * rel_tol;
  int rank = 0;
  for (int k = 0; k < m; ++k) {
    int pr = -1, pc = -1; double best = 0.0;
    for (int i = k; i < m; ++i)
      for (int j = k; j < m; ++j) { double v = c_abs2(A[i * m + j]); if (v > best) { best = v; pr = i; pc = j; } }
    if (pr < 0 || best <= thresh2 || best == 0.0) break;
    if (pr != k) { for (int j = 0; j < m; ++j) { cplx t = A[k * m + j]; A[k * m + j] = A[pr * m + j]; A[pr * m + j] = t; }
                   cplx t = b[k]; b[k] = b[pr]; b[pr] = t; }
    if (pc != k) { for (int i = 0; i < m; ++i) { cplx t = A[i * m + k]; A[i * m + k] = A[i * m + pc]; A[i * m + pc] = t; }
                   int t = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = t; }
    for (int i = k + 1; i < m; ++i) {
      cplx f = c_div(A[i * m + k], A[k * m + k]);
      for (int j = k; j < m; ++j) A[i * m + j] = c_sub(A[i * m + j], c_mul(f, A[k * m + j]));
      b[i] = c_sub(b[i], c_mul(f, b[k]));
    }
    ++rank;
  }
  cplx y[DRC_LOC_MAX_N];
  for (int i = 0; i < m; ++i) y[i] = c_make(0.0, 0.0);
  for (int k = rank - 1; k >= 0; --k) {
    cplx acc = b[k];
    for (int j = k + 1; j < rank; ++j) acc = c_sub(acc, c_mul(A[k * m + j], y[j]));
    y[k] = c_div(acc, A[k * m + k]);
  }
  for (int i = 0; i < m; ++i) x[colperm[i]] = y[i];
  return rank;
}

// Full locate step for one tensor.
//   E[n]      : projected codeword rows  E_i = sum_k R[i,k] f[k]   (complex)
//   v[n]      : out, recombination vector (zero outside the chosen healthy rows)
//   healthy   : out, bitmask of the rows used
// Returns the number of rows flagged Byzantine (locator ~ 0).
DRC_HD int locate_and_recombine(const cplx* E, int n, int s, double rel_tol, cplx* v, unsigned int* healthy_mask) {
  const double PI2 = 6.283185307179586476925286766559;
  const int k = n - 2 * s;
  const double inv_sqrt_n = 1.0 / sqrt((double)n);
  double pmag[DRC_LOC_MAX_N];
  int flagged = 0;
  if (s == 0) {
    for (int t = 0; t < n; ++t) pmag[t] = 1.0;
  } else {
    // syndrome_j = sum_i conj(C[i, k + j]) E_i,  C[i, q] = exp(-2 pi i * i q / n) / sqrt(n)
    cplx synd[2 * DRC_LOC_MAX_S];
    double emax = 0.0, smax = 0.0;
    for (int i = 0; i < n; ++i) { double a = c_abs2(E[i]); if (a > emax) emax = a; }
    for (int j = 0; j < 2 * s; ++j) {
      cplx acc = c_make(0.0, 0.0);
      for (int i = 0; i < n; ++i) {
        long long e = ((long long)i * (k + j)) % n;
        acc = c_add(acc, c_mul(c_polar(PI2 * (double)e / n), E[i]));
      }
      synd[j] = c_make(acc.re * inv_sqrt_n, acc.im * inv_sqrt_n);
      double a = c_abs2(synd[j]); if (a > smax) smax = a;
    }
    cplx alpha[DRC_LOC_MAX_S];
    for (int i = 0; i < s; ++i) alpha[i] = c_make(0.0, 0.0);
    // a syndrome at rounding-noise level means "no liar": p(z) = z^s, every row healthy
    if (smax > emax * 1e-12) {
      cplx A[DRC_LOC_MAX_S * DRC_LOC_MAX_S], b[DRC_LOC_MAX_S];
      for (int i = 0; i < s; ++i) {
        for (int j = 0; j < s; ++j) A[i * s + j] = synd[s - i - 1 + j];
        b[i] = synd[2 * s - i - 1];
      }
      c_solve_pivoted(A, b, alpha, s, 1e-6);
    }
    double pmax = 0.0;
    for (int t = 0; t < n; ++t) {
      // p(z_t) = z^s - sum_j alpha_j z^j at z_t = exp(+2 pi i t / n)
      cplx p = c_polar(PI2 * (double)(((long long)t * s) % n) / n);
      for (int j = 0; j < s; ++j)
        p = c_sub(p, c_mul(alpha[j], c_polar(PI2 * (double)(((long long)t * j) % n) / n)));
      pmag[t] = sqrt(c_abs2(p));
      if (pmag[t] > pmax) pmax = pmag[t];
    }
    for (int t = 0; t < n; ++t) if (!(pmag[t] > rel_tol * pmax)) ++flagged;
    if (n - flagged < k) {                      // threshold too aggressive: fall back to the k largest values
      flagged = 0;
      double cut = 0.0;                         // k-th largest magnitude
      for (int t = 0; t < n; ++t) {
        int larger = 0;
        for (int u = 0; u < n; ++u) if (pmag[u] > pmag[t] || (pmag[u] == pmag[t] && u < t)) ++larger;
        if (larger == k - 1) cut = pmag[t];
      }
      for (int t = 0; t < n; ++t) if (pmag[t] < cut) { pmag[t] = 0.0; ++flagged; }
      pmax = 1.0; rel_tol = 0.0;
      for (int t = 0; t < n; ++t) if (pmag[t] > 0.0) pmag[t] = 1.0;
    } else {
      for (int t = 0; t < n; ++t) pmag[t] = (pmag[t] > rel_tol * pmax) ? 1.0 : 0.0;
    }
  }
  // first k healthy rows in index order (reference: cyclic_master.py:164,169)
  int h[DRC_LOC_MAX_N]; int nh = 0;
  for (int t = 0; t < n && nh < k; ++t) if (pmag[t] > 0.0) h[nh++] = t;
  // solve C1[h]^T v_h = e1 :  M[a][b] = C1[h_b, a] = exp(-2 pi i h_b a / n) / sqrt(n)
  cplx M[DRC_LOC_MAX_N * DRC_LOC_MAX_N], rhs[DRC_LOC_MAX_N], vh[DRC_LOC_MAX_N];
  for (int a = 0; a < k; ++a) {
    for (int b = 0; b < k; ++b) {
      long long e = ((long long)h[b] * a) % n;
      cplx w = c_polar(-PI2 * (double)e / n);
      M[a * k + b] = c_make(w.re * inv_sqrt_n, w.im * inv_sqrt_n);
    }
    rhs[a] = c_make(a == 0 ? 1.0 : 0.0, 0.0);
  }
  c_solve_pivoted(M, rhs, vh, k, 1e-13);
  unsigned int mask = 0u;
  for (int t = 0; t < n; ++t) v[t] = c_make(0.0, 0.0);
  for (int b = 0; b < k; ++b) { v[h[b]] = vh[b]; mask |= 1u << h[b]; }
  if (healthy_mask) *healthy_mask = mask;
  return flagged;
}
