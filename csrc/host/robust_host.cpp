// Host twins of the PS-side robust aggregators (used by the CPU / Gloo path and as a second oracle for the
// CUDA kernels).  N5 replacement: hdmedians.geomedian (Cython Weiszfeld; reference use at
// src/master/baseline_master.py:271-276).  Also exact-equality majority vote (rep_master.py:154-168) and Krum
// (baseline_master.py:278-296) over row-major float32 slabs.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

extern "C" {

// X: [P][d] float32 rows with stride `stride`.  out: [d].  Returns iterations used.
int drc_host_geomedian(const float* X, int P, long long d, long long stride, double eps, int max_iter, float* out) {
  std::vector<double> m(d), mn(d), w(P);
  for (long long k = 0; k < d; ++k) {
    double s = 0.0;
    for (int i = 0; i < P; ++i) s += X[i * stride + k];
    m[k] = s / P;
  }
  int it = 0;
  for (; it < max_iter; ++it) {
    double wsum = 0.0; bool any = false;
    for (int i = 0; i < P; ++i) {
      double s = 0.0;
      for (long long k = 0; k < d; ++k) { double t = X[i * stride + k] - m[k]; s += t * t; }
      double dist = sqrt(s);
      w[i] = dist > 1e-300 ? 1.0 / dist : 0.0;
      any |= dist > 1e-300;
      wsum += w[i];
    }
    if (!any) break;
    double move = 0.0, norm = 0.0;
    for (long long k = 0; k < d; ++k) {
      double s = 0.0;
      for (int i = 0; i < P; ++i) s += w[i] * X[i * stride + k];
      mn[k] = s / wsum;
      double t = mn[k] - m[k];
      move += t * t; norm += m[k] * m[k];
    }
    m.swap(mn);
    if (sqrt(move) <= eps * std::max(1.0, sqrt(norm))) { ++it; break; }
  }
  for (long long k = 0; k < d; ++k) out[k] = (float)m[k];
  return it;
}

// Boyer-Moore vote with whole-row equality (float compare: NaN != NaN, +0 == -0).  rows: indices into X.
int drc_host_vote(const float* X, long long d, long long stride, const int* rows, int r) {
  int cand = 0, count = 0;
  for (int k = 0; k < r; ++k) {
    if (count == 0) { cand = k; count = 1; continue; }
    const float* a = X + (long long)rows[k] * stride;
    const float* b = X + (long long)rows[cand] * stride;
    bool eq = true;
    for (long long i = 0; i < d; ++i) if (a[i] != b[i]) { eq = false; break; }
    count += eq ? 1 : -1;
  }
  return cand;
}

int drc_host_krum(const float* X, int P, long long d, long long stride, int s) {
  std::vector<double> d2((size_t)P * P, 0.0);
  for (int i = 0; i < P; ++i)
    for (int j = i + 1; j < P; ++j) {
      double acc = 0.0;
      for (long long k = 0; k < d; ++k) { double t = (double)X[i * stride + k] - X[j * stride + k]; acc += t * t; }
      d2[(size_t)i * P + j] = d2[(size_t)j * P + i] = acc;
    }
  int keep = std::max(P - s - 2, 0), best_i = 0; double best = 0.0;
  for (int i = 0; i < P; ++i) {
    std::vector<double> nb;
    for (int j = 0; j < P; ++j) if (j != i) nb.push_back(d2[(size_t)i * P + j]);
    std::sort(nb.begin(), nb.end());
    double sc = 0.0;
    for (int x = 0; x < keep && x < (int)nb.size(); ++x) sc += nb[x];
    if (i == 0 || sc < best) { best = sc; best_i = i; }
  }
  return best_i;
}

}  // extern "C"
