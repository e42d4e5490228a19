// K10 (normalisation part): fused training-mode BatchNorm + residual add + ReLU for NHWC bf16 activations.
//
// The activations of the CIFAR models are small (2-17 MB per tensor at B=128) and usually still L2-resident when the
// normalisation runs, so these kernels are latency-bound, not bandwidth-bound: what matters is memory-level parallelism
// and a short reduction tail.  Design (second iteration, see profiles/bn_fused.md for the measurements that drove it):
//
//   * 1024-thread CTAs, at most one per SM, each thread owning 4 consecutive channels (8-byte accesses) of a strided
//     set of rows with the row loop unrolled x4  -> ~32 KB of loads in flight per SM;
//   * the grid is sized from the tensor (one CTA per ~128 KB, capped at the SM count) so that the deterministic fold of
//     the per-CTA partials by the last CTA is <= a handful of dependent L2 round trips;
//   * forward : bn_stats_kernel  (x -> mean, invstd, running stats)        bn_apply_kernel (x, res -> y, ReLU fused)
//     backward: bn_bwd_reduce_kernel (dy, y, x -> dgamma, dbeta)            bn_bwd_apply_kernel (dy, y, x -> dx, dres)
//   * every reduction has a FIXED summation order (no float atomics): replicas on different GPUs stay bit-identical,
//     which the exact-equality majority vote requires (reference: src/master/rep_master.py:162).
//
// Reference counterpart: nn.BatchNorm2d / F.relu inside src/model_ops/resnet.py:14-64 and vgg.py:46-59 (PyTorch-0.3 CPU).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int BN_THREADS = 1024;
constexpr int BN_VEC = 4;                      // channels per thread
constexpr int BN_UNROLL = 4;

struct BnFwdArgs {
  const __nv_bfloat16* x;         // [M][C]
  const __nv_bfloat16* res;       // optional residual [M][C]
  __nv_bfloat16* y;               // [M][C]
  const float* gamma;             // [C]
  const float* beta;              // [C]
  float* running_mean;            // [C] (may be null)
  float* running_var;             // [C]
  float* mean;                    // [C] out (saved for backward)
  float* invstd;                  // [C] out
  float* partial;                 // [nblk][2][C] workspace
  unsigned int* counter;          // zero on entry, reset by the last CTA
  long long M;
  int C;
  int rows_per_cta;
  float eps, momentum;
  int relu;
};

struct BnBwdArgs {
  const __nv_bfloat16* dy;        // [M][C]
  const __nv_bfloat16* y;         // forward output (ReLU mask), null when relu == 0
  const __nv_bfloat16* x;         // [M][C]
  const float* gamma;
  const float* mean;
  const float* invstd;
  __nv_bfloat16* dx;              // [M][C]
  __nv_bfloat16* dres;            // optional: gradient of the residual input
  float* dgamma;                  // [C]
  float* dbeta;                   // [C]
  float* partial;                 // [nblk][2][C]
  float* sums;                    // [2][C] : mean(dy_r), mean(dy_r * xhat)
  unsigned int* counter;
  long long M;
  int C;
  int rows_per_cta;
  int relu;
};

__device__ __forceinline__ void unpack4(const uint2& v, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
  float2 a = __bfloat1622float2(p[0]), b = __bfloat1622float2(p[1]);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ uint2 pack4(const float* f) {
  uint2 v;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
  p[0] = __floats2bfloat162_rn(f[0], f[1]);
  p[1] = __floats2bfloat162_rn(f[2], f[3]);
  return v;
}
__device__ __forceinline__ uint2 ldg8(const __nv_bfloat16* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

extern __shared__ float bn_smem[];

// CTA-level fold of per-thread accumulators acc[2][4] over the row groups into partial_blk[2][C] (fixed order).
__device__ __forceinline__ void cta_fold(const float (&acc)[2][BN_VEC], float* partial_blk, int C, int rgroups, int rg, int cv) {
  // smem layout: [rgroups][2][C]
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int i = 0; i < BN_VEC; ++i) bn_smem[(rg * 2 + a) * C + cv * BN_VEC + i] = acc[a][i];
  __syncthreads();
  // all 1024 threads participate: idx = (which, c), the row groups are split into `parts` interleaved subsets
  const int n_idx = 2 * C;
  const int parts = BN_THREADS / n_idx > 0 ? BN_THREADS / n_idx : 1;
  if (parts > 1) {
    const int idx = threadIdx.x % n_idx, part = threadIdx.x / n_idx;
    float s = 0.f;
    for (int g = part; g < rgroups; g += parts) s += bn_smem[g * n_idx + idx];
    __syncthreads();
    bn_smem[part * n_idx + idx] = s;
    __syncthreads();
    if (part == 0) {
      float t = 0.f;
      for (int p = 0; p < parts; ++p) t += bn_smem[p * n_idx + idx];
      partial_blk[idx] = t;
    }
  } else {
    for (int idx = threadIdx.x; idx < n_idx; idx += BN_THREADS) {
      float s = 0.f;
      for (int g = 0; g < rgroups; ++g) s += bn_smem[g * n_idx + idx];
      partial_blk[idx] = s;
    }
  }
}

__device__ __forceinline__ bool last_cta(unsigned int* counter) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int prev = atomicAdd(counter, 1u);
    s_last = (prev == gridDim.x - 1);
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// Grid-level fold by the last CTA: out[idx] = sum_b partial[b][idx], idx in [0, 2C), fixed order, all threads busy.
// Result is left in bn_smem[0 .. 2C).
__device__ __forceinline__ void grid_fold(const float* partial, int C) {
  const int n_idx = 2 * C;
  const unsigned int nblk = gridDim.x;
  const int parts = BN_THREADS / n_idx > 0 ? BN_THREADS / n_idx : 1;
  __syncthreads();
  if (parts > 1) {
    const int idx = threadIdx.x % n_idx, part = threadIdx.x / n_idx;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    unsigned int b = part;
    for (; b + 3 * parts < nblk; b += 4 * parts) {
      s0 += __ldcg(partial + (long long)b * n_idx + idx);
      s1 += __ldcg(partial + (long long)(b + parts) * n_idx + idx);
      s2 += __ldcg(partial + (long long)(b + 2 * parts) * n_idx + idx);
      s3 += __ldcg(partial + (long long)(b + 3 * parts) * n_idx + idx);
    }
    for (; b < nblk; b += parts) s0 += __ldcg(partial + (long long)b * n_idx + idx);
    bn_smem[n_idx + part * n_idx + idx] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (part == 0) {
      float t = 0.f;
      for (int p = 0; p < parts; ++p) t += bn_smem[n_idx + p * n_idx + idx];
      bn_smem[idx] = t;
    }
  } else {
    for (int idx = threadIdx.x; idx < n_idx; idx += BN_THREADS) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      unsigned int b = 0;
      for (; b + 4 <= nblk; b += 4) {
        s0 += __ldcg(partial + (long long)b * n_idx + idx);
        s1 += __ldcg(partial + (long long)(b + 1) * n_idx + idx);
        s2 += __ldcg(partial + (long long)(b + 2) * n_idx + idx);
        s3 += __ldcg(partial + (long long)(b + 3) * n_idx + idx);
      }
      for (; b < nblk; ++b) s0 += __ldcg(partial + (long long)b * n_idx + idx);
      bn_smem[idx] = (s0 + s1) + (s2 + s3);
    }
  }
  __syncthreads();
}

#define BN_ROW_LOOP(BODY_LOAD, BODY_USE)                                                            \
  {                                                                                                 \
    long long r = r0 + rg;                                                                          \
    for (; r + (long long)(BN_UNROLL - 1) * rgroups < r1; r += (long long)BN_UNROLL * rgroups) {    \
      _Pragma("unroll") for (int u = 0; u < BN_UNROLL; ++u) { const long long off = (r + (long long)u * rgroups) * a.C + cv * BN_VEC; BODY_LOAD }  \
      _Pragma("unroll") for (int u = 0; u < BN_UNROLL; ++u) { const long long off = (r + (long long)u * rgroups) * a.C + cv * BN_VEC; BODY_USE }   \
    }                                                                                               \
    for (; r < r1; r += rgroups) {                                                                  \
      const int u = 0; const long long off = r * a.C + cv * BN_VEC; BODY_LOAD BODY_USE              \
    }                                                                                               \
  }

__global__ void __launch_bounds__(BN_THREADS, 1) bn_stats_kernel(const BnFwdArgs a) {
  const int tpc = a.C / BN_VEC;
  const int rgroups = BN_THREADS / tpc;
  const int cv = threadIdx.x % tpc, rg = threadIdx.x / tpc;
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  float acc[2][BN_VEC];
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
  uint2 xv[BN_UNROLL];
  BN_ROW_LOOP(
      { xv[u] = ldg8(a.x + off); },
      { float f[BN_VEC]; unpack4(xv[u], f);
        _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) { acc[0][i] += f[i]; acc[1][i] = fmaf(f[i], f[i], acc[1][i]); } (void)off; })
  cta_fold(acc, a.partial + (long long)blockIdx.x * 2 * a.C, a.C, rgroups, rg, cv);
  if (last_cta(a.counter)) {
    grid_fold(a.partial, a.C);
    const float inv_m = 1.0f / (float)a.M;
    for (int c = threadIdx.x; c < a.C; c += BN_THREADS) {
      const float m = bn_smem[c] * inv_m;
      float var = fmaf(-m, m, bn_smem[a.C + c] * inv_m);
      var = var < 0.f ? 0.f : var;
      a.mean[c] = m;
      a.invstd[c] = rsqrtf(var + a.eps);
      if (a.running_mean) {
        const float unbiased = a.M > 1 ? var * ((float)a.M / (float)(a.M - 1)) : var;
        a.running_mean[c] = fmaf(a.momentum, m - a.running_mean[c], a.running_mean[c]);
        a.running_var[c] = fmaf(a.momentum, unbiased - a.running_var[c], a.running_var[c]);
      }
    }
  }
}

__global__ void __launch_bounds__(BN_THREADS, 1) bn_apply_kernel(const BnFwdArgs a) {
  const int tpc = a.C / BN_VEC;
  const int rgroups = BN_THREADS / tpc;
  const int cv = threadIdx.x % tpc, rg = threadIdx.x / tpc;
  float scale[BN_VEC], shift[BN_VEC];
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) {
    const int c = cv * BN_VEC + i;
    const float sc = a.gamma[c] * a.invstd[c];
    scale[i] = sc;
    shift[i] = fmaf(-a.mean[c], sc, a.beta[c]);
  }
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  uint2 xv[BN_UNROLL], rv[BN_UNROLL];
  const bool has_res = a.res != nullptr;
  BN_ROW_LOOP(
      { xv[u] = ldg8(a.x + off); if (has_res) rv[u] = ldg8(a.res + off); },
      { float f[BN_VEC]; unpack4(xv[u], f);
        _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) f[i] = fmaf(f[i], scale[i], shift[i]);
        if (has_res) { float g[BN_VEC]; unpack4(rv[u], g); _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) f[i] += g[i]; }
        if (a.relu) { _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) f[i] = fmaxf(f[i], 0.f); }
        *reinterpret_cast<uint2*>(a.y + off) = pack4(f); })
}

__global__ void __launch_bounds__(BN_THREADS, 1) bn_bwd_reduce_kernel(const BnBwdArgs a) {
  const int tpc = a.C / BN_VEC;
  const int rgroups = BN_THREADS / tpc;
  const int cv = threadIdx.x % tpc, rg = threadIdx.x / tpc;
  float mean[BN_VEC], istd[BN_VEC];
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) { mean[i] = a.mean[cv * BN_VEC + i]; istd[i] = a.invstd[cv * BN_VEC + i]; }
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  float acc[2][BN_VEC];
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
  uint2 dv[BN_UNROLL], xv[BN_UNROLL], yv[BN_UNROLL];
  const bool relu = a.relu != 0;
  BN_ROW_LOOP(
      { dv[u] = ldg8(a.dy + off); xv[u] = ldg8(a.x + off); if (relu) yv[u] = ldg8(a.y + off); },
      { float d[BN_VEC]; float xf[BN_VEC]; unpack4(dv[u], d); unpack4(xv[u], xf);
        if (relu) { float yf[BN_VEC]; unpack4(yv[u], yf); _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) d[i] = yf[i] > 0.f ? d[i] : 0.f; }
        _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) { acc[0][i] += d[i]; acc[1][i] = fmaf(d[i], (xf[i] - mean[i]) * istd[i], acc[1][i]); } (void)off; })
  cta_fold(acc, a.partial + (long long)blockIdx.x * 2 * a.C, a.C, rgroups, rg, cv);
  if (last_cta(a.counter)) {
    grid_fold(a.partial, a.C);
    const float inv_m = 1.0f / (float)a.M;
    for (int c = threadIdx.x; c < a.C; c += BN_THREADS) {
      const float s = bn_smem[c], q = bn_smem[a.C + c];
      a.dbeta[c] = s;
      a.dgamma[c] = q;
      a.sums[c] = s * inv_m;
      a.sums[a.C + c] = q * inv_m;
    }
  }
}

__global__ void __launch_bounds__(BN_THREADS, 1) bn_bwd_apply_kernel(const BnBwdArgs a) {
  const int tpc = a.C / BN_VEC;
  const int rgroups = BN_THREADS / tpc;
  const int cv = threadIdx.x % tpc, rg = threadIdx.x / tpc;
  // dx = gs * (d - m1 - xhat * m2) with xhat = (x - mean) * istd   ==   A * d + B * x + Cc
  float cA[BN_VEC], cB[BN_VEC], cC[BN_VEC];
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) {
    const int c = cv * BN_VEC + i;
    const float istd = a.invstd[c], gs = a.gamma[c] * istd, m1 = a.sums[c], m2 = a.sums[a.C + c], mean = a.mean[c];
    cA[i] = gs;
    cB[i] = -gs * m2 * istd;
    cC[i] = gs * (m2 * istd * mean - m1);
  }
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  uint2 dv[BN_UNROLL], xv[BN_UNROLL], yv[BN_UNROLL];
  const bool relu = a.relu != 0;
  BN_ROW_LOOP(
      { dv[u] = ldg8(a.dy + off); xv[u] = ldg8(a.x + off); if (relu) yv[u] = ldg8(a.y + off); },
      { float d[BN_VEC]; float xf[BN_VEC]; unpack4(dv[u], d); unpack4(xv[u], xf);
        if (relu) { float yf[BN_VEC]; unpack4(yv[u], yf); _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) d[i] = yf[i] > 0.f ? d[i] : 0.f; }
        if (a.dres) *reinterpret_cast<uint2*>(a.dres + off) = pack4(d);
        float o[BN_VEC];
        _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) o[i] = fmaf(cA[i], d[i], fmaf(cB[i], xf[i], cC[i]));
        *reinterpret_cast<uint2*>(a.dx + off) = pack4(o); })
}


// ------------------------------------------------------------------------------------------------------------------
// Cooperative single-kernel variants: pass 1 (statistics / gradient sums), a grid-wide barrier, then pass 2 (apply) over
// the SAME rows by the same CTA -- the second read of x (dy, y) is an L2 hit, one launch and one reduction tail are gone.
// Launched with cudaLaunchCooperativeKernel (grid <= #SMs, one 1024-thread CTA per SM) so that all CTAs are co-resident
// and the barrier cannot deadlock even while push kernels of the side stream share the SMs.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int* bar) {
  __threadfence();                                 // every thread publishes its partial sums before the CTA arrives
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&bar[0], 1u);
    while (*reinterpret_cast<volatile unsigned int*>(&bar[0]) < gridDim.x) { __nanosleep(32); }
    __threadfence();
  }
  __syncthreads();
}
__device__ __forceinline__ void grid_barrier_release(unsigned int* bar) {       // at kernel end: reset for the next launch
  if (threadIdx.x == 0) {
    unsigned int prev = atomicAdd(&bar[1], 1u);
    if (prev == gridDim.x - 1) { bar[0] = 0; bar[1] = 0; __threadfence(); }
  }
}

__global__ void __launch_bounds__(BN_THREADS, 1) bn_fwd_coop_kernel(const BnFwdArgs a) {
  const int tpc = a.C / BN_VEC;
  const int rgroups = BN_THREADS / tpc;
  const int cv = threadIdx.x % tpc, rg = threadIdx.x / tpc;
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  {
    float acc[2][BN_VEC];
#pragma unroll
    for (int i = 0; i < BN_VEC; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
    uint2 xv[BN_UNROLL];
    BN_ROW_LOOP(
        { xv[u] = ldg8(a.x + off); },
        { float f[BN_VEC]; unpack4(xv[u], f);
          _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) { acc[0][i] += f[i]; acc[1][i] = fmaf(f[i], f[i], acc[1][i]); } (void)off; })
    cta_fold(acc, a.partial + (long long)blockIdx.x * 2 * a.C, a.C, rgroups, rg, cv);
  }
  grid_barrier(a.counter + 2);
  grid_fold(a.partial, a.C);                       // every CTA folds the same partials in the same order
  const float inv_m = 1.0f / (float)a.M;
  float scale[BN_VEC], shift[BN_VEC];
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) {
    const int c = cv * BN_VEC + i;
    const float m = bn_smem[c] * inv_m;
    float var = fmaf(-m, m, bn_smem[a.C + c] * inv_m);
    var = var < 0.f ? 0.f : var;
    const float istd = rsqrtf(var + a.eps);
    const float sc = a.gamma[c] * istd;
    scale[i] = sc;
    shift[i] = fmaf(-m, sc, a.beta[c]);
    if (blockIdx.x == 0 && rg == 0) {
      a.mean[c] = m;
      a.invstd[c] = istd;
      if (a.running_mean) {
        const float unbiased = a.M > 1 ? var * ((float)a.M / (float)(a.M - 1)) : var;
        a.running_mean[c] = fmaf(a.momentum, m - a.running_mean[c], a.running_mean[c]);
        a.running_var[c] = fmaf(a.momentum, unbiased - a.running_var[c], a.running_var[c]);
      }
    }
  }
  uint2 xv[BN_UNROLL], rv[BN_UNROLL];
  const bool has_res = a.res != nullptr;
  BN_ROW_LOOP(
      { xv[u] = ldg8(a.x + off); if (has_res) rv[u] = ldg8(a.res + off); },
      { float f[BN_VEC]; unpack4(xv[u], f);
        _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) f[i] = fmaf(f[i], scale[i], shift[i]);
        if (has_res) { float g[BN_VEC]; unpack4(rv[u], g); _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) f[i] += g[i]; }
        if (a.relu) { _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) f[i] = fmaxf(f[i], 0.f); }
        *reinterpret_cast<uint2*>(a.y + off) = pack4(f); })
  grid_barrier_release(a.counter + 2);
}

__global__ void __launch_bounds__(BN_THREADS, 1) bn_bwd_coop_kernel(const BnBwdArgs a) {
  const int tpc = a.C / BN_VEC;
  const int rgroups = BN_THREADS / tpc;
  const int cv = threadIdx.x % tpc, rg = threadIdx.x / tpc;
  float mean[BN_VEC], istd[BN_VEC];
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) { mean[i] = a.mean[cv * BN_VEC + i]; istd[i] = a.invstd[cv * BN_VEC + i]; }
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  const bool relu = a.relu != 0;
  {
    float acc[2][BN_VEC];
#pragma unroll
    for (int i = 0; i < BN_VEC; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
    uint2 dv[BN_UNROLL], xv[BN_UNROLL], yv[BN_UNROLL];
    BN_ROW_LOOP(
        { dv[u] = ldg8(a.dy + off); xv[u] = ldg8(a.x + off); if (relu) yv[u] = ldg8(a.y + off); },
        { float d[BN_VEC]; float xf[BN_VEC]; unpack4(dv[u], d); unpack4(xv[u], xf);
          if (relu) { float yf[BN_VEC]; unpack4(yv[u], yf); _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) d[i] = yf[i] > 0.f ? d[i] : 0.f; }
          _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) { acc[0][i] += d[i]; acc[1][i] = fmaf(d[i], (xf[i] - mean[i]) * istd[i], acc[1][i]); } (void)off; })
    cta_fold(acc, a.partial + (long long)blockIdx.x * 2 * a.C, a.C, rgroups, rg, cv);
  }
  grid_barrier(a.counter + 2);
  grid_fold(a.partial, a.C);
  const float inv_m = 1.0f / (float)a.M;
  float cA[BN_VEC], cB[BN_VEC], cC[BN_VEC];
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) {
    const int c = cv * BN_VEC + i;
    const float s = bn_smem[c], q = bn_smem[a.C + c];
    if (blockIdx.x == 0 && rg == 0) { a.dbeta[c] = s; a.dgamma[c] = q; }
    const float gs = a.gamma[c] * istd[i], m1 = s * inv_m, m2 = q * inv_m;
    cA[i] = gs;
    cB[i] = -gs * m2 * istd[i];
    cC[i] = gs * (m2 * istd[i] * mean[i] - m1);
  }
  uint2 dv[BN_UNROLL], xv[BN_UNROLL], yv[BN_UNROLL];
  BN_ROW_LOOP(
      { dv[u] = ldg8(a.dy + off); xv[u] = ldg8(a.x + off); if (relu) yv[u] = ldg8(a.y + off); },
      { float d[BN_VEC]; float xf[BN_VEC]; unpack4(dv[u], d); unpack4(xv[u], xf);
        if (relu) { float yf[BN_VEC]; unpack4(yv[u], yf); _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) d[i] = yf[i] > 0.f ? d[i] : 0.f; }
        if (a.dres) *reinterpret_cast<uint2*>(a.dres + off) = pack4(d);
        float o[BN_VEC];
        _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) o[i] = fmaf(cA[i], d[i], fmaf(cB[i], xf[i], cC[i]));
        *reinterpret_cast<uint2*>(a.dx + off) = pack4(o); })
  grid_barrier_release(a.counter + 2);
}

bool supported(int C) { return C >= BN_VEC && C <= BN_VEC * BN_THREADS && (C & (C - 1)) == 0; }

int plan_rows(long long M, int C, int num_sms, int* grid) {
  const int rgroups = BN_THREADS / (C / BN_VEC);
  long long by_bytes = (M * (long long)C * 2 + 131071) / 131072;           // one CTA per ~128 KB of activations
  long long target = by_bytes < num_sms ? by_bytes : num_sms;
  if (target < 1) target = 1;
  long long rows = (M + target - 1) / target;
  rows = (rows + rgroups - 1) / rgroups * rgroups;
  if (rows < rgroups) rows = rgroups;
  *grid = (int)((M + rows - 1) / rows);
  return (int)rows;
}

size_t smem_bytes(int C) {
  // cta_fold: rgroups * 2 * C floats = BN_THREADS * BN_VEC * 2 floats (32 KB); grid_fold: 2C + parts * 2C <= 2C + BN_THREADS
  size_t a = (size_t)BN_THREADS * BN_VEC * 2 * sizeof(float);
  size_t b = ((size_t)2 * C + BN_THREADS + 2 * C) * sizeof(float);
  return a > b ? a : b;
}

}  // namespace

extern "C" int drc_bn_supported(int C) { return supported(C) ? 1 : 0; }

// workspace floats needed for `partial`
extern "C" long long drc_bn_workspace(long long M, int C, int num_sms) {
  if (!supported(C)) return -1;
  int grid; plan_rows(M, C, num_sms, &grid);
  return (long long)grid * 2 * C;
}

extern "C" int drc_bn_fwd(const void* x, const void* res, void* y, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, float* mean, float* invstd, float* partial, unsigned int* counter, long long M, int C,
                          float eps, float momentum, int relu, int num_sms, int coop, cudaStream_t stream) {
  if (!supported(C)) return -1;
  BnFwdArgs a;
  a.x = (const __nv_bfloat16*)x; a.res = (const __nv_bfloat16*)res; a.y = (__nv_bfloat16*)y; a.gamma = gamma; a.beta = beta;
  a.running_mean = running_mean; a.running_var = running_var; a.mean = mean; a.invstd = invstd; a.partial = partial;
  a.counter = counter; a.M = M; a.C = C; a.eps = eps; a.momentum = momentum; a.relu = relu;
  int grid; a.rows_per_cta = plan_rows(M, C, num_sms, &grid);
  if (coop) {
    void* kargs[] = {&a};
    return (int)cudaLaunchCooperativeKernel((const void*)bn_fwd_coop_kernel, dim3(grid), dim3(BN_THREADS), kargs, smem_bytes(C), stream);
  }
  bn_stats_kernel<<<grid, BN_THREADS, smem_bytes(C), stream>>>(a);
  bn_apply_kernel<<<grid, BN_THREADS, 0, stream>>>(a);
  return (int)cudaGetLastError();
}

extern "C" int drc_bn_bwd(const void* dy, const void* y, const void* x, const float* gamma, const float* mean, const float* invstd,
                          void* dx, void* dres, float* dgamma, float* dbeta, float* partial, float* sums, unsigned int* counter,
                          long long M, int C, int relu, int num_sms, int coop, cudaStream_t stream) {
  if (!supported(C)) return -1;
  BnBwdArgs a;
  a.dy = (const __nv_bfloat16*)dy; a.y = (const __nv_bfloat16*)y; a.x = (const __nv_bfloat16*)x; a.gamma = gamma; a.mean = mean;
  a.invstd = invstd; a.dx = (__nv_bfloat16*)dx; a.dres = (__nv_bfloat16*)dres; a.dgamma = dgamma; a.dbeta = dbeta;
  a.partial = partial; a.sums = sums; a.counter = counter; a.M = M; a.C = C; a.relu = relu;
  int grid; a.rows_per_cta = plan_rows(M, C, num_sms, &grid);
  if (coop) {
    void* kargs[] = {&a};
    return (int)cudaLaunchCooperativeKernel((const void*)bn_bwd_coop_kernel, dim3(grid), dim3(BN_THREADS), kargs, smem_bytes(C), stream);
  }
  bn_bwd_reduce_kernel<<<grid, BN_THREADS, smem_bytes(C), stream>>>(a);
  bn_bwd_apply_kernel<<<grid, BN_THREADS, 0, stream>>>(a);
  return (int)cudaGetLastError();
}
