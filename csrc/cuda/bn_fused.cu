// K10 (normalisation part): fused training-mode BatchNorm + residual add + ReLU for NHWC bf16 activations.
//
// A launch list of the flagship step (profiles/launches_r1.md) shows ATen's channels-last BatchNorm kernels and the
// separate ReLU / add / threshold_backward passes taking ~60 % of a worker's forward+backward while running at 5-30 %
// of HBM bandwidth.  These kernels do the same math in the minimum number of passes with 16-byte accesses:
//
//   forward   bn_stats_kernel     : read x           -> per-channel mean / invstd (+ running statistics)
//             bn_apply_kernel     : read x (+res)    -> y = relu?( gamma * (x - mean) * invstd + beta (+ res) )
//   backward  bn_bwd_reduce_kernel: read dy, y, x    -> dgamma, dbeta (ReLU mask recomputed from y > 0)
//             bn_bwd_apply_kernel : read dy, y, x    -> dx (and d_residual)
//
// x is viewed as [M = N*H*W][C]; a thread owns 8 consecutive channels (one 16-byte word) and walks rows, so every
// access is a full-width coalesced vector.  Reductions are two-stage with a FIXED summation order (per-CTA partials in
// a workspace, the last CTA to finish folds them in index order): bit-identical results on every replica, which the
// exact-equality majority vote requires (reference: src/master/rep_master.py:162).  No float atomics.
//
// Reference counterpart: nn.BatchNorm2d / F.relu inside src/model_ops/resnet.py:14-64 and vgg.py:46-59 (PyTorch-0.3 CPU).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int BN_THREADS = 256;

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

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(p[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// Fold the per-thread accumulators of the CTA's row groups into partial[blockIdx][which][channel], fixed order.
template <int NACC>
__device__ __forceinline__ void cta_fold(float (&acc)[NACC][8], float* partial_blk, int C, int tpc, int rgroups, int rg, int cv,
                                         float* smem) {
  // smem layout: [rgroups][NACC][C]
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) smem[(rg * NACC + a) * C + cv * 8 + i] = acc[a][i];
  __syncthreads();
  for (int idx = threadIdx.x; idx < NACC * C; idx += BN_THREADS) {
    const int a = idx / C, c = idx - a * C;
    float s = 0.f;
    for (int g = 0; g < rgroups; ++g) s += smem[(g * NACC + a) * C + c];
    partial_blk[a * C + c] = s;
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

// Sum partial[b][which][c] over b with four interleaved accumulators (fixed order -> deterministic, 4x the MLP).
__device__ __forceinline__ float fold_partials(const float* partial, unsigned int nblk, int C, int which, int c) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const long long stride = 2LL * C;
  const float* p = partial + (long long)which * C + c;
  unsigned int b = 0;
  for (; b + 4 <= nblk; b += 4) {
    s0 += p[(long long)b * stride]; s1 += p[(long long)(b + 1) * stride];
    s2 += p[(long long)(b + 2) * stride]; s3 += p[(long long)(b + 3) * stride];
  }
  for (; b < nblk; ++b) s0 += p[(long long)b * stride];
  return (s0 + s1) + (s2 + s3);
}

extern __shared__ float bn_smem[];

__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const BnFwdArgs a) {
  const int tpc = a.C >> 3;                         // threads per row
  const int rgroups = BN_THREADS / tpc;
  const int cv = threadIdx.x % tpc, rg = threadIdx.x / tpc;
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
  if (rg < rgroups) {
    for (long long r = r0 + rg; r < r1; r += rgroups) {
      float f[8];
      unpack8(ldg16(a.x + r * a.C + cv * 8), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc[0][i] += f[i]; acc[1][i] = fmaf(f[i], f[i], acc[1][i]); }
    }
  }
  cta_fold<2>(acc, a.partial + (long long)blockIdx.x * 2 * a.C, a.C, tpc, rgroups, rg, cv, bn_smem);
  if (last_cta(a.counter)) {
    const float inv_m = 1.0f / (float)a.M;
    for (int c = threadIdx.x; c < a.C; c += BN_THREADS) {
      const float s = fold_partials(a.partial, gridDim.x, a.C, 0, c), q = fold_partials(a.partial, gridDim.x, a.C, 1, c);
      const float m = s * inv_m;
      float var = fmaf(-m, m, q * inv_m);
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

__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const BnFwdArgs a) {
  const int tpc = a.C >> 3;
  const int rgroups = BN_THREADS / tpc;
  const int cv = threadIdx.x % tpc, rg = threadIdx.x / tpc;
  float scale[8], shift[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = cv * 8 + i;
    const float sc = a.gamma[c] * a.invstd[c];
    scale[i] = sc;
    shift[i] = fmaf(-a.mean[c], sc, a.beta[c]);
  }
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  for (long long r = r0 + rg; r < r1; r += rgroups) {
    const long long off = r * a.C + cv * 8;
    float f[8];
    unpack8(ldg16(a.x + off), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], scale[i], shift[i]);
    if (a.res) {
      float g[8];
      unpack8(ldg16(a.res + off), g);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += g[i];
    }
    if (a.relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    *reinterpret_cast<uint4*>(a.y + off) = pack8(f);
  }
}

__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(const BnBwdArgs a) {
  const int tpc = a.C >> 3;
  const int rgroups = BN_THREADS / tpc;
  const int cv = threadIdx.x % tpc, rg = threadIdx.x / tpc;
  float mean[8], istd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { mean[i] = a.mean[cv * 8 + i]; istd[i] = a.invstd[cv * 8 + i]; }
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
  for (long long r = r0 + rg; r < r1; r += rgroups) {
    const long long off = r * a.C + cv * 8;
    float d[8], xv[8];
    unpack8(ldg16(a.dy + off), d);
    unpack8(ldg16(a.x + off), xv);
    if (a.relu) {
      float yv[8];
      unpack8(ldg16(a.y + off), yv);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = yv[i] > 0.f ? d[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[0][i] += d[i];
      acc[1][i] = fmaf(d[i], (xv[i] - mean[i]) * istd[i], acc[1][i]);
    }
  }
  cta_fold<2>(acc, a.partial + (long long)blockIdx.x * 2 * a.C, a.C, tpc, rgroups, rg, cv, bn_smem);
  if (last_cta(a.counter)) {
    const float inv_m = 1.0f / (float)a.M;
    for (int c = threadIdx.x; c < a.C; c += BN_THREADS) {
      const float s = fold_partials(a.partial, gridDim.x, a.C, 0, c), q = fold_partials(a.partial, gridDim.x, a.C, 1, c);
      a.dbeta[c] = s;
      a.dgamma[c] = q;
      a.sums[c] = s * inv_m;
      a.sums[a.C + c] = q * inv_m;
    }
  }
}

__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const BnBwdArgs a) {
  const int tpc = a.C >> 3;
  const int rgroups = BN_THREADS / tpc;
  const int cv = threadIdx.x % tpc, rg = threadIdx.x / tpc;
  float mean[8], istd[8], gs[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = cv * 8 + i;
    mean[i] = a.mean[c]; istd[i] = a.invstd[c]; gs[i] = a.gamma[c] * a.invstd[c];
    m1[i] = a.sums[c]; m2[i] = a.sums[a.C + c];
  }
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  for (long long r = r0 + rg; r < r1; r += rgroups) {
    const long long off = r * a.C + cv * 8;
    float d[8], xv[8];
    unpack8(ldg16(a.dy + off), d);
    unpack8(ldg16(a.x + off), xv);
    if (a.relu) {
      float yv[8];
      unpack8(ldg16(a.y + off), yv);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = yv[i] > 0.f ? d[i] : 0.f;
    }
    if (a.dres) *reinterpret_cast<uint4*>(a.dres + off) = pack8(d);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xhat = (xv[i] - mean[i]) * istd[i];
      o[i] = gs[i] * (d[i] - m1[i] - xhat * m2[i]);
    }
    *reinterpret_cast<uint4*>(a.dx + off) = pack8(o);
  }
}

int plan_rows(long long M, int C, int num_sms, int* grid) {
  // These tensors are small (2-17 MB for ResNet-18 at B=128): a CTA should stream >= 64 KB so that the serial fold of
  // the per-CTA partials by the last CTA (grid * 2 * C floats) stays a small fraction of the kernel.
  const int rgroups = BN_THREADS / (C >> 3);
  long long by_bytes = (M * (long long)C * 2 + 65535) / 65536;
  long long cap = 32768 / C; if (cap > 2LL * num_sms) cap = 2LL * num_sms; if (cap < 1) cap = 1;
  long long target = by_bytes < cap ? by_bytes : cap; if (target < 1) target = 1;
  long long rows = (M + target - 1) / target;
  rows = (rows + rgroups - 1) / rgroups * rgroups;               // whole row-group iterations
  if (rows < rgroups) rows = rgroups;
  *grid = (int)((M + rows - 1) / rows);
  return (int)rows;
}

bool supported(int C) { return C >= 8 && C <= 2048 && (C & (C - 1)) == 0; }

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
                          float eps, float momentum, int relu, int num_sms, cudaStream_t stream) {
  if (!supported(C)) return -1;
  BnFwdArgs a;
  a.x = (const __nv_bfloat16*)x; a.res = (const __nv_bfloat16*)res; a.y = (__nv_bfloat16*)y; a.gamma = gamma; a.beta = beta;
  a.running_mean = running_mean; a.running_var = running_var; a.mean = mean; a.invstd = invstd; a.partial = partial;
  a.counter = counter; a.M = M; a.C = C; a.eps = eps; a.momentum = momentum; a.relu = relu;
  int grid; a.rows_per_cta = plan_rows(M, C, num_sms, &grid);
  const int rgroups = BN_THREADS / (C >> 3);
  const size_t smem = (size_t)rgroups * 2 * C * sizeof(float);    // = 256/ (C/8) * 2 * C * 4 = 16 KB
  bn_stats_kernel<<<grid, BN_THREADS, smem, stream>>>(a);
  bn_apply_kernel<<<grid, BN_THREADS, 0, stream>>>(a);
  return (int)cudaGetLastError();
}

extern "C" int drc_bn_bwd(const void* dy, const void* y, const void* x, const float* gamma, const float* mean, const float* invstd,
                          void* dx, void* dres, float* dgamma, float* dbeta, float* partial, float* sums, unsigned int* counter,
                          long long M, int C, int relu, int num_sms, cudaStream_t stream) {
  if (!supported(C)) return -1;
  BnBwdArgs a;
  a.dy = (const __nv_bfloat16*)dy; a.y = (const __nv_bfloat16*)y; a.x = (const __nv_bfloat16*)x; a.gamma = gamma; a.mean = mean;
  a.invstd = invstd; a.dx = (__nv_bfloat16*)dx; a.dres = (__nv_bfloat16*)dres; a.dgamma = dgamma; a.dbeta = dbeta;
  a.partial = partial; a.sums = sums; a.counter = counter; a.M = M; a.C = C; a.relu = relu;
  int grid; a.rows_per_cta = plan_rows(M, C, num_sms, &grid);
  const int rgroups = BN_THREADS / (C >> 3);
  const size_t smem = (size_t)rgroups * 2 * C * sizeof(float);
  bn_bwd_reduce_kernel<<<grid, BN_THREADS, smem, stream>>>(a);
  bn_bwd_apply_kernel<<<grid, BN_THREADS, 0, stream>>>(a);
  return (int)cudaGetLastError();
}
