// K10 (normalisation part): fused training-mode BatchNorm + residual add + ReLU for NHWC bf16 activations (third iteration).
//
// What changed against round 1 (profiles/ncu_bn.md: 10-14 us per launch at 14-35 % of HBM, four launches per layer):
//   * the forward statistics normally come out of the producing convolution's epilogue (conv_epilogue.cuh), so a layer's
//     forward is ONE streaming kernel: bn_apply_kernel (x, mean, invstd, residual -> y, ReLU fused).  bn_stats_kernel remains
//     for producers that are not ours (GEMM-served 1x1 convolutions, library fallbacks);
//   * every access is 16 bytes (8 channels per thread), 4 independent rows in flight per thread, 512-thread CTAs with up to
//     four resident per SM, and the grid is sized from the tensor (one CTA per 32 KB up to 4 per SM): the 2-4 MB late-stage
//     tensors spread over 64-128 CTAs instead of 16-32, the 16.8 MB ones keep ~64 KB of loads in flight per SM;
//   * reductions: per-thread fp32 accumulators -> fixed-order shared-memory fold -> per-CTA partial [2][C] -> the last CTA
//     (atomic ticket) folds the partials with 16-byte loads in a fixed order.  No cooperative launch, no grid barrier, no
//     cluster: safe next to any other kernel on any number of concurrent streams, and bit-deterministic (replicas of a batch
//     on different GPUs must agree exactly for the majority vote, reference: src/master/rep_master.py:162).
//
// backward: bn_bwd_reduce_kernel (dy, y, x -> dgamma, dbeta, means)    bn_bwd_apply_kernel (dy, y, x -> dx, dres)
//
// Reference counterpart: nn.BatchNorm2d / F.relu inside src/model_ops/resnet.py:14-64 and vgg.py:46-59 (PyTorch-0.3 CPU).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int BN_THREADS = 512;
constexpr int BN_VEC = 8;                      // channels per thread (one 16-byte access)
constexpr int BN_UNROLL = 4;

struct BnFwdArgs {
  const __nv_bfloat16* x;         // [M][C]
  const __nv_bfloat16* res;       // optional residual [M][C]
  __nv_bfloat16* y;               // [M][C]
  const float* gamma;             // [C]
  const float* beta;              // [C]
  float* running_mean;            // [C] (may be null)
  float* running_var;
  float* mean;                    // [C] saved for backward (written by the statistics kernel, read by apply)
  float* invstd;                  // [C]
  float* partial;                 // [grid][2][C] workspace (statistics kernel)
  unsigned int* counter;          // zero on entry, reset by the last CTA
  long long M;
  int C;
  int rows_per_cta;
  float eps, momentum;
  int relu;
};

struct BnBwdArgs {
  const __nv_bfloat16* dy;        // [M][C]
  const __nv_bfloat16* y;         // forward output (ReLU mask); null when relu == 0 -- or when the mask is recomputed from x
  const float* beta;              // relu && !y: mask = (bf16(fma(x, gamma*invstd, beta - mean*gamma*invstd)) > 0), the forward's own
                                  // arithmetic (layers without a residual input) -- one tensor less to read in both kernels
  const __nv_bfloat16* x;         // [M][C]
  const float* gamma;
  const float* mean;
  const float* invstd;
  __nv_bfloat16* dx;              // [M][C]
  __nv_bfloat16* dres;            // optional: gradient of the residual input
  float* dgamma;                  // [C]
  float* dbeta;                   // [C]
  float* partial;                 // [grid][2][C]
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
  for (int i = 0; i < 4; ++i) { const float2 a = __bfloat1622float2(p[i]); f[2 * i] = a.x; f[2 * i + 1] = a.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
// did the forward's BatchNorm + ReLU output survive?  Same fp32 arithmetic and the same bf16 rounding as bn_apply_kernel.
__device__ __forceinline__ bool relu_mask(float x, float scale, float shift) {
  return __bfloat162float(__float2bfloat16_rn(fmaf(x, scale, shift))) > 0.f;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

extern __shared__ __align__(16) float bn_smem[];

// CTA-level fold of the per-thread accumulators acc[2][8] over the row groups, then this CTA's partial[2][C] (fixed order).
__device__ __forceinline__ void cta_fold(const float (&acc)[2][BN_VEC], float* partial_blk, int C, int rgroups, int rg, int cv) {
  // smem layout: [rgroups][2][C]  (rgroups * 2 * C == BN_THREADS * 16 floats = 32 KB)
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    float4* dst = reinterpret_cast<float4*>(&bn_smem[(rg * 2 + a) * C + cv * BN_VEC]);
    dst[0] = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    dst[1] = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
  }
  __syncthreads();
  const int n_idx = 2 * C;
  for (int idx = threadIdx.x; idx < n_idx; idx += BN_THREADS) {
    float s0 = 0.f, s1 = 0.f;
    int g = 0;
    for (; g + 1 < rgroups; g += 2) { s0 += bn_smem[g * n_idx + idx]; s1 += bn_smem[(g + 1) * n_idx + idx]; }
    if (g < rgroups) s0 += bn_smem[g * n_idx + idx];
    partial_blk[idx] = s0 + s1;
  }
}

__device__ __forceinline__ bool last_cta(unsigned int* counter) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    // release the CTA's partial (ordered before this thread by the barrier) with the ticket, acquire everybody else's
    unsigned int prev;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(counter) : "memory");
    s_last = (prev == gridDim.x - 1);
    if (s_last) *counter = 0;
  }
  __syncthreads();
  return s_last != 0;
}

// Grid-level fold by the last CTA: totals[idx] = sum_b partial[b][idx], idx in [0, 2C): 16-byte loads, the CTAs' partials split
// into interleaved subsets that are combined in a fixed order.  Result in bn_smem[0 .. 2C).
__device__ __forceinline__ void grid_fold(const float* partial, int C) {
  const int items4 = (2 * C) >> 2;
  const int nblk = (int)gridDim.x;
  const float4* p4 = reinterpret_cast<const float4*>(partial);
  float4* s4 = reinterpret_cast<float4*>(bn_smem);
  __syncthreads();
  if (items4 <= BN_THREADS) {
    const int SUB = BN_THREADS / items4;
    const int item = threadIdx.x % items4, sub = threadIdx.x / items4;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sub < SUB) {
      constexpr int FB = 16;                                        // independent 16-byte loads in flight
      for (int g0 = sub; g0 < nblk; g0 += FB * SUB) {
        float4 v[FB];
#pragma unroll
        for (int u = 0; u < FB; ++u) {
          const int g = g0 + u * SUB;
          v[u] = g < nblk ? __ldcg(p4 + (long long)g * items4 + item) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < FB; ++u) { t.x += v[u].x; t.y += v[u].y; t.z += v[u].z; t.w += v[u].w; }    // fixed order
      }
    }
    s4[items4 + threadIdx.x] = t;                                  // scratch behind the totals
    __syncthreads();
    if (sub == 0) {
      float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s2 = 0; s2 < SUB; ++s2) {
        const float4 v = s4[items4 + s2 * items4 + item];
        tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
      }
      s4[item] = tot;
    }
  } else {
    for (int item = threadIdx.x; item < items4; item += BN_THREADS) {
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
      int g = 0;
      for (; g + 1 < nblk; g += 2) {
        const float4 v0 = __ldcg(p4 + (long long)g * items4 + item), v1 = __ldcg(p4 + (long long)(g + 1) * items4 + item);
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
      }
      if (g < nblk) { const float4 v0 = __ldcg(p4 + (long long)g * items4 + item); a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w; }
      s4[item] = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
    }
  }
  __syncthreads();
}

// rows [r0, r1) of this CTA, row groups interleaved, BN_UNROLL independent rows in flight per thread
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

__global__ void __launch_bounds__(BN_THREADS, 2) bn_stats_kernel(const BnFwdArgs a) {
  const int tpc = a.C / BN_VEC;
  const int rgroups = BN_THREADS / tpc;
  const int cv = threadIdx.x % tpc, rg = threadIdx.x / tpc;
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  float acc[2][BN_VEC];
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
  uint4 xv[BN_UNROLL];
  BN_ROW_LOOP(
      { xv[u] = ldg16(a.x + off); },
      { float f[BN_VEC]; unpack8(xv[u], f);
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

__global__ void __launch_bounds__(BN_THREADS, 2) bn_apply_kernel(const BnFwdArgs a) {
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
  uint4 xv[BN_UNROLL], rv[BN_UNROLL];
  const bool has_res = a.res != nullptr;
  BN_ROW_LOOP(
      { xv[u] = ldg16(a.x + off); if (has_res) rv[u] = ldg16(a.res + off); },
      { float f[BN_VEC]; unpack8(xv[u], f);
        _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) f[i] = fmaf(f[i], scale[i], shift[i]);
        if (has_res) { float g[BN_VEC]; unpack8(rv[u], g); _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) f[i] += g[i]; }
        if (a.relu) { _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) f[i] = fmaxf(f[i], 0.f); }
        *reinterpret_cast<uint4*>(a.y + off) = pack8(f); })
}

// MASK: 0 = no ReLU, 1 = ReLU mask from the saved output y, 2 = mask recomputed from x (specialised: the apply kernel has to fit
// 64 registers for two CTAs per SM)
template <int MASK>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(const BnBwdArgs a) {
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
  uint4 dv[BN_UNROLL], xv[BN_UNROLL], yv[BN_UNROLL];
  constexpr bool relu_y = MASK == 1, relu_x = MASK == 2;
  float sc[BN_VEC], sh[BN_VEC];
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) {
    sc[i] = relu_x ? a.gamma[cv * BN_VEC + i] * istd[i] : 0.f;
    sh[i] = relu_x ? fmaf(-mean[i], sc[i], a.beta[cv * BN_VEC + i]) : 0.f;
  }
  BN_ROW_LOOP(
      { dv[u] = ldg16(a.dy + off); xv[u] = ldg16(a.x + off); if (relu_y) yv[u] = ldg16(a.y + off); },
      { float d[BN_VEC]; float xf[BN_VEC]; unpack8(dv[u], d); unpack8(xv[u], xf);
        if (relu_y) { float yf[BN_VEC]; unpack8(yv[u], yf); _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) d[i] = yf[i] > 0.f ? d[i] : 0.f; }
        if (relu_x) { _Pragma("unroll") for (int i = 0; i < BN_VEC; ++i) d[i] = relu_mask(xf[i], sc[i], sh[i]) ? d[i] : 0.f; }
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

template <int MASK>
__global__ void __launch_bounds__(BN_THREADS, 2) bn_bwd_apply_kernel(const BnBwdArgs a) {
  const int tpc = a.C / BN_VEC;
  const int rgroups = BN_THREADS / tpc;
  const int cv = threadIdx.x % tpc, rg = threadIdx.x / tpc;
  // dx = gs * (d - m1 - xhat * m2) with xhat = (x - mean) * istd   ==   A * d + B * x + Cc
  float cA[BN_VEC], cB[BN_VEC], cC[BN_VEC];
  constexpr bool relu_y = MASK == 1, relu_x = MASK == 2;
  __shared__ __align__(16) float s_sh[BN_THREADS * BN_VEC];           // the forward's shift per channel (mask recomputation), read per row: keeps
  if (relu_x) {                                        // the kernel inside 64 registers
    for (int c = threadIdx.x; c < a.C; c += BN_THREADS) s_sh[c] = fmaf(-a.mean[c], a.gamma[c] * a.invstd[c], a.beta[c]);
    __syncthreads();
  }
  const float* sh = s_sh + cv * BN_VEC;
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) {
    const int c = cv * BN_VEC + i;
    const float istd = a.invstd[c], gs = a.gamma[c] * istd, m1 = a.sums[c], m2 = a.sums[a.C + c], mean = a.mean[c];
    cA[i] = gs;                                        // == the forward's scale
    cB[i] = -gs * m2 * istd;
    cC[i] = gs * (m2 * istd * mean - m1);
  }
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  constexpr int BU = 2;                                            // three streams in flight: keep the register count moderate
  long long r = r0 + rg;
  for (; r + (long long)(BU - 1) * rgroups < r1; r += (long long)BU * rgroups) {
    uint4 dv[BU], xv[BU], yv[BU];
#pragma unroll
    for (int u = 0; u < BU; ++u) {
      const long long off = (r + (long long)u * rgroups) * a.C + cv * BN_VEC;
      dv[u] = ldg16(a.dy + off); xv[u] = ldg16(a.x + off); if (relu_y) yv[u] = ldg16(a.y + off);
    }
#pragma unroll
    for (int u = 0; u < BU; ++u) {
      const long long off = (r + (long long)u * rgroups) * a.C + cv * BN_VEC;
      float d[BN_VEC], xf[BN_VEC];
      unpack8(dv[u], d); unpack8(xv[u], xf);
      if (relu_y) { float yf[BN_VEC]; unpack8(yv[u], yf);
#pragma unroll
        for (int i = 0; i < BN_VEC; ++i) d[i] = yf[i] > 0.f ? d[i] : 0.f; }
      if (relu_x) {
#pragma unroll
        for (int i = 0; i < BN_VEC; ++i) d[i] = relu_mask(xf[i], cA[i], sh[i]) ? d[i] : 0.f; }
      if (a.dres) *reinterpret_cast<uint4*>(a.dres + off) = pack8(d);
      float o[BN_VEC];
#pragma unroll
      for (int i = 0; i < BN_VEC; ++i) o[i] = fmaf(cA[i], d[i], fmaf(cB[i], xf[i], cC[i]));
      *reinterpret_cast<uint4*>(a.dx + off) = pack8(o);
    }
  }
  for (; r < r1; r += rgroups) {
    const long long off = r * a.C + cv * BN_VEC;
    float d[BN_VEC], xf[BN_VEC];
    const uint4 dvv = ldg16(a.dy + off), xvv = ldg16(a.x + off);
    unpack8(dvv, d); unpack8(xvv, xf);
    if (relu_y) { float yf[BN_VEC]; const uint4 yvv = ldg16(a.y + off); unpack8(yvv, yf);
#pragma unroll
      for (int i = 0; i < BN_VEC; ++i) d[i] = yf[i] > 0.f ? d[i] : 0.f; }
    if (relu_x) {
#pragma unroll
      for (int i = 0; i < BN_VEC; ++i) d[i] = relu_mask(xf[i], cA[i], sh[i]) ? d[i] : 0.f; }
    if (a.dres) *reinterpret_cast<uint4*>(a.dres + off) = pack8(d);
    float o[BN_VEC];
#pragma unroll
    for (int i = 0; i < BN_VEC; ++i) o[i] = fmaf(cA[i], d[i], fmaf(cB[i], xf[i], cC[i]));
    *reinterpret_cast<uint4*>(a.dx + off) = pack8(o);
  }
}

// C / 8 vector lanes must tile the 512-thread CTA
bool supported(int C) { return C >= BN_VEC && C % BN_VEC == 0 && BN_THREADS % (C / BN_VEC) == 0; }

// Grid / rows per CTA.  `reduce`: every CTA adds a partial the last CTA has to fold, so the grid is capped at the SM count;
// streaming kernels go up to 4 CTAs per SM (one CTA per ~32 KB of the tensor).
int plan_rows(long long M, int C, int num_sms, bool reduce, int* grid) {
  const int rgroups = BN_THREADS / (C / BN_VEC);
  const long long bytes = M * (long long)C * 2;
  long long target = (bytes + 32767) / 32768;
  const long long cap = reduce ? num_sms : 4LL * num_sms;
  if (target > cap) target = cap;
  if (target < 1) target = 1;
  long long rows = (M + target - 1) / target;
  rows = (rows + rgroups - 1) / rgroups * rgroups;
  if (rows < rgroups) rows = rgroups;
  *grid = (int)((M + rows - 1) / rows);
  return (int)rows;
}

size_t smem_bytes(int C) {
  // cta_fold: BN_THREADS * 16 floats (32 KB); grid_fold: totals 2C floats + BN_THREADS float4 of scratch
  size_t a = (size_t)BN_THREADS * 2 * BN_VEC * sizeof(float);
  size_t b = ((size_t)2 * C + 4 * BN_THREADS) * sizeof(float);
  return a > b ? a : b;
}

}  // namespace

extern "C" int drc_bn_supported(int C) { return supported(C) ? 1 : 0; }

// workspace floats needed for `partial`
extern "C" long long drc_bn_workspace(long long M, int C, int num_sms) {
  if (!supported(C)) return -1;
  int grid; plan_rows(M, C, num_sms, true, &grid);
  return (long long)grid * 2 * C;
}

// Forward.  have_stats != 0: mean / invstd were produced by the convolution epilogue -> only the apply kernel runs.
extern "C" int drc_bn_fwd(const void* x, const void* res, void* y, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, float* mean, float* invstd, float* partial, unsigned int* counter, long long M, int C,
                          float eps, float momentum, int relu, int num_sms, int have_stats, cudaStream_t stream) {
  if (!supported(C)) return -1;
  BnFwdArgs a;
  a.x = (const __nv_bfloat16*)x; a.res = (const __nv_bfloat16*)res; a.y = (__nv_bfloat16*)y; a.gamma = gamma; a.beta = beta;
  a.running_mean = running_mean; a.running_var = running_var; a.mean = mean; a.invstd = invstd; a.partial = partial;
  a.counter = counter; a.M = M; a.C = C; a.eps = eps; a.momentum = momentum; a.relu = relu;
  int grid;
  if (!have_stats) {
    a.rows_per_cta = plan_rows(M, C, num_sms, true, &grid);
    bn_stats_kernel<<<grid, BN_THREADS, smem_bytes(C), stream>>>(a);
  }
  a.rows_per_cta = plan_rows(M, C, num_sms, false, &grid);
  bn_apply_kernel<<<grid, BN_THREADS, 0, stream>>>(a);
  return (int)cudaGetLastError();
}

extern "C" int drc_bn_bwd(const void* dy, const void* y, const void* x, const float* gamma, const float* beta, const float* mean,
                          const float* invstd,
                          void* dx, void* dres, float* dgamma, float* dbeta, float* partial, float* sums, unsigned int* counter,
                          long long M, int C, int relu, int num_sms, cudaStream_t stream) {
  if (!supported(C)) return -1;
  BnBwdArgs a;
  a.dy = (const __nv_bfloat16*)dy; a.y = (const __nv_bfloat16*)y; a.x = (const __nv_bfloat16*)x; a.gamma = gamma; a.beta = beta;
  a.mean = mean; a.invstd = invstd;
  if (relu && !y && !beta) return -2; a.dx = (__nv_bfloat16*)dx; a.dres = (__nv_bfloat16*)dres; a.dgamma = dgamma; a.dbeta = dbeta;
  a.partial = partial; a.sums = sums; a.counter = counter; a.M = M; a.C = C; a.relu = relu;
  int grid;
  a.rows_per_cta = plan_rows(M, C, num_sms, true, &grid);
  const int mask = !relu ? 0 : (y ? 1 : 2);
  if (mask == 0) bn_bwd_reduce_kernel<0><<<grid, BN_THREADS, smem_bytes(C), stream>>>(a);
  else if (mask == 1) bn_bwd_reduce_kernel<1><<<grid, BN_THREADS, smem_bytes(C), stream>>>(a);
  else bn_bwd_reduce_kernel<2><<<grid, BN_THREADS, smem_bytes(C), stream>>>(a);
  a.rows_per_cta = plan_rows(M, C, num_sms, false, &grid);
  if (mask == 0) bn_bwd_apply_kernel<0><<<grid, BN_THREADS, 0, stream>>>(a);
  else if (mask == 1) bn_bwd_apply_kernel<1><<<grid, BN_THREADS, 0, stream>>>(a);
  else bn_bwd_apply_kernel<2><<<grid, BN_THREADS, 0, stream>>>(a);
  return (int)cudaGetLastError();
}

// Backward when the reduction already happened in the dgrad epilogue of the convolution that consumed this layer's output
// (conv_epilogue.cuh, BatchNorm-backward mode): dz arrives masked, sums = [mean(dz), mean(dz * xhat)] -> only the apply kernel.
extern "C" int drc_bn_bwd_apply(const void* dz, const void* x, const float* gamma, const float* mean, const float* invstd,
                                const float* sums, void* dx, long long M, int C, int num_sms, cudaStream_t stream) {
  if (!supported(C)) return -1;
  BnBwdArgs a;
  a.dy = (const __nv_bfloat16*)dz; a.y = nullptr; a.x = (const __nv_bfloat16*)x; a.gamma = gamma; a.beta = nullptr;
  a.mean = mean; a.invstd = invstd; a.dx = (__nv_bfloat16*)dx; a.dres = nullptr; a.dgamma = nullptr; a.dbeta = nullptr;
  a.partial = nullptr; a.sums = const_cast<float*>(sums); a.counter = nullptr; a.M = M; a.C = C; a.relu = 0;
  int grid;
  a.rows_per_cta = plan_rows(M, C, num_sms, false, &grid);
  bn_bwd_apply_kernel<0><<<grid, BN_THREADS, 0, stream>>>(a);
  return (int)cudaGetLastError();
}
