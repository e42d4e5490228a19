// K10 (normalisation part, continued): single-launch BatchNorm (+ residual + ReLU) on thread-block CLUSTERS for the small
// late-stage activations.
//
// Why: half of the 20 BatchNorm layers of ResNet-18 (layers 3 and 4, 2-4 MB per tensor at B = 128) are pure latency in the
// streaming kernels of bn_fused.cu -- ~7-13 us per launch for ~1 us of data movement, four launches per layer per step
// (profiles/worker_profile_ResNet18_fused.txt: 0.8 ms of a 2.0 ms step in BatchNorm).  A tensor that small fits in the
// shared memory of one cluster, so the whole layer can be ONE launch with ONE pass over global memory:
//
//   * the channel range is cut into slices (32-128 channels); a cluster of 8 (or 16) CTAs owns one slice for ALL rows, so
//     clusters never talk to each other -- no grid barrier, no "last CTA", safe next to any other kernel on the GPU;
//   * forward : every CTA streams its rows of the slice into shared memory while accumulating sum / sum-of-squares, the
//     per-CTA partials are exchanged through distributed shared memory after one cluster barrier and folded in rank
//     order by every CTA (identical result everywhere, fixed order => bit-deterministic), then y = relu(bn(x) + res) is
//     produced straight from the shared-memory copy: x is read from global memory once;
//   * backward: pass 1 keeps the ReLU-masked dy in shared memory while accumulating sum(dy), sum(dy * xhat); after the
//     cluster fold pass 2 re-reads only x (an L2 hit) and writes dx (and dres).
//
// Numerics are those of bn_fused.cu (fp32 sums, biased variance for normalisation, unbiased for the running estimate).
//
// STATUS: compiled for sm_100a, not yet run on hardware (written after the round's GPU budget was spent): opt-in via
// DRACO_BN_CLUSTER=1, test gated by DRACO_EXPERIMENTAL=1.
//
// Reference counterpart: nn.BatchNorm2d / F.relu inside src/model_ops/resnet.py:14-64 (PyTorch-0.3 CPU).
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cg = cooperative_groups;

namespace {

constexpr int CL_THREADS = 512;
constexpr int CL_VEC = 4;
constexpr int CL_SCRATCH_FLOATS = CL_THREADS * 2 * CL_VEC;      // [rgroups][2][CS] == 512 * 8 floats = 16 KB for every CS
constexpr int CL_MAX_CS = 128;

struct ClFwdArgs {
  const __nv_bfloat16* x;
  const __nv_bfloat16* res;       // optional
  __nv_bfloat16* y;
  const float* gamma;
  const float* beta;
  float* running_mean;            // may be null
  float* running_var;
  float* mean;                    // [C] out
  float* invstd;                  // [C] out
  long long M;
  int C, CS, K;                   // channels, channels per cluster, CTAs per cluster
  int rows_per_cta;
  float eps, momentum;
  int relu;
};

struct ClBwdArgs {
  const __nv_bfloat16* dy;
  const __nv_bfloat16* y;         // forward output (ReLU mask) or null
  const __nv_bfloat16* x;
  const float* gamma;
  const float* mean;
  const float* invstd;
  __nv_bfloat16* dx;
  __nv_bfloat16* dres;            // optional
  float* dgamma;
  float* dbeta;
  long long M;
  int C, CS, K;
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

extern __shared__ __align__(16) uint8_t cl_smem[];

// shared-memory carve-up: [scratch 16 KB][partial 2*CS floats][total 2*CS floats][coef 3*CS floats][tile rows*CS bf16]
struct Smem {
  float* scratch; float* partial; float* total; float* coef; __nv_bfloat16* tile;
};
__device__ __forceinline__ Smem carve(int CS) {
  Smem s;
  s.scratch = reinterpret_cast<float*>(cl_smem);
  s.partial = s.scratch + CL_SCRATCH_FLOATS;
  s.total = s.partial + 2 * CL_MAX_CS;
  s.coef = s.total + 2 * CL_MAX_CS;
  s.tile = reinterpret_cast<__nv_bfloat16*>(s.coef + 3 * CL_MAX_CS);
  (void)CS;
  return s;
}
constexpr int CL_HEADER_BYTES = (CL_SCRATCH_FLOATS + 7 * CL_MAX_CS) * 4;

// fold the per-thread accumulators of the CTA (fixed order), then the CTAs of the cluster in rank order -> s.total[2*CS]
__device__ __forceinline__ void fold(const Smem& s, const float (&acc)[2][CL_VEC], int CS, int rgroups, int rg, int cv,
                                     cg::cluster_group& cluster, int K) {
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int i = 0; i < CL_VEC; ++i) s.scratch[(rg * 2 + a) * CS + cv * CL_VEC + i] = acc[a][i];
  __syncthreads();
  const int n_idx = 2 * CS;
  for (int idx = threadIdx.x; idx < n_idx; idx += CL_THREADS) {
    float t = 0.f;
    for (int g = 0; g < rgroups; ++g) t += s.scratch[g * n_idx + idx];
    s.partial[idx] = t;
  }
  cluster.sync();                                             // every CTA's partial is visible cluster-wide
  for (int idx = threadIdx.x; idx < n_idx; idx += CL_THREADS) {
    float t = 0.f;
    for (int k = 0; k < K; ++k) t += cluster.map_shared_rank(s.partial, k)[idx];      // DSMEM reads, rank order
    s.total[idx] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(CL_THREADS, 1) bn_fwd_cluster_kernel(const ClFwdArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int slice = blockIdx.x / a.K;
  const int c0 = slice * a.CS;
  const Smem s = carve(a.CS);
  const int tpr = a.CS / CL_VEC, rgroups = CL_THREADS / tpr;
  const int cv = threadIdx.x % tpr, rg = threadIdx.x / tpr;
  const long long r0 = (long long)rank * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;

  float acc[2][CL_VEC];
#pragma unroll
  for (int i = 0; i < CL_VEC; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
  {
    long long r = r0 + rg;
    for (; r + 3LL * rgroups < r1; r += 4LL * rgroups) {
      uint2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ldg8(a.x + (r + (long long)u * rgroups) * a.C + c0 + cv * CL_VEC);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        *reinterpret_cast<uint2*>(s.tile + (r + (long long)u * rgroups - r0) * a.CS + cv * CL_VEC) = v[u];
        float f[CL_VEC]; unpack4(v[u], f);
#pragma unroll
        for (int i = 0; i < CL_VEC; ++i) { acc[0][i] += f[i]; acc[1][i] = fmaf(f[i], f[i], acc[1][i]); }
      }
    }
    for (; r < r1; r += rgroups) {
      const uint2 v = ldg8(a.x + r * a.C + c0 + cv * CL_VEC);
      *reinterpret_cast<uint2*>(s.tile + (r - r0) * a.CS + cv * CL_VEC) = v;
      float f[CL_VEC]; unpack4(v, f);
#pragma unroll
      for (int i = 0; i < CL_VEC; ++i) { acc[0][i] += f[i]; acc[1][i] = fmaf(f[i], f[i], acc[1][i]); }
    }
  }
  fold(s, acc, a.CS, rgroups, rg, cv, cluster, a.K);

  const float inv_m = 1.0f / (float)a.M;
  for (int c = threadIdx.x; c < a.CS; c += CL_THREADS) {
    const float m = s.total[c] * inv_m;
    float var = fmaf(-m, m, s.total[a.CS + c] * inv_m);
    var = var < 0.f ? 0.f : var;
    const float istd = rsqrtf(var + a.eps);
    const float sc = a.gamma[c0 + c] * istd;
    s.coef[c] = sc;
    s.coef[a.CS + c] = fmaf(-m, sc, a.beta[c0 + c]);
    if (rank == 0) {
      a.mean[c0 + c] = m;
      a.invstd[c0 + c] = istd;
      if (a.running_mean) {
        const float unbiased = a.M > 1 ? var * ((float)a.M / (float)(a.M - 1)) : var;
        a.running_mean[c0 + c] = fmaf(a.momentum, m - a.running_mean[c0 + c], a.running_mean[c0 + c]);
        a.running_var[c0 + c] = fmaf(a.momentum, unbiased - a.running_var[c0 + c], a.running_var[c0 + c]);
      }
    }
  }
  __syncthreads();
  float scale[CL_VEC], shift[CL_VEC];
#pragma unroll
  for (int i = 0; i < CL_VEC; ++i) { scale[i] = s.coef[cv * CL_VEC + i]; shift[i] = s.coef[a.CS + cv * CL_VEC + i]; }
  const bool has_res = a.res != nullptr;
  for (long long r = r0 + rg; r < r1; r += rgroups) {
    const long long off = r * a.C + c0 + cv * CL_VEC;
    float f[CL_VEC];
    unpack4(*reinterpret_cast<const uint2*>(s.tile + (r - r0) * a.CS + cv * CL_VEC), f);
#pragma unroll
    for (int i = 0; i < CL_VEC; ++i) f[i] = fmaf(f[i], scale[i], shift[i]);
    if (has_res) {
      float g[CL_VEC]; unpack4(ldg8(a.res + off), g);
#pragma unroll
      for (int i = 0; i < CL_VEC; ++i) f[i] += g[i];
    }
    if (a.relu) {
#pragma unroll
      for (int i = 0; i < CL_VEC; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    *reinterpret_cast<uint2*>(a.y + off) = pack4(f);
  }
  cluster.sync();                                             // nobody leaves while a peer may still read its partials
}

__global__ void __launch_bounds__(CL_THREADS, 1) bn_bwd_cluster_kernel(const ClBwdArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int slice = blockIdx.x / a.K;
  const int c0 = slice * a.CS;
  const Smem s = carve(a.CS);
  const int tpr = a.CS / CL_VEC, rgroups = CL_THREADS / tpr;
  const int cv = threadIdx.x % tpr, rg = threadIdx.x / tpr;
  const long long r0 = (long long)rank * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.M) r1 = a.M;
  float mean[CL_VEC], istd[CL_VEC];
#pragma unroll
  for (int i = 0; i < CL_VEC; ++i) { mean[i] = a.mean[c0 + cv * CL_VEC + i]; istd[i] = a.invstd[c0 + cv * CL_VEC + i]; }
  const bool relu = a.relu != 0;

  float acc[2][CL_VEC];
#pragma unroll
  for (int i = 0; i < CL_VEC; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
  // pass 1: d = relu-masked dy -> shared memory (and dres); sums of d and d * xhat
  for (long long r = r0 + rg; r < r1; r += 2LL * rgroups) {
    const long long ra = r, rb = r + rgroups;
    const bool hb = rb < r1;
    const long long offa = ra * a.C + c0 + cv * CL_VEC, offb = rb * a.C + c0 + cv * CL_VEC;
    uint2 dva = ldg8(a.dy + offa), xva = ldg8(a.x + offa), yva = make_uint2(0, 0);
    uint2 dvb = make_uint2(0, 0), xvb = make_uint2(0, 0), yvb = make_uint2(0, 0);
    if (relu) yva = ldg8(a.y + offa);
    if (hb) { dvb = ldg8(a.dy + offb); xvb = ldg8(a.x + offb); if (relu) yvb = ldg8(a.y + offb); }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !hb) break;
      const uint2 dv = h ? dvb : dva, xv = h ? xvb : xva, yv = h ? yvb : yva;
      const long long rr = h ? rb : ra, off = h ? offb : offa;
      float d[CL_VEC], xf[CL_VEC];
      unpack4(dv, d); unpack4(xv, xf);
      if (relu) {
        float yf[CL_VEC]; unpack4(yv, yf);
#pragma unroll
        for (int i = 0; i < CL_VEC; ++i) d[i] = yf[i] > 0.f ? d[i] : 0.f;
      }
      const uint2 dm = pack4(d);                               // exact: every d[i] is a bf16 value or zero
      *reinterpret_cast<uint2*>(s.tile + (rr - r0) * a.CS + cv * CL_VEC) = dm;
      if (a.dres) *reinterpret_cast<uint2*>(a.dres + off) = dm;
#pragma unroll
      for (int i = 0; i < CL_VEC; ++i) { acc[0][i] += d[i]; acc[1][i] = fmaf(d[i], (xf[i] - mean[i]) * istd[i], acc[1][i]); }
    }
  }
  fold(s, acc, a.CS, rgroups, rg, cv, cluster, a.K);

  const float inv_m = 1.0f / (float)a.M;
  for (int c = threadIdx.x; c < a.CS; c += CL_THREADS) {
    const float sd = s.total[c], sq = s.total[a.CS + c];
    if (rank == 0) { a.dbeta[c0 + c] = sd; a.dgamma[c0 + c] = sq; }
    // dx = gs * (d - m1 - xhat * m2) with xhat = (x - mean) * istd   ==   A * d + B * x + Cc
    const float is = a.invstd[c0 + c], gs = a.gamma[c0 + c] * is, m1 = sd * inv_m, m2 = sq * inv_m, mu = a.mean[c0 + c];
    s.coef[c] = gs;
    s.coef[a.CS + c] = -gs * m2 * is;
    s.coef[2 * a.CS + c] = gs * (m2 * is * mu - m1);
  }
  __syncthreads();
  float cA[CL_VEC], cB[CL_VEC], cC[CL_VEC];
#pragma unroll
  for (int i = 0; i < CL_VEC; ++i) {
    cA[i] = s.coef[cv * CL_VEC + i]; cB[i] = s.coef[a.CS + cv * CL_VEC + i]; cC[i] = s.coef[2 * a.CS + cv * CL_VEC + i];
  }
  // pass 2: x again (L2), d from shared memory
  for (long long r = r0 + rg; r < r1; r += rgroups) {
    const long long off = r * a.C + c0 + cv * CL_VEC;
    float d[CL_VEC], xf[CL_VEC], o[CL_VEC];
    unpack4(ldg8(a.x + off), xf);
    unpack4(*reinterpret_cast<const uint2*>(s.tile + (r - r0) * a.CS + cv * CL_VEC), d);
#pragma unroll
    for (int i = 0; i < CL_VEC; ++i) o[i] = fmaf(cA[i], d[i], fmaf(cB[i], xf[i], cC[i]));
    *reinterpret_cast<uint2*>(a.dx + off) = pack4(o);
  }
  cluster.sync();
}

struct Plan { int CS, K, rows_per_cta, smem; };

bool make_plan(long long M, int C, Plan* p) {
  if (M < 16 || C < 32 || (C & (C - 1))) return false;
  // candidates: 64-channel slices first (128-byte row chunks), then 128, then 32; clusters of 8 (portable) or 16 CTAs.
  // Pick the plan with the most CTAs in flight (<= 148); ties go to the earlier candidate.
  const int css[3] = {64, 128, 32};
  const int ks[2] = {8, 16};
  int best = 0;
  for (int ci = 0; ci < 3; ++ci) {
    const int cs = css[ci];
    if (cs > C) continue;
    for (int ki = 0; ki < 2; ++ki) {
      const int K = ks[ki];
      const long long rows = (M + K - 1) / K;
      const long long smem = CL_HEADER_BYTES + rows * cs * 2;
      const int ctas = (C / cs) * K;
      if (smem > 200 * 1024 || ctas > 148) continue;
      if (ctas > best) {
        best = ctas;
        p->CS = cs; p->K = K; p->rows_per_cta = (int)rows; p->smem = (int)smem;
      }
    }
  }
  return best > 0;
}

template <typename KernT, typename ArgT>
int launch_cluster(KernT kern, const ArgT& a, const Plan& p, int C, cudaStream_t stream) {
  static bool configured = false;                             // one instantiation of this template per kernel
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((C / p.CS) * p.K), 1, 1);
  cfg.blockDim = dim3(CL_THREADS, 1, 1);
  cfg.dynamicSmemBytes = (size_t)p.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.K; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kern, a);
}

}  // namespace

// 1 if (M rows, C channels) is served by the cluster kernels; CS / K report the plan (channels per cluster, CTAs per cluster).
extern "C" int drc_bn_cluster_plan(long long M, int C, int* CS, int* K) {
  Plan p;
  if (!make_plan(M, C, &p)) return 0;
  if (CS) *CS = p.CS;
  if (K) *K = p.K;
  return 1;
}

extern "C" int drc_bn_fwd_cluster(const void* x, const void* res, void* y, const float* gamma, const float* beta, float* running_mean,
                                  float* running_var, float* mean, float* invstd, long long M, int C, float eps, float momentum,
                                  int relu, cudaStream_t stream) {
  Plan p;
  if (!make_plan(M, C, &p)) return -1;
  ClFwdArgs a;
  a.x = (const __nv_bfloat16*)x; a.res = (const __nv_bfloat16*)res; a.y = (__nv_bfloat16*)y; a.gamma = gamma; a.beta = beta;
  a.running_mean = running_mean; a.running_var = running_var; a.mean = mean; a.invstd = invstd;
  a.M = M; a.C = C; a.CS = p.CS; a.K = p.K; a.rows_per_cta = p.rows_per_cta; a.eps = eps; a.momentum = momentum; a.relu = relu;
  return launch_cluster(bn_fwd_cluster_kernel, a, p, C, stream);
}

extern "C" int drc_bn_bwd_cluster(const void* dy, const void* y, const void* x, const float* gamma, const float* mean, const float* invstd,
                                  void* dx, void* dres, float* dgamma, float* dbeta, long long M, int C, int relu, cudaStream_t stream) {
  Plan p;
  if (!make_plan(M, C, &p)) return -1;
  ClBwdArgs a;
  a.dy = (const __nv_bfloat16*)dy; a.y = (const __nv_bfloat16*)y; a.x = (const __nv_bfloat16*)x; a.gamma = gamma; a.mean = mean;
  a.invstd = invstd; a.dx = (__nv_bfloat16*)dx; a.dres = (__nv_bfloat16*)dres; a.dgamma = dgamma; a.dbeta = dbeta;
  a.M = M; a.C = C; a.CS = p.CS; a.K = p.K; a.rows_per_cta = p.rows_per_cta; a.relu = relu;
  return launch_cluster(bn_bwd_cluster_kernel, a, p, C, stream);
}
