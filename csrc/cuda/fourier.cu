// K4: PS-side decode of the cyclic (Fourier) gradient code.
//
// Reference flow per parameter tensor (src/master/cyclic_master.py:152-173 + src/c_coding.cpp): E = R f
// (numpy complex128 dot over d), locator solve in C++/Eigen, healthy-row pick, scipy lsq_linear for the
// recombination vector v, then decoded = v^T R and Re(.)/n.  The two passes over the n x d matrix R are the
// heavy part; everything else is O(n^3) per tensor.
//
// Device formulation: (a) `cyclic_project_kernel`: one streaming pass over the complex64 arena R[n][D]
// producing E[T][n] with fp64 accumulation (segmented by tensor through the tile table);
// (b) `cyclic_locate_kernel`: one thread per tensor runs the shared fp64 locator core
// (csrc/common/locator_core.h) and emits v[T][n] as complex64; (c) the recombination Re(v^T R)/n is fused
// with SGD + broadcast in aggregate_update.cu (mode 1).  No host round trip between (a), (b) and (c).
#include "common.cuh"
#include "../common/locator_core.h"

struct ProjectArgs {
  const float* R;                 // [n][2*slot_stride] complex64 interleaved
  long long slot_stride;          // complex elements per worker slot
  int n;
  const float* f;                 // [D] random projection factors (N(1,1), fixed at build time)
  TileView tv;
  double* E;                      // [T][n][2] (re, im), must be zero on entry
};

__global__ void __launch_bounds__(DRC_THREADS) cyclic_project_kernel(const __grid_constant__ ProjectArgs a) {
  __shared__ double s_part[DRC_THREADS / 32][2];
  for (int tile = blockIdx.x; tile < a.tv.ntiles; tile += gridDim.x) {
    int tensor;
    const int valid = tile_valid(a.tv, tile, tensor);
    const long long idx = (long long)tile * DRC_TILE + threadIdx.x * 4;
    const bool active = (int)threadIdx.x * 4 < valid;
    float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) f4 = *reinterpret_cast<const float4*>(a.f + idx);
    for (int i = 0; i < a.n; ++i) {
      double re = 0.0, im = 0.0;
      if (active) {
        const float4* src = reinterpret_cast<const float4*>(a.R + 2 * (i * a.slot_stride + idx));
        float4 c0 = ld_f4(src), c1 = ld_f4(src + 1);
        re = (double)c0.x * f4.x + (double)c0.z * f4.y + (double)c1.x * f4.z + (double)c1.z * f4.w;
        im = (double)c0.y * f4.x + (double)c0.w * f4.y + (double)c1.y * f4.z + (double)c1.w * f4.w;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
      }
      if ((threadIdx.x & 31) == 0) { s_part[threadIdx.x >> 5][0] = re; s_part[threadIdx.x >> 5][1] = im; }
      __syncthreads();
      if (threadIdx.x == 0) {
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int w = 0; w < DRC_THREADS / 32; ++w) { sr += s_part[w][0]; si += s_part[w][1]; }
        atomicAdd(&a.E[((long long)tensor * a.n + i) * 2 + 0], sr);
        atomicAdd(&a.E[((long long)tensor * a.n + i) * 2 + 1], si);
      }
      __syncthreads();
    }
  }
}

extern "C" int drc_cyclic_project(const ProjectArgs* args, int grid, cudaStream_t stream) {
  cyclic_project_kernel<<<grid, DRC_THREADS, 0, stream>>>(*args);
  return (int)cudaGetLastError();
}

struct LocateArgs {
  double* E;                      // [T][n][2]; zeroed after use so the next step can accumulate again
  int T, n, s;
  double rel_tol;
  float2* recomb;                 // [T][n] out
  unsigned int* healthy;          // [T] out bitmask of rows used
  int* flagged;                   // [T] out number of rows flagged Byzantine
};

__global__ void cyclic_locate_kernel(const __grid_constant__ LocateArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.T) return;
  cplx E[DRC_LOC_MAX_N], v[DRC_LOC_MAX_N];
  for (int i = 0; i < a.n; ++i) {
    E[i] = c_make(a.E[((long long)t * a.n + i) * 2], a.E[((long long)t * a.n + i) * 2 + 1]);
    a.E[((long long)t * a.n + i) * 2] = 0.0;
    a.E[((long long)t * a.n + i) * 2 + 1] = 0.0;
  }
  unsigned int mask = 0u;
  int fl = locate_and_recombine(E, a.n, a.s, a.rel_tol, v, &mask);
  for (int i = 0; i < a.n; ++i) a.recomb[(long long)t * a.n + i] = make_float2((float)v[i].re, (float)v[i].im);
  if (a.healthy) a.healthy[t] = mask;
  if (a.flagged) a.flagged[t] = fl;
}

extern "C" int drc_cyclic_locate(const LocateArgs* args, cudaStream_t stream) {
  if (args->n > DRC_LOC_MAX_N || args->s > DRC_LOC_MAX_S) return (int)cudaErrorInvalidValue;
  cyclic_locate_kernel<<<(args->T + 31) / 32, 32, 0, stream>>>(*args);
  return (int)cudaGetLastError();
}

extern "C" int drc_sizeof_ProjectArgs() { return (int)sizeof(ProjectArgs); }
extern "C" int drc_sizeof_LocateArgs() { return (int)sizeof(LocateArgs); }
