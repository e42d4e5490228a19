// Classifier-head helpers of the ResNet family (reference: F.avg_pool2d(out, 4) + nn.Linear in src/model_ops/resnet.py:96-105).
//
//   * global average pool over channels-last activations, forward and backward: x [N, HW, C] bf16 <-> y [N, C] bf16.  ATen's
//     avg_pool2d backward writes the 2 MB gradient of the last ResNet-18 block in ~16 us; one 16-byte store per thread does it
//     at copy speed.
//   * head_prep: the backward of a narrow Linear (10 classes) needs dy zero-padded to a 16-byte row (TMA) and the bias gradient;
//     one small CTA produces both (fixed summation order -> bit-identical on every replica) instead of fill + copy + reduce.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

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

// one thread per (image, 8 channels): sums its HW pixels in order
__global__ void gap_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int HW, int C, float inv) {
  const int cv = C / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * cv) return;
  const int n = i / cv, c = (i - n * cv) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const __nv_bfloat16* p = x + ((long long)n * HW) * C + c;
  for (int s = 0; s < HW; ++s) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(p + (long long)s * C), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] *= inv;
  *reinterpret_cast<uint4*>(y + (long long)n * C + c) = pack8(acc);
}

// one thread per (image, pixel, 8 channels)
__global__ void gap_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, long long total, int HW, int C,
                               float inv) {
  const int cv = C / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % cv);
  const long long n = i / ((long long)cv * HW);
  float f[8];
  unpack8(*reinterpret_cast<const uint4*>(dy + n * C + c8 * 8), f);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] *= inv;
  *reinterpret_cast<uint4*>(dx + i * 8) = pack8(f);
}

// dy [B, n] -> dyp [B, np] (zero padded), db[j] = sum_b dy[b, j] (fp32 accumulation in row order, written in the dtype of the bias)
__global__ void head_prep_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dyp, void* db, int db_bf16,
                                 int B, int n, int np) {
  for (int i = threadIdx.x; i < B * np; i += blockDim.x) {
    const int b = i / np, j = i - b * np;
    dyp[i] = j < n ? dy[b * n + j] : __float2bfloat16_rn(0.f);
  }
  if (db && threadIdx.x < n) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += __bfloat162float(dy[b * n + threadIdx.x]);
    if (db_bf16) ((__nv_bfloat16*)db)[threadIdx.x] = __float2bfloat16_rn(s);
    else ((float*)db)[threadIdx.x] = s;
  }
}

}  // namespace

extern "C" int drc_gap_fwd(const void* x, void* y, int N, int HW, int C, cudaStream_t stream) {
  if (C % 8 || N <= 0) return -1;
  const int total = N * (C / 8);
  gap_fwd_kernel<<<(total + 127) / 128, 128, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, HW, C, 1.0f / (float)HW);
  return (int)cudaGetLastError();
}

extern "C" int drc_gap_bwd(const void* dy, void* dx, int N, int HW, int C, cudaStream_t stream) {
  if (C % 8 || N <= 0) return -1;
  const long long total = (long long)N * HW * (C / 8);
  gap_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, total, HW, C,
                                                                       1.0f / (float)HW);
  return (int)cudaGetLastError();
}

extern "C" int drc_head_prep(const void* dy, void* dyp, void* db, int db_bf16, int B, int n, int np, cudaStream_t stream) {
  if (n > 1024 || np < n) return -1;
  head_prep_kernel<<<1, 1024, 0, stream>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dyp, db, db_bf16, B, n, np);
  return (int)cudaGetLastError();
}
