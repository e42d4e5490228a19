// Fused input preparation: uint8 NCHW image batch -> normalised NHWC tensor in the compute dtype, one launch.
//
// Replaces the chain `x.float().div_(255).sub_(mean).div_(std).to(bf16).contiguous(channels_last)` (six elementwise
// launches over the batch per worker per sub-batch) with one pass: every thread produces one output pixel (all C channels,
// C <= 4), reading the C planes of the uint8 source and writing C consecutive elements of the NHWC destination.
// Reference counterpart: the ToTensor + Normalize transforms of src/util.py:28-52.
//
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

struct PrepArgs {
  const uint8_t* src;              // [N, C, H, W] uint8
  void* dst;                       // [N, H, W, C] bf16 or fp32 (the channels-last storage of an [N, C, H, W] tensor)
  int out_bf16;
  float scale[4], shift[4];        // y = x * scale[c] + shift[c]   (scale = 1 / (255 std), shift = -mean / std)
  long long npix;                  // N * H * W
  int C, HW;
};

__global__ void prep_input_kernel(const PrepArgs a) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.npix) return;
  const long long n = p / a.HW, hw = p - n * a.HW;
  const uint8_t* s = a.src + n * a.C * a.HW + hw;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c < a.C) {
      const float v = fmaf((float)s[(long long)c * a.HW], a.scale[c], a.shift[c]);
      if (a.out_bf16) reinterpret_cast<__nv_bfloat16*>(a.dst)[p * a.C + c] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(a.dst)[p * a.C + c] = v;
    }
  }
}

}  // namespace

extern "C" int drc_prep_input(const void* src, void* dst, int out_bf16, const float* mean, const float* std_, int N, int C, int H,
                              int W, cudaStream_t stream) {
  if (C < 1 || C > 4) return (int)cudaErrorInvalidValue;
  PrepArgs a;
  a.src = (const uint8_t*)src; a.dst = dst; a.out_bf16 = out_bf16; a.C = C; a.HW = H * W; a.npix = (long long)N * H * W;
  for (int c = 0; c < 4; ++c) {
    a.scale[c] = c < C ? 1.0f / (255.0f * std_[c]) : 0.f;
    a.shift[c] = c < C ? -mean[c] / std_[c] : 0.f;
  }
  prep_input_kernel<<<(unsigned)((a.npix + 255) / 256), 256, 0, stream>>>(a);
  return (int)cudaGetLastError();
}
