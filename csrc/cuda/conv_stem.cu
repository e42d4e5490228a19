// K10 (convolution part, continued): the 3-channel stem convolution (3x3 / stride 1 / pad 1, Cin = 3) on CUDA cores.
//
// K = 27 is far too small for the tensor-core path (a UMMA K-step is 16 and a TMA row must be a multiple of 16 bytes; a
// 3-channel bf16 pixel is 6 bytes), and the work is tiny: 0.45 GFLOP per ResNet-18 step at B = 128.  cuDNN serves it with
// legacy kernels that take 35 us (fprop) and 47 us (wgrad) -- 4 % of a 2.0 ms worker step
// (profiles/worker_profile_ResNet18_fused.txt).  Here:
//
//   stem_fprop_kernel : one thread = one output pixel x 16 output channels.  The 27 inputs of the pixel live in registers,
//                       the 27 x 64 weights in shared memory (fp32, read as broadcast float4), the 16 results leave as one
//                       32-byte store -- the kernel is bound by writing the 16.8 MB output.
//   stem_wgrad_kernel : persistent CTAs walk 128-pixel tiles; the dy tile [128][64] and the x halo patch sit in shared
//                       memory, thread (co, g) keeps 7 of the 27 filter entries of output channel co in registers.  Every CTA
//                       writes ONE fp32 partial [64][27]; stem_wgrad_reduce_kernel folds them in a fixed order
//                       (bit-deterministic: the tile -> CTA assignment depends only on the grid size).
//   (no dgrad: the network input needs no gradient)
//
// STATUS: numerics validated on a B200 (tests/test_gemm_gpu.py::test_conv_stem_native_kernels) at the very end of round 1; not
// yet timed against cuDNN, hence still opt-in via DRACO_CONV_STEM=native.
//
// Reference counterpart: `self.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)` of
// src/model_ops/resnet.py:70-72 (PyTorch-0.3 CPU THNN).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int CIN = 3;
constexpr int KTOT = 27;            // 3 x 3 x 3, index k = (r*3 + s)*3 + ci  (the arena layout of a [Cout,3,3,3] weight row)
constexpr int COUT = 64;

struct StemArgs {
  const __nv_bfloat16* x;          // [N, H, W, 3]
  const __nv_bfloat16* w;          // [64, 3, 3, 3]
  __nv_bfloat16* y;                // [N, H, W, 64]
  const float* bias;               // [64] or null
  int N, H, W;
};

__global__ void __launch_bounds__(256) stem_fprop_kernel(const StemArgs a) {
  __shared__ __align__(16) float ws[KTOT][COUT];
  for (int i = threadIdx.x; i < KTOT * COUT; i += blockDim.x) {
    const int co = i / KTOT, k = i - co * KTOT;
    ws[k][co] = __bfloat162float(a.w[i]);
  }
  __syncthreads();
  const long long npix = (long long)a.N * a.H * a.W;
  const long long p = (long long)blockIdx.x * 64 + (threadIdx.x >> 2);
  const int cg = threadIdx.x & 3;                                  // 16-channel group
  if (p >= npix) return;
  const int pw = (int)(p % a.W), ph = (int)((p / a.W) % a.H);
  const long long pn = p / ((long long)a.W * a.H);
  float in[KTOT];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int hh = ph + r - 1, wv = pw + s - 1;
      const bool ok = hh >= 0 && hh < a.H && wv >= 0 && wv < a.W;
      const __nv_bfloat16* src = a.x + ((pn * a.H + hh) * a.W + wv) * CIN;
#pragma unroll
      for (int c = 0; c < CIN; ++c) in[(r * 3 + s) * 3 + c] = ok ? __bfloat162float(src[c]) : 0.f;
    }
  }
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = a.bias ? a.bias[cg * 16 + j] : 0.f;
#pragma unroll
  for (int k = 0; k < KTOT; ++k) {
    const float4* wr = reinterpret_cast<const float4*>(&ws[k][cg * 16]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 wv = wr[q];
      acc[4 * q + 0] = fmaf(in[k], wv.x, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(in[k], wv.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(in[k], wv.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(in[k], wv.w, acc[4 * q + 3]);
    }
  }
  uint32_t o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __nv_bfloat162 v = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
    o[j] = *reinterpret_cast<uint32_t*>(&v);
  }
  uint4* dst = reinterpret_cast<uint4*>(a.y + p * COUT + cg * 16);
  dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
  dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

struct StemWgradArgs {
  const __nv_bfloat16* dy;         // [N, H, W, 64]
  const __nv_bfloat16* x;          // [N, H, W, 3]
  float* partial;                  // [grid][64][27]
  int N, H, W;
  int rows_per_tile;               // rows_per_tile * W == 128
  int tiles;                       // N * H / rows_per_tile
};

constexpr int TILE_PX = 128;
constexpr int MAX_PATCH = 1184;                        // max over W in {4..128} of (128/W + 2) * (W + 2) * 3 = 1170 floats (W = 128)

__global__ void __launch_bounds__(256) stem_wgrad_kernel(const StemWgradArgs a) {
  __shared__ __align__(16) __nv_bfloat16 sdy[TILE_PX][COUT];                 // 16 KB
  __shared__ float sx[MAX_PATCH];                                             // (rows + 2) x (W + 2) x 3 halo patch
  const int co = threadIdx.x & 63, g = threadIdx.x >> 6;                      // g: which 7 of the 27 filter entries
  const int W = a.W, RT = a.rows_per_tile, PWD = W + 2;
  int koff[7];
  bool kval[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int k = g * 7 + j;
    kval[j] = k < KTOT;
    const int kk = kval[j] ? k : 0;
    const int tap = kk / 3, ci = kk - 3 * tap, r = tap / 3, s = tap - 3 * r;
    koff[j] = (r * PWD + s) * CIN + ci;
  }
  float acc[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) acc[j] = 0.f;

  for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
    const int tiles_per_img = a.H / RT;
    const int n = tile / tiles_per_img, h0 = (tile - n * tiles_per_img) * RT;
    __syncthreads();                                                          // previous tile's smem no longer read
    // dy tile: 128 pixels x 64 channels = 16 KB, 16-byte loads
    {
      const uint4* src = reinterpret_cast<const uint4*>(a.dy + (((long long)n * a.H + h0) * W) * COUT);
      uint4* dst = reinterpret_cast<uint4*>(&sdy[0][0]);
      for (int i = threadIdx.x; i < TILE_PX * COUT / 8; i += blockDim.x) dst[i] = src[i];
    }
    // x halo patch (rows h0-1 .. h0+RT, cols -1 .. W), zero outside the image
    for (int i = threadIdx.x; i < (RT + 2) * PWD * CIN; i += blockDim.x) {
      const int c = i % CIN, col = (i / CIN) % PWD - 1, row = i / (CIN * PWD) - 1 + h0;
      float v = 0.f;
      if (row >= 0 && row < a.H && col >= 0 && col < W) v = __bfloat162float(a.x[(((long long)n * a.H + row) * W + col) * CIN + c]);
      sx[i] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int px = 0; px < TILE_PX; ++px) {
      const float d = __bfloat162float(sdy[px][co]);
      const int base = ((px / W) * PWD + (px % W)) * CIN;                     // patch coordinates of the (r=0, s=0) neighbour
#pragma unroll
      for (int j = 0; j < 7; ++j) acc[j] = fmaf(d, sx[base + koff[j]], acc[j]);
    }
  }
  float* out = a.partial + ((long long)blockIdx.x * COUT + co) * KTOT;
#pragma unroll
  for (int j = 0; j < 7; ++j)
    if (kval[j]) out[g * 7 + j] = acc[j];
}

__global__ void stem_wgrad_reduce_kernel(const float* partial, int parts, __nv_bfloat16* dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= COUT * KTOT) return;
  float s = 0.f;
  for (int p = 0; p < parts; ++p) s += partial[(long long)p * COUT * KTOT + i];      // fixed order
  dw[i] = __float2bfloat16_rn(s);
}

bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

extern "C" int drc_conv_stem_supported(int H, int W, int Cin, int Cout) {
  return Cin == 3 && Cout == 64 && pow2(W) && W >= 4 && W <= 128 && (128 % W) == 0 && H % (128 / W) == 0;
}

// grid of the wgrad kernel (= number of fp32 partials [64][27] in the workspace)
extern "C" int drc_conv_stem_wgrad_parts(int N, int H, int W, int num_sms) {
  const int tiles = N * (H / (128 / W));
  const int grid = 2 * num_sms;
  return tiles < grid ? tiles : grid;
}

extern "C" int drc_conv_stem_fprop(const void* x, const void* w, void* y, const float* bias, int N, int H, int W, int device,
                                   cudaStream_t stream) {
  if (!drc_conv_stem_supported(H, W, 3, 64)) return -1;
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return (int)e; }
  StemArgs a;
  a.x = (const __nv_bfloat16*)x; a.w = (const __nv_bfloat16*)w; a.y = (__nv_bfloat16*)y; a.bias = bias;
  a.N = N; a.H = H; a.W = W;
  const long long npix = (long long)N * H * W;
  stem_fprop_kernel<<<(unsigned)((npix + 63) / 64), 256, 0, stream>>>(a);
  return (int)cudaGetLastError();
}

// ws: fp32 workspace of drc_conv_stem_wgrad_parts(...) * 64 * 27 elements; dw: [64,3,3,3] bf16.
extern "C" int drc_conv_stem_wgrad(const void* dy, const void* x, void* dw, float* ws, int N, int H, int W, int num_sms, int device,
                                   cudaStream_t stream) {
  if (!drc_conv_stem_supported(H, W, 3, 64)) return -1;
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return (int)e; }
  StemWgradArgs a;
  a.dy = (const __nv_bfloat16*)dy; a.x = (const __nv_bfloat16*)x; a.partial = ws;
  a.N = N; a.H = H; a.W = W; a.rows_per_tile = 128 / W; a.tiles = N * (H / a.rows_per_tile);
  const int parts = drc_conv_stem_wgrad_parts(N, H, W, num_sms);
  stem_wgrad_kernel<<<parts, 256, 0, stream>>>(a);
  int rc = (int)cudaGetLastError();
  if (rc) return rc;
  stem_wgrad_reduce_kernel<<<(COUT * KTOT + 255) / 256, 256, 0, stream>>>(ws, parts, (__nv_bfloat16*)dw);
  return (int)cudaGetLastError();
}
