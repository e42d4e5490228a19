// Replica-deterministic, CUDA-graph-safe dropout for the VGG classifier (reference: nn.Dropout() in src/model_ops/vgg.py:24-31).
//
// Under the repetition and cyclic codes every holder of a batch must apply the SAME dropout mask or honest gradients stop
// being bit-identical and the exact-equality vote / the Fourier syndrome break.  torch's nn.Dropout draws from the device's
// Philox stream, which inside a captured graph advances per replay and per call site -- different for every worker.  Here the
// mask is a pure function of (job seed, STEP READ FROM DEVICE MEMORY, batch identity, layer salt, element index): a counter-based
// hash (two rounds of a 64-bit mix, SplitMix64 finaliser) thresholded at p * 2^32.  Graph replays advance by themselves because
// the step lives in device memory; forward saves nothing -- backward recomputes the mask from the same key.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct DropArgs {
  const void* x;                  // input (forward) or upstream gradient (backward)
  void* y;
  const long long* step;          // device step counter
  long long n;
  unsigned long long key;         // mix of (seed, batch id, layer salt)
  unsigned int threshold;         // drop when hash32 < threshold  (p * 2^32)
  float scale;                    // 1 / (1 - p)
  int is_bf16;
};

__global__ void dropout_kernel(const DropArgs a) {
  const uint64_t base = mix64(a.key ^ mix64((uint64_t)*a.step));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int h = (unsigned int)(mix64(base + (uint64_t)i) >> 32);
    const float m = h < a.threshold ? 0.f : a.scale;
    if (a.is_bf16) {
      const __nv_bfloat16* x = (const __nv_bfloat16*)a.x;
      ((__nv_bfloat16*)a.y)[i] = __float2bfloat16_rn(__bfloat162float(x[i]) * m);
    } else {
      ((float*)a.y)[i] = ((const float*)a.x)[i] * m;
    }
  }
}

}  // namespace

extern "C" int drc_dropout(const void* x, void* y, const long long* step, long long n, unsigned long long key, float p, int is_bf16,
                           cudaStream_t stream) {
  if (n <= 0) return 0;
  DropArgs a;
  a.x = x; a.y = y; a.step = step; a.n = n; a.key = key; a.is_bf16 = is_bf16;
  double t = (double)p * 4294967296.0;
  a.threshold = t >= 4294967295.0 ? 4294967295u : (unsigned int)t;
  a.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  long long blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  dropout_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a);
  return (int)cudaGetLastError();
}
