// Micro-benchmark: cost of a K block (4 x tcgen05.mma 128 x N x 16, both operands in shared memory) as seen by the issuing
// thread + tensor pipe, without any TMA traffic: (a) back-to-back issue, one commit at the end; (b) + one tcgen05.commit per
// K block (what a pipelined kernel does to release the stage); (c) + one mbarrier try_wait per K block on an already completed
// barrier (the full-barrier wait).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I csrc/cuda -o umma_issue umma_issue.cu -lcuda
#include <cstdio>
#include "tcgen05_common.cuh"
using namespace tc;

template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  for (int i = threadIdx.x; i < (128 * 64 * 2 + N * 64 * 2) * 4 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    mbar_arrive(&bar[1]);                                             // bar[1]: phase 0 complete (for the try_wait variant)
    const uint32_t idesc = make_idesc(128, N, false, false);
    constexpr int STAGE = 128 * 64 * 2 + N * 64 * 2;
    const long long t0 = clock64();
    int stage = 0;
    for (int it = 0; it < iters; ++it) {
      if (mode >= 2) { mbar_wait(&bar[1], 0); tcgen05_fence_after(); }
      const uint32_t sa = smem_u32(smem + stage * STAGE), sb = sa + 128 * 64 * 2;
      const uint64_t da = desc_kmajor(sa), db = desc_kmajor(sb);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16(tmem, da + (uint64_t)((kk * 32) >> 4), db + (uint64_t)((kk * 32) >> 4), idesc, (it | kk) ? 1u : 0u);
      if (mode >= 1) tcgen05_commit(&bar[2]);                         // nobody waits on bar[2]; its arrivals just complete phases
      if (++stage == 4) stage = 0;
    }
    const long long t1 = clock64();
    tcgen05_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int N>
void run(long long* d, int mode) {
  const int iters = 2000;
  const int smem = (128 * 64 * 2 + N * 64 * 2) * 4 + 2048;
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N><<<148, 128, smem>>>(d, iters, mode);
  k<N><<<148, 128, smem>>>(d, iters, mode);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d mode=%d : issue loop %6.1f clk / K block, until complete %6.1f clk / K block (floor %d)  %s\n", N, mode,
         (double)h[0] / iters, (double)h[1] / iters, 4 * N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  for (int mode = 0; mode < 3; ++mode) { run<64>(d, mode); run<128>(d, mode); run<256>(d, mode); }
  return 0;
}
