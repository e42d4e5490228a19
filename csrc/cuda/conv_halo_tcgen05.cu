// K10 (convolution part, continued): HALO-REUSE 3x3 / stride-1 convolution on tcgen05 for the 64 -> 64 channel layers.
//
// profiles/ncu_conv.md: on the layer1 shape (128 x 64 x 32 x 32) the per-tap kernel of conv_tap_tcgen05.cu -- and cuDNN's --
// is bound by L2 -> SM traffic: every filter tap re-loads the shifted activation patch (9 x 16.9 MB = 147 MB at 7.5 TB/s)
// and re-loads its 8 KB weight slice for every tile.  With Cin = Cout = 64 everything a CTA needs fits in shared memory:
//
//   * the nine weight taps (9 x 8 KB) are loaded ONCE per CTA and stay resident (the CTA is persistent over tiles);
//   * an M tile is an 8 x 16 pixel block; its (8+2) x (16+2) HALO patch (180 pixel rows x 128 B = 23 KB) is loaded ONCE per
//     tile by one 4-D TMA box (out-of-image pixels zero-filled);
//   * the A operand of tap (r, s) is that same shared-memory patch addressed through a row-shifted UMMA descriptor: start
//     address + (r*10 + s) * 128 B, stride between 8-row groups = 10 rows = 1280 B (an 8-pixel image row is one 8-row group,
//     the next image row starts 10 patch rows later).  The 128B swizzle is an XOR of ADDRESS bits [4:6] with bits [7:9] and the
//     patch base is 1024-byte aligned, so what TMA wrote and what the MMA unit reads agree for any 128-byte-aligned start
//     (validated on a B200: tests/test_gemm_gpu.py::test_conv3x3_halo_reuse_kernels; the descriptor base-offset field stays 0).
//
// L2 traffic per tile drops from 9 x (16 + 8) KB to 23 KB; the 36 MMAs of a tile (9 taps x 4 K-steps) run back to back on
// one resident patch.  dgrad is the same kernel with mirrored taps and the weight tile read MN-major (same shared-memory image).
// The epilogue is the shared staged TMA-store epilogue of conv_epilogue.cuh, including the fused BatchNorm statistics.
//
// Reference counterpart: the 64 -> 64 BasicBlock convolutions of src/model_ops/resnet.py:14-36.
#include "conv_epilogue.cuh"
#include "tcgen05_common.cuh"

namespace {

using namespace tc;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 64;          // Cout (fprop) / Cin (dgrad)
constexpr int BLOCK_K = 64;          // the 64 reduction channels: ONE K-slice per tap
constexpr int NUM_THREADS = 256;
constexpr int TW = 8, TH = 16;       // tile = 8 x 16 pixels
constexpr int PH = TH + 2, PW = TW + 2;              // halo patch: 18 image rows x 10 pixels
constexpr int PATCH_BYTES = PW * PH * 128;           // 23040
constexpr int PATCH_STRIDE = 23 * 1024;              // stage pitch (1024-byte aligned)
constexpr int W_TAP_BYTES = BLOCK_N * 128;           // 8 KB
constexpr int W_BYTES = 9 * W_TAP_BYTES;             // 72 KB
constexpr int STAGES = 4;
constexpr int EPI_BYTES = convepi::staging_bytes(BLOCK_N) + convepi::stat_bytes();
constexpr int SMEM_BYTES = W_BYTES + STAGES * PATCH_STRIDE + EPI_BYTES + 1024 + 256;

struct HaloArgs {
  int N, H, W;
  const float* bias_f32;
  const __nv_bfloat16* bias_bf16;
  convepi::BnStatArgs stat;
};

template <bool DGRAD>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_halo_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                         const __grid_constant__ CUtensorMap tmap_out, const HaloArgs a) {
  constexpr int TMEM_COLS = 128;                    // 2 accumulators x 64 columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;                              // 9 taps x [64 rows][128 B]
  uint8_t* s_patch = smem + W_BYTES;                // STAGES x PATCH_STRIDE
  uint8_t* sbuf = s_patch + STAGES * PATCH_STRIDE;  // staging tile [128][128 B]
  float* s_stat = reinterpret_cast<float*>(sbuf + convepi::staging_bytes(BLOCK_N));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbuf + EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_bar = tmem_empty + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wt = a.W / TW, ht = a.H / TH;
  const int num_tiles = wt * ht * a.N;
  const bool want_stats = a.stat.partial != nullptr;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
    mbar_init(w_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_base_slot);
  if (warp >= 4) {
    for (int i = threadIdx.x - 128; i < 2 * convepi::STAT_PARTS * convepi::STAT_MAX_C; i += convepi::EPI_THREADS) s_stat[i] = 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer: weights once, then one halo patch per tile =====================
    if (elect_one()) {
      mbar_expect_tx(w_bar, W_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(s_w + tap * W_TAP_BYTES, &tmap_w, tap * 64, 0, w_bar);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int w0 = (tile % wt) * TW, h0 = ((tile / wt) % ht) * TH, n0 = tile / (wt * ht);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], (uint32_t)PATCH_BYTES);
        tma_load_4d(s_patch + stage * PATCH_STRIDE, &tmap_x, 0, w0 - 1, h0 - 1, n0, &full_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 9 taps x 4 K-steps on one resident patch =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, false, DGRAD);
      mbar_wait(w_bar, 0);
      tcgen05_fence_after();
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        const uint32_t patch = smem_u32(s_patch + stage * PATCH_STRIDE);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int r = tap / 3, s = tap - 3 * r;
          // fprop: input pixel (y + r - 1, x + s - 1) = patch row (y + r) * 10 + (x + s); dgrad: (y + 1 - r, x + 1 - s) -> (2 - r, 2 - s)
          const int pr = DGRAD ? 2 - r : r, ps = DGRAD ? 2 - s : s;
          const uint32_t a_addr = patch + (uint32_t)(pr * PW + ps) * 128u;
          const uint32_t b_addr = smem_u32(s_w + tap * W_TAP_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = desc_kmajor(a_addr + k * UMMA_K * 2, PW * 128);
            // weights: K-major rows = Cout (fprop) or MN-major rows = Cout = K (dgrad); 64 rows -> 8 groups of 1024 B
            const uint64_t db = DGRAD ? desc_mnmajor(b_addr + k * UMMA_K * 128, BLOCK_K * 128) : desc_kmajor(b_addr + k * UMMA_K * 2);
            umma_f16(tmem_d, da, db, idesc, (tap | k) ? 1u : 0u);
          }
        }
        tcgen05_commit(&empty_bar[stage]);
        tcgen05_commit(&tmem_full[acc]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: staged TMA store (+ BatchNorm statistics) =====================
    const int et = threadIdx.x - 128;
    int acc = 0; uint32_t acc_phase = 0;
    convepi::StatAcc<BLOCK_N> sacc;
    sacc.clear();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int w0 = (tile % wt) * TW, h0 = ((tile / wt) % ht) * TH, n0 = tile / (wt * ht);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      convepi::drain_tile<BLOCK_N>(tmem_base + (uint32_t)(acc * BLOCK_N), sbuf, s_stat, et, BLOCK_M, 0, BLOCK_N, a.bias_f32, a.bias_bf16,
                                   &tmem_empty[acc]);
      if (et == 0) {
        tma_store_4d(&tmap_out, sbuf, 0, w0, h0, n0);
        tma_store_commit();
      }
      if (want_stats) sacc.add_tile(sbuf, et, BLOCK_M);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (want_stats) sacc.flush(s_stat, et, 0, BLOCK_N);
    if (et == 0) tma_store_wait<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
  if (want_stats)
    convepi::finalize_stats<NUM_THREADS>(a.stat, s_stat, convepi::STAT_PARTS, BLOCK_N, (int)blockIdx.x, (int)gridDim.x, 0,
                                         BLOCK_N, reinterpret_cast<float*>(sbuf));
}

}  // namespace

extern "C" int drc_conv_halo_supported(int H, int W, int Cin, int Cout) {
  return Cin == 64 && Cout == 64 && W % TW == 0 && H % TH == 0 && W >= TW && H >= TH;
}

extern "C" int drc_conv_halo_stat_slots(int N, int H, int W, int num_sms) {
  const int tiles = (W / TW) * (H / TH) * N;
  return tiles < num_sms ? tiles : num_sms;
}

// act: [N,H,W,64] bf16 (x for fprop, dy for dgrad); wgt: [64,3,3,64] bf16 (arena layout); out: [N,H,W,64] bf16.
// stat_*: optional fused BatchNorm statistics of the output (fprop only), see drc_convg.
extern "C" int drc_conv_halo(const void* act, const void* wgt, void* out, int N, int H, int W, int dgrad, const float* bias_f32,
                             const void* bias_bf16, float* stat_partial, unsigned int* stat_counter, float* stat_mean,
                             float* stat_invstd, float* running_mean, float* running_var, float eps, float momentum, int num_sms,
                             int device, cudaStream_t stream) {
  if (!drc_conv_halo_supported(H, W, 64, 64)) return -1;
  if (stat_partial && dgrad) return -4;
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return (int)e; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -2;
  HaloArgs a;
  a.N = N; a.H = H; a.W = W; a.bias_f32 = bias_f32; a.bias_bf16 = (const __nv_bfloat16*)bias_bf16;
  a.stat.partial = stat_partial; a.stat.counter = stat_counter; a.stat.mean = stat_mean; a.stat.invstd = stat_invstd;
  a.stat.running_mean = running_mean; a.stat.running_var = running_var; a.stat.count = (long long)N * H * W;
  a.stat.eps = eps; a.stat.momentum = momentum;
  CUtensorMap tx, tw, tout;
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    cuuint32_t box[4] = {64, (cuuint32_t)PW, PH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(act), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1000 + (int)r;
  }
  int r = encode_mat(&tw, wgt, 64, 9 * 64, 9 * 64, 64);
  if (r) return 2000 + r;
  r = encode_act(&tout, out, 64, W, H, N, TW, TH, 1, 1);
  if (r) return 3000 + r;
  const int tiles = (W / TW) * (H / TH) * N;
  const int grid = tiles < num_sms ? tiles : num_sms;
  static bool cfg0 = false, cfg1 = false;
  if (!dgrad) {
    auto kern = conv_halo_tcgen05_kernel<false>;
    if (!cfg0) { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); if (e != cudaSuccess) return (int)e; cfg0 = true; }
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tx, tw, tout, a);
  } else {
    auto kern = conv_halo_tcgen05_kernel<true>;
    if (!cfg1) { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); if (e != cudaSuccess) return (int)e; cfg1 = true; }
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tx, tw, tout, a);
  }
  return (int)cudaGetLastError();
}
