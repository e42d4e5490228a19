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
  const __nv_bfloat16* resid;     // optional [N,H,W,64] tensor added to the output in the epilogue (dgrad: the other branch's gradient)
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
      if (want_stats && a.stat.bwd_x) {            // BatchNorm-backward mode: this tile's rows of x / y into L2 under the MMAs
        const long long off = (((long long)n0 * a.H + (h0 + et / TW)) * a.W + (w0 + et % TW)) * BLOCK_N;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.stat.bwd_x + off));
        if (a.stat.bwd_mask) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.stat.bwd_mask + off));
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      auto row_off = [&](int r) -> long long { return (((long long)n0 * a.H + (h0 + r / TW)) * a.W + (w0 + r % TW)) * BLOCK_N; };
      const __nv_bfloat16* rrow = a.resid ? a.resid + row_off(et) : nullptr;
      const __nv_bfloat16* mrow = (want_stats && a.stat.bwd_mask) ? a.stat.bwd_mask + row_off(et) : nullptr;
      convepi::drain_tile<BLOCK_N>(tmem_base + (uint32_t)(acc * BLOCK_N), sbuf, s_stat, et, BLOCK_M, 0, BLOCK_N, a.bias_f32, a.bias_bf16,
                                   &tmem_empty[acc], 0u, rrow, mrow);
      if (et == 0) {
        tma_store_4d(&tmap_out, sbuf, 0, w0, h0, n0);
        tma_store_commit();
      }
      if (want_stats) {
        if (a.stat.bwd_x) sacc.add_tile_bwd(sbuf, et, BLOCK_M, a.stat.bwd_x, row_off, 0, BLOCK_N, a.stat.bwd_mean, a.stat.bwd_invstd);
        else sacc.add_tile(sbuf, et, BLOCK_M);
      }
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

// ---------------------------------------------------------------------------------------------------------------------------
// Weight gradient of the same layers: dW[co][tap][ci] = sum_px dy[px][co] * x[px + tap][ci], ALL NINE TAPS FROM ONE RESIDENT PATCH.
//
// The per-tap split-K kernel (conv_tap_tcgen05.cu) re-reads dy and a shifted x patch for every tap and wastes half of every MMA
// (Cout = 64 rows of a 128-row instruction): 48 us against cuDNN's 29 us on the layer1 shape.  Here the roles are swapped and the
// taps are stacked in M:
//   D_g[m = (tap in pair, ci)][n = co]  +=  A_g[px][m] * B[px][n]        K = pixels of an 8 x 16 tile, 16 per MMA
//   A_g : the x HALO patch (the very box the forward kernel loads) read MN-major: K rows = pixels; the two 64-channel M atoms of a
//         tap pair are the SAME patch at two start offsets -- the descriptor's leading-dimension offset (LBO) is the byte distance
//         between the taps, its stride between 8-pixel groups (SBO) the patch pitch of one image row (1280 B);
//   B   : the dy tile [128 px][64 co], MN-major, dense.
// Five accumulators (tap pairs (0,1) (2,3) (4,5) (6,7) (7,8); 320 TMEM columns) live across ALL tiles of the persistent CTA: one
// 39 KB TMA stage and 40 full-rate MMAs per 128 pixels, no epilogue until the end.  Every CTA then writes one fp32 partial
// [64][9][64] and wgrad_halo_reduce_kernel folds the per-CTA partials in a fixed order (bit-deterministic).
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int WG_DY_BYTES = BLOCK_M * 128;                        // 16 KB: [128 px][64 co]
constexpr int WG_STAGE = PATCH_STRIDE + WG_DY_BYTES;               // 39 KB
constexpr int WG_STAGES = 4;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE + 1024 + 256;
constexpr int WG_GROUPS = 5;
constexpr int WG_TMEM_COLS = 512;

struct HaloWgradArgs {
  int N, H, W;
  float* partial;                  // [grid][64 co][9 taps][64 ci]
};

__device__ __forceinline__ int wg_tap(int g, int half) { return g < 4 ? 2 * g + half : 7 + half; }

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_halo_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy, const HaloWgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* tmem_full = empty_bar + WG_STAGES;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wt = a.W / TW, ht = a.H / TH;
  const int num_tiles = wt * ht * a.N;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_dy);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<WG_TMEM_COLS>(tmem_base_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int w0 = (tile % wt) * TW, h0 = ((tile / wt) % ht) * TH, n0 = tile / (wt * ht);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sp = smem + stage * WG_STAGE;
        mbar_expect_tx(&full_bar[stage], (uint32_t)(PATCH_BYTES + WG_DY_BYTES));
        tma_load_4d(sp, &tmap_x, 0, w0 - 1, h0 - 1, n0, &full_bar[stage]);               // x halo patch (zero-filled outside)
        tma_load_4d(sp + PATCH_STRIDE, &tmap_dy, 0, w0, h0, n0, &full_bar[stage]);       // dy tile
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, true, true);                   // both operands MN-major
      int stage = 0; uint32_t phase = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t patch = smem_u32(smem + stage * WG_STAGE);
        const uint32_t dyt = patch + PATCH_STRIDE;
#pragma unroll 1
        for (int ks = 0; ks < BLOCK_M / UMMA_K; ++ks) {                                  // 16 pixels = two image rows of the tile
          const uint64_t db = make_desc(dyt + ks * (UMMA_K * 128), 1024, 1);
#pragma unroll
          for (int g = 0; g < WG_GROUPS; ++g) {
            const int t0 = wg_tap(g, 0), t1 = wg_tap(g, 1);
            const int o0 = (t0 / 3) * PW + (t0 % 3), o1 = (t1 / 3) * PW + (t1 % 3);     // patch pixel offsets of the two taps
            const uint32_t a_addr = patch + (uint32_t)((2 * ks) * PW + o0) * 128u;
            const uint64_t da = make_desc(a_addr, PW * 128, (uint32_t)((o1 - o0) * 128) >> 4);
            umma_f16(tmem_base + g * BLOCK_N, da, db, idesc, (first && ks == 0) ? 0u : 1u);
          }
        }
        first = false;
        tcgen05_commit(&empty_bar[stage]);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      tcgen05_commit(tmem_full);
    }
  } else if (warp >= 4) {
    // ===================== epilogue (once): TMEM -> fp32 partial [co][tap][ci] =====================
    const int q = warp & 3;
    mbar_wait(tmem_full, 0);
    tcgen05_fence_after();
    const int m = q * 32 + lane;                                   // accumulator row = (tap in pair, ci)
    const int half = m >> 6, ci = m & 63;
    float* base = a.partial + (long long)blockIdx.x * (64 * 9 * 64) + ci;
#pragma unroll 1
    for (int g = 0; g < WG_GROUPS; ++g) {
      const int tap = wg_tap(g, half);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BLOCK_N + c), v);
        if (g == 4 && half == 0) continue;                         // tap 7 was already produced by group 3
#pragma unroll
        for (int j = 0; j < 32; ++j) base[((long long)(c + j) * 9 + tap) * 64] = __uint_as_float(v[j]);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<WG_TMEM_COLS>(tmem_base);
}

// out[e] = sum over the per-CTA partials (8 interleaved subsets, each with up to 16 independent 16-byte loads in flight, combined
// in a fixed order), written as bf16.  elems = 64 * 9 * 64; grid = elems / 4 / 32 CTAs of 256 threads.
__global__ void __launch_bounds__(256) wgrad_halo_reduce_kernel(const float* partial, int parts, __nv_bfloat16* out) {
  constexpr int ELEMS4 = 64 * 9 * 64 / 4;
  __shared__ float4 s[8][32];
  const int il = threadIdx.x & 31, sub = threadIdx.x >> 5;
  const int e4 = blockIdx.x * 32 + il;
  const float4* p4 = reinterpret_cast<const float4*>(partial);
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int g0 = sub; g0 < parts; g0 += 8 * 16) {
    float4 v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int g = g0 + 8 * u;
      v[u] = g < parts ? __ldcg(p4 + (long long)g * ELEMS4 + e4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) { t.x += v[u].x; t.y += v[u].y; t.z += v[u].z; t.w += v[u].w; }
  }
  s[sub][il] = t;
  __syncthreads();
  if (sub == 0) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float4 v = s[k][il]; r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w; }
    __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo); o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + (long long)e4 * 4) = o;
  }
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
                             const void* bias_bf16, const void* resid, float* stat_partial, unsigned int* stat_counter, float* stat_mean,
                             float* stat_invstd, float* running_mean, float* running_var, float eps, float momentum,
                             const void* bwd_x, const void* bwd_mask, const float* bwd_mean, const float* bwd_invstd, float* bwd_sums,
                             float* bwd_dgamma, float* bwd_dbeta, int num_sms, int device, cudaStream_t stream) {
  if (!drc_conv_halo_supported(H, W, 64, 64)) return -1;
  if (stat_partial && ((dgrad != 0) != (bwd_x != nullptr))) return -4;
  if (bwd_x && !(stat_partial && bwd_mean && bwd_invstd && bwd_sums && bwd_dgamma && bwd_dbeta)) return -7;
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return (int)e; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -2;
  HaloArgs a;
  a.N = N; a.H = H; a.W = W; a.bias_f32 = bias_f32; a.bias_bf16 = (const __nv_bfloat16*)bias_bf16;
  a.resid = (const __nv_bfloat16*)resid;
  a.stat.partial = stat_partial; a.stat.counter = stat_counter; a.stat.mean = stat_mean; a.stat.invstd = stat_invstd;
  a.stat.running_mean = running_mean; a.stat.running_var = running_var; a.stat.count = (long long)N * H * W;
  a.stat.eps = eps; a.stat.momentum = momentum;
  a.stat.bwd_x = (const __nv_bfloat16*)bwd_x; a.stat.bwd_mask = (const __nv_bfloat16*)bwd_mask; a.stat.bwd_mean = bwd_mean;
  a.stat.bwd_invstd = bwd_invstd; a.stat.bwd_sums = bwd_sums; a.stat.bwd_dgamma = bwd_dgamma; a.stat.bwd_dbeta = bwd_dbeta;
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

// dy: [N,H,W,64] bf16, x: [N,H,W,64] bf16 -> dw: [64,3,3,64] bf16 (arena layout); ws: drc_conv_halo_stat_slots(...) * 36864 floats.
extern "C" int drc_conv_halo_wgrad(const void* dy, const void* x, void* dw, float* ws, int N, int H, int W, int num_sms, int device,
                                   cudaStream_t stream) {
  if (!drc_conv_halo_supported(H, W, 64, 64)) return -1;
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return (int)e; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -2;
  CUtensorMap tx, tdy;
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    cuuint32_t box[4] = {64, (cuuint32_t)PW, PH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1000 + (int)r;
  }
  int r = encode_act(&tdy, dy, 64, W, H, N, TW, TH, 1, 1);
  if (r) return 2000 + r;
  HaloWgradArgs a;
  a.N = N; a.H = H; a.W = W; a.partial = ws;
  const int grid = drc_conv_halo_stat_slots(N, H, W, num_sms);
  static bool cfgd = false;
  if (!cfgd) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_halo_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    if (e != cudaSuccess) return (int)e;
    cfgd = true;
  }
  wgrad_halo_tcgen05_kernel<<<grid, NUM_THREADS, WG_SMEM, stream>>>(tx, tdy, a);
  int rc = (int)cudaGetLastError();
  if (rc) return rc;
  wgrad_halo_reduce_kernel<<<64 * 9 * 64 / 4 / 32, 256, 0, stream>>>(ws, grid, (__nv_bfloat16*)dw);
  return (int)cudaGetLastError();
}
