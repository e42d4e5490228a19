// K10 (convolution part, continued): generalised tap-table implicit-GEMM convolutions on tcgen05 -- stride 1 or 2, 1x1 or 3x3.
//
// Same machinery as conv_tcgen05.cu (TMA patch loads into 128B-swizzled tiles, UMMA 128xNx16 with TMEM double buffering,
// pixel-mapped epilogue); what is new is how a filter tap reaches the tensor core:
//
//   * strided fprop : the activation tensor map carries elementStrides {1, s, s, 1}, so ONE box load delivers the
//                     BW x BH x BN *output* pixels' inputs for a tap (every s-th pixel, halo zero-filled);
//   * strided dgrad : dx is split into its s x s parity classes; for class (ph, pw) only the taps with matching parity
//                     contribute and they read dy with unit stride -- 1 + 2 + 2 + 4 = 9 taps over the four launches of a
//                     3x3 / stride-2 layer (no zero-insertion, no wasted MMAs); the epilogue writes with pixel stride s;
//   * strided wgrad : K-blocks are 64-pixel patches of dy; the matching x patch is a strided box shifted by the tap.
//
// A launch is described by a tap table (input offset + weight column per tap), so the stride-1 3x3 case is the
// 9-tap instance of the same kernel.  Motivation (profiles/worker_profile_ResNet18_fused.txt): the three stride-2 dgrads
// of ResNet-18 cost 220 us of a 2.0 ms step in cuDNN (93 us each for the two large ones) -- 5-10x their FLOP time.
//
// STATUS: numerics validated on a B200 (tests/test_gemm_gpu.py::test_convg_*: fprop / dgrad / wgrad for stride 1 and 2, 1x1 and
// 3x3, plus the autograd path) at the very end of round 1; not yet timed against cuDNN, hence still opt-in via
// DRACO_CONV_STRIDED=tcgen05 (round-2 first item: `bash tools/gpu_ci.sh experimental worker_native`).
//
// Reference counterpart: the strided nn.Conv2d layers of src/model_ops/resnet.py:14-64 (downsampling blocks + shortcuts).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 256;
constexpr int MAX_TAPS = 9;

struct TapConvArgs {
  int N, OH, OW;               // iteration space: one GEMM row per (n, i, j), tiled in BW x BH x BN patches
  int Cred, Cn;                // reduction channels per tap / output channels
  int BW, BH, BN;
  int in_mul;                  // input coordinate = patch origin * in_mul + tap offset   (stride for fprop, 1 for dgrad)
  int ntaps;
  int tap_dw[MAX_TAPS], tap_dh[MAX_TAPS];
  int tap_wcol[MAX_TAPS];      // column of the tap inside a weight row (tap index * Cin)
  int out_H, out_W;            // output tensor geometry
  int out_mul, out_oh, out_ow; // output pixel = (i * out_mul + out_oh, j * out_mul + out_ow)
  __nv_bfloat16* out;          // [N, out_H, out_W, Cn]
  const float* bias_f32;
  const __nv_bfloat16* bias_bf16;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 r;\n\t"
      "elect.sync r|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  const uint64_t sbo = 1024 >> 4;
  const uint64_t lbo = MN_MAJOR ? ((BLOCK_K * 128) >> 4) : 1;
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= lbo << 16;
  d |= sbo << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

template <int BLOCK_N, bool B_MN>
__device__ __forceinline__ uint32_t make_idesc() {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (B_MN ? 1u : 0u) << 16;
  d |= (uint32_t)(BLOCK_N >> 3) << 17;
  d |= (uint32_t)(BLOCK_M >> 4) << 24;
  return d;
}

template <int BLOCK_N, int STAGES, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
convg_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const TapConvArgs a) {
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wt = a.OW / a.BW, ht = a.OH / a.BH, nt = (a.N + a.BN - 1) / a.BN;
  const int m_tiles = wt * ht * nt;
  const int n_tiles = (a.Cn + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int c_blocks = a.Cred / BLOCK_K;           // 64-channel slices per tap
  const int k_blocks = a.ntaps * c_blocks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / n_tiles, n0 = (tile % n_tiles) * BLOCK_N;
        const int w0 = (mt % wt) * a.BW, h0 = ((mt / wt) % ht) * a.BH, nb0 = (mt / (wt * ht)) * a.BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          const int tap = kb / c_blocks, c0 = (kb - tap * c_blocks) * BLOCK_K;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          // activation patch of this tap: (strided) 4-D box, out-of-range pixels zero-filled by TMA
          tma_load_4d(sa, &tmap_x, c0, w0 * a.in_mul + a.tap_dw[tap], h0 * a.in_mul + a.tap_dh[tap], nb0, &full_bar[stage]);
          if (!B_MN) {
            tma_load_2d(sb, &tmap_w, a.tap_wcol[tap] + c0, n0, &full_bar[stage]);                 // rows = Cout tile, K-major
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)                                               // rows = Cout (K), cols = Cin (N)
              tma_load_2d(sb + j * (BLOCK_K * 128), &tmap_w, a.tap_wcol[tap] + n0 + 64 * j, c0, &full_bar[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc<BLOCK_N, B_MN>();
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint64_t da = make_smem_desc<false>(sa), db = make_smem_desc<B_MN>(sb);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adv_a = (uint64_t)((k * UMMA_K * 2) >> 4);
            const uint64_t adv_b = (uint64_t)((B_MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
            umma_f16(tmem_d, da + adv_a, db + adv_b, idesc, (kb | k) ? 1u : 0u);
          }
          tcgen05_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / n_tiles, n0 = (tile % n_tiles) * BLOCK_N;
      const int w0 = (mt % wt) * a.BW, h0 = ((mt / wt) % ht) * a.BH, nb0 = (mt / (wt * ht)) * a.BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const int m = q * 32 + lane;                             // row of the tile = pixel of the patch (w fastest)
      const int pw = w0 + m % a.BW, ph = h0 + (m / a.BW) % a.BH, pn = nb0 + m / (a.BW * a.BH);
      const bool row_ok = pn < a.N;
      __nv_bfloat16* orow = a.out + (((long long)pn * a.out_H + (ph * a.out_mul + a.out_oh)) * a.out_W + (pw * a.out_mul + a.out_ow)) * a.Cn;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c), v);
        const int col0 = n0 + c;
        if (row_ok && col0 < a.Cn) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (a.bias_f32 || a.bias_bf16) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < a.Cn) f[j] += a.bias_f32 ? a.bias_f32[col0 + j] : __bfloat162float(a.bias_bf16[col0 + j]);
          }
          __nv_bfloat16* dst = orow + col0;
          if (col0 + 32 <= a.Cn) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              __nv_bfloat162 p0 = __floats2bfloat162_rn(f[j], f[j + 1]), p1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), p3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
              uint4 o;
              o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
              o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
              *reinterpret_cast<uint4*>(dst + j) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < a.Cn) dst[j] = __float2bfloat16_rn(f[j]);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}



struct WgradGArgs {
  int N, H, W, Cin, Cout;      // H, W: spatial size of dy (the K-block patch lives in dy space)
  int ks, stride, pad;         // filter size (1 or 3), stride, padding of the forward convolution
  int PW, PH, PN;               // 64-pixel patch shape
  int splits;                   // K splits
  float* partial;               // [splits][Cout][9*Cin]
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
convg_wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgradGArgs a) {
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;            // 2 atoms of [64 px][64 co]
  constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wt = a.W / a.PW, ht = a.H / a.PH, nt = (a.N + a.PN - 1) / a.PN;
  const int k_total = wt * ht * nt;                          // 64-pixel K-blocks
  const int m_tiles = (a.Cout + BLOCK_M - 1) / BLOCK_M, n_tiles = (a.Cin + BLOCK_N - 1) / BLOCK_N;
  const int ntaps = a.ks * a.ks;
  const int out_tiles = m_tiles * ntaps * n_tiles;
  const int num_work = out_tiles * a.splits;
  const int k_per_split = (k_total + a.splits - 1) / a.splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // work item -> (split, m tile, tap, n tile); splits of one output tile are spread over different CTAs
  auto decode = [&](int wi, int& split, int& co0, int& tap, int& ci0) {
    split = wi / out_tiles;
    int t = wi - split * out_tiles;
    const int ntile = t % n_tiles; t /= n_tiles;
    tap = t % ntaps; t /= ntaps;
    co0 = t * BLOCK_M; ci0 = ntile * BLOCK_N;
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
        int split, co0, tap, ci0; decode(wi, split, co0, tap, ci0);
        const int r = tap / a.ks, s = tap - a.ks * r;
        const int kb0 = split * k_per_split, kb1 = min(k_total, kb0 + k_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int w0 = (kb % wt) * a.PW, h0 = ((kb / wt) % ht) * a.PH, n0 = (kb / (wt * ht)) * a.PN;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          // a 64-channel atom entirely beyond Cout is never read back: skip its load (its TMEM rows hold junk)
          const int a_atoms = min(BLOCK_M / 64, (a.Cout - co0 + 63) / 64);
          mbar_expect_tx(&full_bar[stage], a_atoms * (BLOCK_K * 128) + B_BYTES);
          for (int j = 0; j < a_atoms; ++j) tma_load_4d(sa + j * (BLOCK_K * 128), &tmap_dy, co0 + 64 * j, w0, h0, n0, &full_bar[stage]);
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j) tma_load_4d(sb + j * (BLOCK_K * 128), &tmap_x, ci0 + 64 * j, w0 * a.stride + s - a.pad, h0 * a.stride + r - a.pad, n0, &full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      uint32_t idesc = make_idesc<BLOCK_N, true>();
      idesc |= 1u << 15;                                      // A is MN-major too
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
        int split, co0, tap, ci0; decode(wi, split, co0, tap, ci0);
        const int kb0 = split * k_per_split, kb1 = min(k_total, kb0 + k_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint64_t da = make_smem_desc<true>(sa), db = make_smem_desc<true>(sb);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 128) >> 4);
            umma_f16(tmem_d, da + adv, db + adv, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tcgen05_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    const long long row_pitch = (long long)ntaps * a.Cin;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      int split, co0, tap, ci0; decode(wi, split, co0, tap, ci0);
      const int kb0 = split * k_per_split, kb1 = min(k_total, kb0 + k_per_split);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const int co = co0 + q * 32 + lane;
      float* orow = a.partial + ((long long)split * a.Cout + co) * row_pitch + (long long)tap * a.Cin;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c), v);
        const int col0 = ci0 + c;
        if (co < a.Cout && col0 < a.Cin) {
          // an empty K range (more splits than K-blocks) must still produce zeros
          const bool empty = kb1 <= kb0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j + 4 <= a.Cin) {
              float4 o = empty ? make_float4(0.f, 0.f, 0.f, 0.f)
                               : make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              *reinterpret_cast<float4*>(orow + col0 + j) = o;
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

__global__ void wgradg_reduce_kernel(const float* partial, int splits, long long elems, __nv_bfloat16* out) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= elems) return;
  float4 acc = *reinterpret_cast<const float4*>(partial + i);
  for (int s = 1; s < splits; ++s) {                           // fixed order: bit-deterministic
    float4 v = *reinterpret_cast<const float4*>(partial + (long long)s * elems + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&lo); o.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(out + i) = o;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}


template <int BLOCK_N, bool B_MN>
int launch_g(const CUtensorMap& tx, const CUtensorMap& tw, const TapConvArgs& a, int num_sms, cudaStream_t stream) {
  constexpr int STAGE_BYTES = BLOCK_M * BLOCK_K * 2 + BLOCK_N * BLOCK_K * 2;
  constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  auto kern = convg_tcgen05_kernel<BLOCK_N, STAGES, B_MN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int m_tiles = (a.OW / a.BW) * (a.OH / a.BH) * ((a.N + a.BN - 1) / a.BN);
  const int tiles = m_tiles * ((a.Cn + BLOCK_N - 1) / BLOCK_N);
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, NUM_THREADS, SMEM, stream>>>(tx, tw, a);
  return (int)cudaGetLastError();
}

bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// patch shape for an iteration space of OH x OW pixels per image: BW * BH * BN == pixels
void patch_shape(int OH, int OW, int pixels, int& BW, int& BH, int& BN) {
  BW = OW < pixels ? OW : pixels;
  BH = (pixels / BW) < OH ? (pixels / BW) : OH;
  BN = pixels / (BW * BH);
}

int encode_act(PFN_encodeTiled enc, CUtensorMap* tm, const void* base, int C, int W, int H, int N, int bw, int bh, int bn, int estride) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  // with a traversal stride e the box is given as its bounding size: e * (elements wanted)
  cuuint32_t box[4] = {64, (cuuint32_t)(bw * estride), (cuuint32_t)(bh * estride), (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  return (int)enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace

// The tap table of one launch (also exported so that the host-side algebra can be tested without a GPU).
//   fprop            : every filter tap (r, s); input pixel = output pixel * stride + (r - pad, s - pad)
//   dgrad, class (ph, pw): output pixel (stride*i + ph, stride*j + pw) receives dy[i + dh, j + dw] * W[r, s] from the taps with
//                      (ph + pad - r) and (pw + pad - s) divisible by the stride; dh = (ph + pad - r) / stride, dw likewise
// Returns the number of taps; tap_index[t] = r * ks + s.
extern "C" int drc_convg_taps(int ks, int stride, int dgrad, int ph, int pw, int* dh, int* dw, int* tap_index) {
  const int pad = ks / 2;
  int n = 0;
  for (int rr = 0; rr < ks; ++rr) {
    for (int ss = 0; ss < ks; ++ss) {
      if (!dgrad) {
        dh[n] = rr - pad; dw[n] = ss - pad;
      } else {
        if ((ph + pad - rr) % stride || (pw + pad - ss) % stride) continue;
        dh[n] = (ph + pad - rr) / stride; dw[n] = (pw + pad - ss) / stride;
      }
      tap_index[n++] = rr * ks + ss;
    }
  }
  return n;
}

// 1 if the forward geometry x[N,H,W,Cin] -> y[N,H/stride,W/stride,Cout] (ks x ks filter, pad ks/2) is served.
extern "C" int drc_convg_supported(int H, int W, int Cin, int Cout, int ks, int stride) {
  if (!(ks == 1 || ks == 3) || !(stride == 1 || stride == 2)) return 0;
  if (!pow2(W) || !pow2(H) || H % stride || W % stride) return 0;
  const int OH = H / stride, OW = W / stride;
  if (OW > 64 || OW < 2 || OH < 2) return 0;
  if (Cin % 64 || Cout % 64) return 0;
  if (OW * OH < 128 && 128 % (OW * OH)) return 0;
  return 1;
}

// fprop (dgrad == 0): act = x [N,H,W,Cin]      -> out = y  [N,H/stride,W/stride,Cout]
// dgrad (dgrad == 1): act = dy[N,H/s,W/s,Cout] -> out = dx [N,H,W,Cin]
// wgt: [Cout, ks, ks, Cin] bf16 (arena layout).  H, W are always the spatial size of the forward INPUT x.
extern "C" int drc_convg(const void* act, const void* wgt, void* out, int N, int H, int W, int Cin, int Cout, int ks, int stride,
                         int dgrad, const float* bias_f32, const void* bias_bf16, int num_sms, int device, cudaStream_t stream) {
  if (!drc_convg_supported(H, W, Cin, Cout, ks, stride)) return -1;
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return (int)e; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -2;
  const int OH = H / stride, OW = W / stride;
  TapConvArgs a;
  a.N = N; a.OH = OH; a.OW = OW;                 // both passes iterate over an OH x OW grid per image (dgrad: per parity class)
  a.Cred = dgrad ? Cout : Cin; a.Cn = dgrad ? Cin : Cout;
  patch_shape(OH, OW, BLOCK_M, a.BW, a.BH, a.BN);
  a.out = (__nv_bfloat16*)out; a.bias_f32 = bias_f32; a.bias_bf16 = (const __nv_bfloat16*)bias_bf16;
  const int block_n = a.Cn >= 128 ? 128 : 64;
  CUtensorMap tx, tw;
  {
    // weights as a matrix [Cout rows][ks*ks*Cin cols]
    cuuint64_t dims[2] = {(cuuint64_t)ks * ks * Cin, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)ks * ks * Cin * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(dgrad ? BLOCK_K : block_n)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wgt), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 2000 + (int)r;
  }
  if (!dgrad) {
    int r = encode_act(enc, &tx, act, Cin, W, H, N, a.BW, a.BH, a.BN, stride);
    if (r) return 1000 + r;
    int tidx[MAX_TAPS];
    a.in_mul = stride;
    a.ntaps = drc_convg_taps(ks, stride, 0, 0, 0, a.tap_dh, a.tap_dw, tidx);
    for (int t = 0; t < a.ntaps; ++t) a.tap_wcol[t] = tidx[t] * Cin;
    a.out_H = OH; a.out_W = OW; a.out_mul = 1; a.out_oh = a.out_ow = 0;
    return block_n == 64 ? launch_g<64, false>(tx, tw, a, num_sms, stream) : launch_g<128, false>(tx, tw, a, num_sms, stream);
  }
  // dgrad: dy has OH x OW pixels per image and is read with unit stride
  int r = encode_act(enc, &tx, act, Cout, OW, OH, N, a.BW, a.BH, a.BN, 1);
  if (r) return 1000 + r;
  a.in_mul = 1; a.out_H = H; a.out_W = W; a.out_mul = stride;
  a.bias_f32 = nullptr; a.bias_bf16 = nullptr;
  bool zeroed = false;
  if (ks == 1 && stride > 1) {                    // empty parity classes exist: zero dx before any class writes into it
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)N * H * W * Cin * 2, stream);
    if (e != cudaSuccess) return (int)e;
    zeroed = true;
  }
  for (int ph = 0; ph < stride; ++ph) {
    for (int pw = 0; pw < stride; ++pw) {
      int tidx[MAX_TAPS];
      a.ntaps = drc_convg_taps(ks, stride, 1, ph, pw, a.tap_dh, a.tap_dw, tidx);
      for (int t = 0; t < a.ntaps; ++t) a.tap_wcol[t] = tidx[t] * Cin;
      if (a.ntaps == 0) {
        // this parity class of dx receives nothing (1x1 / stride 2): it has to read as zero
        if (!zeroed) {
          cudaError_t e = cudaMemsetAsync(out, 0, (size_t)N * H * W * Cin * 2, stream);
          if (e != cudaSuccess) return (int)e;
          zeroed = true;
        }
        continue;
      }
      a.out_oh = ph; a.out_ow = pw;
      int rc = block_n == 64 ? launch_g<64, true>(tx, tw, a, num_sms, stream) : launch_g<128, true>(tx, tw, a, num_sms, stream);
      if (rc) return rc;
    }
  }
  return 0;
}

namespace {
void wgrad_patch(int H, int W, int& PW, int& PH, int& PN) { patch_shape(H, W, 64, PW, PH, PN); }
}

// K splits and fp32 workspace elements.  H, W: spatial size of the forward input x.
extern "C" int drc_convg_wgrad_plan(int N, int H, int W, int Cin, int Cout, int ks, int stride, int num_sms, long long* ws_elems) {
  const int OH = H / stride, OW = W / stride, ntaps = ks * ks;
  const int block_n = Cin >= 128 ? 128 : 64;
  const int out_tiles = ((Cout + BLOCK_M - 1) / BLOCK_M) * ntaps * ((Cin + block_n - 1) / block_n);
  int PW, PH, PN;
  wgrad_patch(OH, OW, PW, PH, PN);
  const int k_total = (OW / PW) * (OH / PH) * ((N + PN - 1) / PN);
  int splits = num_sms / out_tiles;
  if (splits > k_total) splits = k_total;
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  if (ws_elems) *ws_elems = (long long)splits * Cout * ntaps * Cin;
  return splits;
}

extern "C" int drc_convg_wgrad_supported(int H, int W, int Cin, int Cout, int ks, int stride) {
  if (!(ks == 1 || ks == 3) || !(stride == 1 || stride == 2)) return 0;
  if (!pow2(W) || !pow2(H) || H % stride || W % stride) return 0;
  const int OH = H / stride, OW = W / stride;
  if (OW > 64 || OW < 2 || OH < 2) return 0;
  if (Cin % 64 || Cout % 64) return 0;
  if (OW * OH < 64 && 64 % (OW * OH)) return 0;
  return 1;
}

// dy: [N,H/s,W/s,Cout] bf16, x: [N,H,W,Cin] bf16 -> dw: [Cout,ks,ks,Cin] bf16 (arena layout).
extern "C" int drc_convg_wgrad(const void* dy, const void* x, void* dw, float* ws, int N, int H, int W, int Cin, int Cout, int ks,
                               int stride, int num_sms, int device, cudaStream_t stream) {
  if (!drc_convg_wgrad_supported(H, W, Cin, Cout, ks, stride)) return -1;
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return (int)e; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -2;
  const int OH = H / stride, OW = W / stride, ntaps = ks * ks;
  WgradGArgs a;
  a.N = N; a.H = OH; a.W = OW; a.Cin = Cin; a.Cout = Cout; a.ks = ks; a.stride = stride; a.pad = ks / 2;
  wgrad_patch(OH, OW, a.PW, a.PH, a.PN);
  a.splits = drc_convg_wgrad_plan(N, H, W, Cin, Cout, ks, stride, num_sms, nullptr);
  a.partial = ws;
  const int block_n = Cin >= 128 ? 128 : 64;
  CUtensorMap tdy, tx;
  int r = encode_act(enc, &tdy, dy, Cout, OW, OH, N, a.PW, a.PH, a.PN, 1);
  if (r) return 1000 + r;
  r = encode_act(enc, &tx, x, Cin, W, H, N, a.PW, a.PH, a.PN, stride);
  if (r) return 2000 + r;
  const int out_tiles = ((Cout + BLOCK_M - 1) / BLOCK_M) * ntaps * ((Cin + block_n - 1) / block_n);
  const int work = out_tiles * a.splits;
  const int grid = work < num_sms ? work : num_sms;
  if (block_n == 64) {
    constexpr int BN_ = 64, SB = BLOCK_M * BLOCK_K * 2 + BN_ * BLOCK_K * 2, ST = (200 * 1024) / SB > 8 ? 8 : (200 * 1024) / SB, SM = ST * SB + 1280;
    auto kern = convg_wgrad_tcgen05_kernel<BN_, ST>;
    static bool cfgd = false;
    if (!cfgd) { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); if (e != cudaSuccess) return (int)e; cfgd = true; }
    kern<<<grid, NUM_THREADS, SM, stream>>>(tdy, tx, a);
  } else {
    constexpr int BN_ = 128, SB = BLOCK_M * BLOCK_K * 2 + BN_ * BLOCK_K * 2, ST = (200 * 1024) / SB > 8 ? 8 : (200 * 1024) / SB, SM = ST * SB + 1280;
    auto kern = convg_wgrad_tcgen05_kernel<BN_, ST>;
    static bool cfgd = false;
    if (!cfgd) { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); if (e != cudaSuccess) return (int)e; cfgd = true; }
    kern<<<grid, NUM_THREADS, SM, stream>>>(tdy, tx, a);
  }
  int rc = (int)cudaGetLastError();
  if (rc) return rc;
  const long long elems = (long long)Cout * ntaps * Cin;
  wgradg_reduce_kernel<<<(unsigned)((elems / 4 + 255) / 256), 256, 0, stream>>>(ws, a.splits, elems, (__nv_bfloat16*)dw);
  return (int)cudaGetLastError();
}
