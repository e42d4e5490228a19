// K10 (convolution part): 3x3 / stride-1 / pad-1 implicit-GEMM convolution on tcgen05 tensor cores, NHWC bf16.
//
//   fprop : y[n,h,w,co]  = sum_{r,s,ci} x [n, h+r-1, w+s-1, ci] * W[co, r, s, ci]
//   dgrad : dx[n,h,w,ci] = sum_{r,s,co} dy[n, h-r+1, w-s+1, co] * W[co, r, s, ci]
//
// No im2col buffer exists.  An M-tile is a 128-pixel *patch* (BW x BH x BN pixels of the NHWC tensor); for every filter
// tap (r,s) and every 64-channel slice the producer issues ONE 4-D tiled TMA load of the patch shifted by the tap offset
// -- out-of-range coordinates (the padding halo) are zero-filled by the TMA unit -- which lands in shared memory as the
// same 128-row x 128-byte swizzled K-major tile a plain GEMM would use.  The weights stay in their arena layout
// [Cout][3][3][Cin]: for fprop a tap is a K-major B tile (rows = Cout), for dgrad the *same* memory is read as an
// MN-major B tile (rows = Cout = K, contiguous Cin = N), so the backward-data pass needs no weight transform at all.
// MMA issue, TMEM double buffering and the epilogue are those of gemm_tcgen05.cu; the epilogue maps a tile row back to
// its pixel (n,h,w) and writes NHWC rows (+ optional bias).
//
// The 9x re-read of the activation patch hits L2 (the whole activation tensor of a CIFAR layer is 2-17 MB).
// Fixed tile order and a single accumulation chain per output element: bit-deterministic, as the vote requires.
//
// Reference counterpart: nn.Conv2d in src/model_ops/resnet.py / vgg.py (PyTorch-0.3 CPU THNN kernels).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 256;

struct ConvArgs {
  int N, H, W;                 // activation geometry (input == output spatial size: stride 1, pad 1)
  int Cred;                    // reduction channels per tap (fprop: Cin, dgrad: Cout)
  int Cn;                      // output channels (fprop: Cout, dgrad: Cin)
  int w_tap_stride;            // elements between taps inside a weight row (= Cin)
  int BW, BH, BN;              // patch shape, BW*BH*BN == 128
  __nv_bfloat16* out;          // [N,H,W,Cn]
  const float* bias_f32;
  const __nv_bfloat16* bias_bf16;
  int dgrad;
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

template <int BLOCK_N, int STAGES, bool DGRAD>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const ConvArgs a) {
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
  constexpr bool B_MN = DGRAD;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wt = a.W / a.BW, ht = a.H / a.BH, nt = (a.N + a.BN - 1) / a.BN;
  const int m_tiles = wt * ht * nt;
  const int n_tiles = (a.Cn + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int c_blocks = a.Cred / BLOCK_K;           // 64-channel slices per tap
  const int k_blocks = 9 * c_blocks;

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
          const int r = tap / 3, s = tap - 3 * r;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          // activation patch shifted by the tap; the halo is zero-filled by TMA
          const int dw = DGRAD ? (1 - s) : (s - 1), dh = DGRAD ? (1 - r) : (r - 1);
          tma_load_4d(sa, &tmap_x, c0, w0 + dw, h0 + dh, nb0, &full_bar[stage]);
          if (!B_MN) {
            tma_load_2d(sb, &tmap_w, tap * a.w_tap_stride + c0, n0, &full_bar[stage]);          // rows = Cout tile, K-major
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)                                               // rows = Cout (K), cols = Cin (N)
              tma_load_2d(sb + j * (BLOCK_K * 128), &tmap_w, tap * a.w_tap_stride + n0 + 64 * j, c0, &full_bar[stage]);
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
      __nv_bfloat16* orow = a.out + (((long long)pn * a.H + ph) * a.W + pw) * a.Cn;
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


// ------------------------------------------------------------------------------------------------------------------
// wgrad:  dW[co, r, s, ci] = sum_{n,h,w} dy[n,h,w,co] * x[n, h+r-1, w+s-1, ci]
//
// GEMM view per filter tap: M = Cout, N = Cin, K = N*H*W pixels.  Both operands are MN-major: a K-block is a 64-pixel
// patch, the A tile is the dy patch [64 px][co] and the B tile is the x patch shifted by the tap [64 px][ci] -- both are
// exactly what a 4-D TMA box delivers (rows = pixels, 128-byte rows of 64 channels), again with the halo zero-filled.
// K is huge and the number of output tiles tiny (9 taps x Cout/128 x Cin/BLOCK_N), so the K range is split across CTAs;
// every split writes its fp32 partial tile and `wgrad_reduce_kernel` folds the splits in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------------------------------
struct WgradArgs {
  int N, H, W, Cin, Cout;
  int PW, PH, PN;               // 64-pixel patch shape
  int splits;                   // K splits
  float* partial;               // [splits][Cout][9*Cin]
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgradArgs a) {
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
  const int out_tiles = m_tiles * 9 * n_tiles;
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
    tap = t % 9; t /= 9;
    co0 = t * BLOCK_M; ci0 = ntile * BLOCK_N;
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
        int split, co0, tap, ci0; decode(wi, split, co0, tap, ci0);
        const int r = tap / 3, s = tap - 3 * r;
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
          for (int j = 0; j < BLOCK_N / 64; ++j) tma_load_4d(sb + j * (BLOCK_K * 128), &tmap_x, ci0 + 64 * j, w0 + s - 1, h0 + r - 1, n0, &full_bar[stage]);
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
    const long long row_pitch = 9LL * a.Cin;
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

__global__ void wgrad_reduce_kernel(const float* partial, int splits, long long elems, __nv_bfloat16* out) {
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

template <int BLOCK_N, bool DGRAD>
int launch(const CUtensorMap& tx, const CUtensorMap& tw, const ConvArgs& a, int num_sms, cudaStream_t stream) {
  constexpr int STAGE_BYTES = BLOCK_M * BLOCK_K * 2 + BLOCK_N * BLOCK_K * 2;
  constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  auto kern = conv3x3_tcgen05_kernel<BLOCK_N, STAGES, DGRAD>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int m_tiles = (a.W / a.BW) * (a.H / a.BH) * ((a.N + a.BN - 1) / a.BN);
  const int tiles = m_tiles * ((a.Cn + BLOCK_N - 1) / BLOCK_N);
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, NUM_THREADS, SMEM, stream>>>(tx, tw, a);
  return (int)cudaGetLastError();
}

}  // namespace

// 1 if the (H, W, Cin, Cout) geometry is served by the tcgen05 kernels.
extern "C" int drc_conv3x3_supported(int H, int W, int Cin, int Cout, int dgrad) {
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  if (!pow2(W) || !pow2(H) || W > 128 || (long long)W * H < 1) return 0;
  const int Cred = dgrad ? Cout : Cin, Cn = dgrad ? Cin : Cout;
  if (Cred % 64 || Cn % 8) return 0;
  if (dgrad && Cn % 64) return 0;                  // MN-major B tiles are 64 channels wide
  if (W * H < 128 && 128 % (W * H)) return 0;
  return 1;
}

// act: [N,H,W,Cred] bf16 (x for fprop, dy for dgrad); wgt: [Cout,3,3,Cin] bf16 (arena layout); out: [N,H,W,Cn] bf16.
extern "C" int drc_conv3x3(const void* act, const void* wgt, void* out, int N, int H, int W, int Cin, int Cout, int dgrad,
                           const float* bias_f32, const void* bias_bf16, int num_sms, int device, cudaStream_t stream) {
  if (!drc_conv3x3_supported(H, W, Cin, Cout, dgrad)) return -1;
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return (int)e; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -2;
  ConvArgs a;
  a.N = N; a.H = H; a.W = W; a.dgrad = dgrad;
  a.Cred = dgrad ? Cout : Cin; a.Cn = dgrad ? Cin : Cout; a.w_tap_stride = Cin;
  a.BW = W < 128 ? W : 128;
  a.BH = (128 / a.BW) < H ? (128 / a.BW) : H;
  a.BN = 128 / (a.BW * a.BH);
  a.out = (__nv_bfloat16*)out; a.bias_f32 = bias_f32; a.bias_bf16 = (const __nv_bfloat16*)bias_bf16;
  int block_n = a.Cn >= 128 ? 128 : 64;
  if (!dgrad && a.Cn < 64) block_n = 32;
  // activation: rank-4 {C, W, H, N}
  CUtensorMap tx, tw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)a.Cred, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)a.Cred * 2, (cuuint64_t)W * a.Cred * 2, (cuuint64_t)H * W * a.Cred * 2};
    cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)a.BW, (cuuint32_t)a.BH, (cuuint32_t)a.BN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(act), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1000 + (int)r;
  }
  {
    // weights as a matrix [Cout rows][9*Cin cols]
    cuuint64_t dims[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)9 * Cin * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(dgrad ? BLOCK_K : block_n)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wgt), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 2000 + (int)r;
  }
  if (!dgrad) {
    switch (block_n) {
      case 32: return launch<32, false>(tx, tw, a, num_sms, stream);
      case 64: return launch<64, false>(tx, tw, a, num_sms, stream);
      default: return launch<128, false>(tx, tw, a, num_sms, stream);
    }
  }
  return block_n == 64 ? launch<64, true>(tx, tw, a, num_sms, stream) : launch<128, true>(tx, tw, a, num_sms, stream);
}

// number of K splits and fp32 workspace elements for wgrad
extern "C" int drc_conv3x3_wgrad_plan(int N, int H, int W, int Cin, int Cout, int num_sms, long long* ws_elems) {
  const int block_n = Cin >= 128 ? 128 : 64;
  const int out_tiles = ((Cout + BLOCK_M - 1) / BLOCK_M) * 9 * ((Cin + block_n - 1) / block_n);
  int PW = W < 64 ? W : 64, PH = (64 / PW) < H ? (64 / PW) : H, PN = 64 / (PW * PH);
  const int k_total = (W / PW) * (H / PH) * ((N + PN - 1) / PN);
  int splits = num_sms / out_tiles;                            // one wave: out_tiles * splits <= SMs
  if (splits > k_total) splits = k_total;
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  if (ws_elems) *ws_elems = (long long)splits * Cout * 9 * Cin;
  return splits;
}

extern "C" int drc_conv3x3_wgrad_supported(int H, int W, int Cin, int Cout) {
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  if (!pow2(W) || !pow2(H) || W > 64) return 0;
  if (Cin % 64 || Cout % 64) return 0;
  if (W * H < 64 && 64 % (W * H)) return 0;
  return 1;
}

// dy: [N,H,W,Cout] bf16, x: [N,H,W,Cin] bf16 -> dw: [Cout,3,3,Cin] bf16 (arena layout).  ws: fp32 workspace from the plan.
extern "C" int drc_conv3x3_wgrad(const void* dy, const void* x, void* dw, float* ws, int N, int H, int W, int Cin, int Cout,
                                 int num_sms, int device, cudaStream_t stream) {
  if (!drc_conv3x3_wgrad_supported(H, W, Cin, Cout)) return -1;
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return (int)e; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -2;
  WgradArgs a;
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
  a.PW = W < 64 ? W : 64; a.PH = (64 / a.PW) < H ? (64 / a.PW) : H; a.PN = 64 / (a.PW * a.PH);
  a.splits = drc_conv3x3_wgrad_plan(N, H, W, Cin, Cout, num_sms, nullptr);
  a.partial = ws;
  const int block_n = Cin >= 128 ? 128 : 64;
  CUtensorMap tdy, tx;
  for (int which = 0; which < 2; ++which) {
    const int C = which ? Cin : Cout;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)a.PW, (cuuint32_t)a.PH, (cuuint32_t)a.PN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(which ? &tx : &tdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(which ? x : dy), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1000 * (which + 1) + (int)r;
  }
  const int out_tiles = ((Cout + BLOCK_M - 1) / BLOCK_M) * 9 * ((Cin + block_n - 1) / block_n);
  const int work = out_tiles * a.splits;
  const int grid = work < num_sms ? work : num_sms;
  int rc;
  if (block_n == 64) {
    constexpr int BN_ = 64, SB = BLOCK_M * BLOCK_K * 2 + BN_ * BLOCK_K * 2, ST = (200 * 1024) / SB > 8 ? 8 : (200 * 1024) / SB, SM = ST * SB + 1280;
    auto kern = conv3x3_wgrad_tcgen05_kernel<BN_, ST>;
    static bool cfgd = false;
    if (!cfgd) { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); if (e != cudaSuccess) return (int)e; cfgd = true; }
    kern<<<grid, NUM_THREADS, SM, stream>>>(tdy, tx, a);
  } else {
    constexpr int BN_ = 128, SB = BLOCK_M * BLOCK_K * 2 + BN_ * BLOCK_K * 2, ST = (200 * 1024) / SB > 8 ? 8 : (200 * 1024) / SB, SM = ST * SB + 1280;
    auto kern = conv3x3_wgrad_tcgen05_kernel<BN_, ST>;
    static bool cfgd = false;
    if (!cfgd) { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); if (e != cudaSuccess) return (int)e; cfgd = true; }
    kern<<<grid, NUM_THREADS, SM, stream>>>(tdy, tx, a);
  }
  rc = (int)cudaGetLastError();
  if (rc) return rc;
  const long long elems = (long long)Cout * 9 * Cin;
  wgrad_reduce_kernel<<<(unsigned)((elems / 4 + 255) / 256), 256, 0, stream>>>(ws, a.splits, elems, (__nv_bfloat16*)dw);
  return (int)cudaGetLastError();
}
