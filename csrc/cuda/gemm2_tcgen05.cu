// K10 (dense part, continued): bf16 GEMM on CTA PAIRS -- tcgen05.mma.cta_group::2, 256 x BLOCK_N tiles.
//
// The single-CTA kernel (gemm_tcgen05.cu) tops out at 0.64x cuBLAS on 8192^3 (profiles/gemm_bench_v1.json): a 128 x 256 UMMA still
// pulls 96 B/clk of operands out of one SM's shared memory and every CTA re-loads the full B tile.  Here two CTAs of a cluster (the
// two SMs of a TPC) execute ONE MMA of M = 256: each CTA stages its own 128 rows of A and HALF of the B tile (BLOCK_N / 2 rows), the
// tensor cores of both SMs read both halves, so per-SM operand traffic (shared memory AND L2 -> SM) for B is halved, and each CTA
// keeps the accumulator of its 128 rows in its own TMEM.
//
// Protocol (one persistent pair per TPC, tiles strided over the pairs):
//   * both CTAs: warp 0 = TMA producer for ITS operand slices; the loads of both CTAs complete on the LEADER's (rank 0) full barrier
//     (cp.async.bulk.tensor ... .cta_group::2 with the leader's barrier address in cluster shared memory); the leader's
//     arrive.expect_tx announces the bytes of both CTAs, the peer never arrives (no cross-CTA fence on the K loop);
//   * leader only: warp 1 issues tcgen05.mma.cta_group::2; tcgen05.commit.cta_group::2 ... multicast::cluster releases the smem
//     stage in BOTH CTAs (each producer waits on its own empty barrier) and publishes the accumulator to BOTH epilogues;
//   * both CTAs: warps 4-7 drain their 128 TMEM lanes; all eight epilogue warps arrive on the leader's tmem_empty barrier
//     (the peer's through mapa + mbarrier.arrive.shared::cluster);
//   * TMEM is allocated / freed with cta_group::2 by one warp of EACH CTA, bracketed by cluster barriers.
// K-major operands (C = A * B^T with A [M,K], B [N,K] row-major); the other operand orders stay on the single-CTA kernel.
//
// The reference has no counterpart (dense math is PyTorch-0.3 CPU THNN).
#include "tcgen05_common.cuh"

namespace {

using namespace tc;

constexpr int BLOCK_M = 128;          // rows per CTA; the pair's tile is 256 rows
constexpr int BLOCK_K = 64;
constexpr int NUM_THREADS = 256;

struct Gemm2Args {
  int M, N, K;
  void* C;
  long long ldc;
  int c_fp32;
  const float* bias_f32;
  const __nv_bfloat16* bias_bf16;
  int relu;
  int accumulate;
};

// Tile rasterisation: tiles are walked in groups of GROUP_M row blocks x all column blocks, column-major inside a group, so the ~74
// tiles in flight at any time cover ~8 x 9 blocks: (8 + 9) operand panels per wave instead of (2 + 32) -- the B panel set of an
// 8192^3 problem no longer streams through L2 once per wave.
constexpr int GROUP_M = 8;
__device__ __forceinline__ void tile_coords(int tile, int m_blocks, int n_blocks, int& mb, int& nb) {
  const int per_group = GROUP_M * n_blocks;
  const int g = tile / per_group, first = g * GROUP_M;
  const int gm = (m_blocks - first) < GROUP_M ? (m_blocks - first) : GROUP_M;
  const int r = tile - g * per_group;
  mb = first + r % gm;
  nb = r / gm;
}

template <int BLOCK_N, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Gemm2Args args) {
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;                    // this CTA's 128 rows of A
  constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;              // this CTA's half of the B tile
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = tmem_cols_for(2 * BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);       // meaningful in the leader
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;                                                // meaningful in the leader
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_blocks = (args.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int n_blocks = (args.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_blocks * n_blocks;
  const int k_blocks = (args.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 8); }
    mbar_fence_init();
  }
  cluster_sync_all();                                               // both CTAs' barriers exist before anyone signals remotely
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, each its own slices) =====================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        int mb, nb;
        tile_coords(tile, m_blocks, n_blocks, mb, nb);
        const int m0 = mb * (2 * BLOCK_M) + (int)rank * BLOCK_M;
        const int n0 = nb * BLOCK_N + (int)rank * (BLOCK_N / 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const uint32_t lead_full = map_to_cta(&full_bar[stage], 0);
          // ONE arrival (the leader's) carrying the bytes of BOTH CTAs; the peer only issues its loads -- they complete on the
          // leader's barrier (first measurement: a per-stage cluster-scope arrive from the peer = a memory barrier per K block,
          // tensor pipe 38 % active, profiles/ncu_gemm2.md)
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          tma_load_2d_2sm(sa, &tmap_a, k0, m0, lead_full);
          tma_load_2d_2sm(sb, &tmap_b, k0, n0, lead_full);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc(2 * BLOCK_M, BLOCK_N, false, false);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint64_t da = desc_kmajor(sa), db = desc_kmajor(sb);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
            umma_f16_2sm(tmem_d, da + adv, db + adv, idesc, (kb | k) ? 1u : 0u);
          }
          commit_2sm(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        commit_2sm(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs: 128 rows each) =====================
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    const bool vec_ok = (args.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(args.C) & 15) == 0);
    for (int tile = pair; tile < num_tiles; tile += npairs) {
      int mb, nb;
      tile_coords(tile, m_blocks, n_blocks, mb, nb);
      const int m0 = mb * (2 * BLOCK_M) + (int)rank * BLOCK_M, n0 = nb * BLOCK_N;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < args.M;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c), v);
        const int col0 = n0 + c;
        if (row_ok && col0 < args.N) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (args.bias_f32 || args.bias_bf16) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < args.N) f[j] += args.bias_f32 ? args.bias_f32[col0 + j] : __bfloat162float(args.bias_bf16[col0 + j]);
          }
          if (args.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          const bool full = col0 + 32 <= args.N;
          if (args.c_fp32) {
            float* dst = reinterpret_cast<float*>(args.C) + (long long)row * args.ldc + col0;
            if (full && vec_ok && !args.accumulate) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (col0 + j < args.N) dst[j] = args.accumulate ? dst[j] + f[j] : f[j];
            }
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(args.C) + (long long)row * args.ldc + col0;
            if (full && vec_ok && !args.accumulate) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) *reinterpret_cast<uint4*>(dst + j) = pack8(f + j);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < args.N) dst[j] = __float2bfloat16_rn(args.accumulate ? __bfloat162float(dst[j]) + f[j] : f[j]);
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[acc]);
        else remote_arrive(map_to_cta(&tmem_empty[acc], 0));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();                                               // no MMA / remote arrive of the pair is still in flight
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

template <int BLOCK_N>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const Gemm2Args& g, int num_sms, cudaStream_t stream) {
  constexpr int STAGE_BYTES = BLOCK_M * BLOCK_K * 2 + (BLOCK_N / 2) * BLOCK_K * 2;
  constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  auto kern = gemm2_bf16_tcgen05_kernel<BLOCK_N, STAGES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int tiles = ((g.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * ((g.N + BLOCK_N - 1) / BLOCK_N);
  const int max_pairs = num_sms / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  kern<<<2 * pairs, NUM_THREADS, SMEM, stream>>>(ta, tb, g);
  return (int)cudaGetLastError();
}

}  // namespace

// C[M,N] (+)= A[M,K] * B[N,K]^T, both operands K-major (row-major), on CTA pairs.  block_n: 128 or 256 (0 = choose).
extern "C" int drc_gemm2_bf16(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int c_fp32, int M, int N,
                              int K, const float* bias_f32, const void* bias_bf16, int relu, int accumulate, int block_n, int num_sms,
                              int device, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return (int)e; }
  if ((lda % 8) || (ldb % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return -3;
  if (block_n == 0) block_n = (N % 256 == 0 || N > 512) ? 256 : 128;
  if (block_n != 128 && block_n != 256) return -4;
  CUtensorMap ta, tb;
  int r = encode_mat(&ta, A, M, K, lda, BLOCK_M);
  if (r) return 1000 + r;
  r = encode_mat(&tb, B, N, K, ldb, block_n / 2);
  if (r) return 2000 + r;
  Gemm2Args g;
  g.M = M; g.N = N; g.K = K; g.C = C; g.ldc = ldc; g.c_fp32 = c_fp32; g.bias_f32 = bias_f32;
  g.bias_bf16 = reinterpret_cast<const __nv_bfloat16*>(bias_bf16); g.relu = relu; g.accumulate = accumulate;
  return block_n == 256 ? launch2<256>(ta, tb, g, num_sms, stream) : launch2<128>(ta, tb, g, num_sms, stream);
}
