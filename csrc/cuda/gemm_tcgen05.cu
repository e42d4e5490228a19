// K10 (dense part): bf16 GEMM on the 5th-generation tensor cores, written directly against the sm_100a
// programming model -- TMA (cp.async.bulk.tensor) stages 128B-swizzled operand tiles in shared memory, a
// single elected thread issues tcgen05.mma with the fp32 accumulator in TMEM, tcgen05.commit hands smem
// stages back to the producer and accumulator stages to the epilogue warps, which drain TMEM with
// tcgen05.ld and apply bias / ReLU / down-conversion on the way to global memory.
//
//   C[M,N] = op(A) * op(B)^T   with  A: M x K,  B: N x K  (both described K-major *or* MN-major, so the
//   three products of a linear layer -- y = x W^T, dx = dy W, dW = dy^T x -- run without a transpose pass)
//
// Role layout (256 threads, one persistent CTA per SM, tiles strided over the grid):
//   warp 0     TMA producer      (one elected lane)
//   warp 1     MMA issuer        (one elected lane; UMMA 128 x BLOCK_N x 16, cta_group::1)
//   warp 2     TMEM allocator
//   warps 4-7  epilogue          (warp q drains TMEM lanes 32q..32q+31 = rows 32q..32q+31 of the tile)
// Pipelines: smem full/empty ring of STAGES stages; TMEM full/empty with two accumulator stages so the
// epilogue of tile i overlaps the main loop of tile i+1.
//
// The reference has no counterpart (its dense math is PyTorch-0.3 CPU THNN); this is the engine behind
// draco_b200.ops.linear and the im2col convolution path.
#include "tcgen05_common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;           // 64 bf16 = 128 bytes = one swizzle-128B atom row
constexpr int NUM_THREADS = 256;

struct GemmArgs {
  int M, N, K;
  void* C;
  long long ldc;                      // elements
  int c_fp32;                         // 1: C is fp32, 0: bf16
  const float* bias_f32;              // optional bias[N]
  const __nv_bfloat16* bias_bf16;     // optional bias[N]
  int relu;
  int accumulate;                     // C += result (fp32 or bf16 read-modify-write)
};

using namespace tc;

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B.
//   K-major  tile [rows][64 k]        : 8-row groups 1024 B apart (SBO); LBO unused (=1 like CUTLASS)
//   MN-major tile [64 k][64 mn] boxes : 8-k-row groups 1024 B apart (SBO); 64-wide MN atoms BLOCK_K*128 B apart (LBO)
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  const uint64_t sbo = 1024 >> 4;
  const uint64_t lbo = MN_MAJOR ? ((BLOCK_K * 128) >> 4) : 1;
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= lbo << 16;
  d |= sbo << 32;
  d |= 1ull << 46;                    // descriptor version 1 (sm_100)
  d |= 2ull << 61;                    // SWIZZLE_128B
  return d;
}

template <int BLOCK_N, bool A_MN, bool B_MN>
__device__ __forceinline__ uint32_t gemm_idesc() {
  uint32_t d = 0;
  d |= 1u << 4;                       // D format  : F32
  d |= 1u << 7;                       // A format  : BF16
  d |= 1u << 10;                      // B format  : BF16
  d |= (A_MN ? 1u : 0u) << 15;        // A major
  d |= (B_MN ? 1u : 0u) << 16;        // B major
  d |= (uint32_t)(BLOCK_N >> 3) << 17;
  d |= (uint32_t)(BLOCK_M >> 4) << 24;
  return d;
}

// Tile rasterisation: groups of GROUP_M row blocks x all column blocks, column-major inside a group (the tiles in flight cover a
// roughly square block of the output: fewer distinct operand panels per wave -> better L2 reuse on large problems).
constexpr int GROUP_M = 8;
__device__ __forceinline__ void tile_coords(int tile, int m_blocks, int n_blocks, int& mb, int& nb) {
  const int per_group = GROUP_M * n_blocks;
  const int g = tile / per_group, first = g * GROUP_M;
  const int gm = (m_blocks - first) < GROUP_M ? (m_blocks - first) : GROUP_M;
  const int r = tile - g * per_group;
  mb = first + r % gm;
  nb = r / gm;
}

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmArgs args) {
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
  static_assert(!B_MN || BLOCK_N % 64 == 0, "MN-major B needs 64-wide atoms");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blocks = (args.M + BLOCK_M - 1) / BLOCK_M;
  const int n_blocks = (args.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_blocks * n_blocks;
  const int k_blocks = (args.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
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
        int mb_, nb_;
        tile_coords(tile, m_blocks, n_blocks, mb_, nb_);
        const int m0 = mb_ * BLOCK_M, n0 = nb_ * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          if (!A_MN) tma_load_2d(sa, &tmap_a, k0, m0, &full_bar[stage]);
          else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j) tma_load_2d(sa + j * (BLOCK_K * 128), &tmap_a, m0 + 64 * j, k0, &full_bar[stage]);
          }
          if (!B_MN) tma_load_2d(sb, &tmap_b, k0, n0, &full_bar[stage]);
          else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j) tma_load_2d(sb + j * (BLOCK_K * 128), &tmap_b, n0 + 64 * j, k0, &full_bar[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = gemm_idesc<BLOCK_N, A_MN, B_MN>();
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
          const uint64_t da = make_smem_desc<A_MN>(sa), db = make_smem_desc<B_MN>(sb);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance along K inside the stage: K-major +32 B per UMMA_K, MN-major +16 rows * 128 B
            const uint64_t adv_a = (uint64_t)((A_MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
            const uint64_t adv_b = (uint64_t)((B_MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
            umma_f16(tmem_d, da + adv_a, db + adv_b, idesc, (kb | k) ? 1u : 0u);
          }
          tcgen05_commit(&empty_bar[stage]);          // frees the smem stage once the MMAs above retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&tmem_full[acc]);              // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    int acc = 0; uint32_t acc_phase = 0;
    const bool vec_ok = (args.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(args.C) & 15) == 0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int mb_, nb_;
        tile_coords(tile, m_blocks, n_blocks, mb_, nb_);
        const int m0 = mb_ * BLOCK_M, n0 = nb_ * BLOCK_N;
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
            for (int j = 0; j < 32; ++j) {
              if (col0 + j < args.N)
                f[j] += args.bias_f32 ? args.bias_f32[col0 + j] : __bfloat162float(args.bias_bf16[col0 + j]);
            }
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
              for (int j = 0; j < 32; ++j)
                if (col0 + j < args.N) dst[j] = __float2bfloat16_rn(args.accumulate ? __bfloat162float(dst[j]) + f[j] : f[j]);
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

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// operand X: logical [rows(MN), K].  mn_major == 0: memory is [rows][ld] (K contiguous).  mn_major == 1: memory is [K][ld] (MN contiguous).
int make_tmap(CUtensorMap* out, const void* ptr, long long rows, long long K, long long ld, int mn_major, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[2], strides[1];
  cuuint32_t box[2], estr[2] = {1, 1};
  if (!mn_major) { dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows; box[0] = BLOCK_K; box[1] = (cuuint32_t)box_rows; }
  else           { dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K; box[0] = 64; box[1] = BLOCK_K; }
  strides[0] = (cuuint64_t)ld * 2;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

template <int BLOCK_N, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g, int num_sms, cudaStream_t stream) {
  constexpr int STAGE_BYTES = BLOCK_M * BLOCK_K * 2 + BLOCK_N * BLOCK_K * 2;
  constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, STAGES, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int tiles = ((g.M + BLOCK_M - 1) / BLOCK_M) * ((g.N + BLOCK_N - 1) / BLOCK_N);
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, NUM_THREADS, SMEM, stream>>>(ta, tb, g);
  return (int)cudaGetLastError();
}

template <int BLOCK_N>
int dispatch_major(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g, int sms, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<BLOCK_N, false, false>(ta, tb, g, sms, s);
  if (a_mn && !b_mn) return launch<BLOCK_N, true, false>(ta, tb, g, sms, s);
  if constexpr (BLOCK_N % 64 == 0) {
    if (!a_mn && b_mn) return launch<BLOCK_N, false, true>(ta, tb, g, sms, s);
    return launch<BLOCK_N, true, true>(ta, tb, g, sms, s);
  }
  return -2;
}

}  // namespace

// C[M,N] (+)= A * B^T.  a_mn / b_mn select the memory order of the operands (see make_tmap).
// Requirements: bf16 operands, 16-byte aligned bases, leading dimensions multiple of 8 elements.
extern "C" int drc_gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, void* C, long long ldc,
                             int c_fp32, int M, int N, int K, const float* bias_f32, const void* bias_bf16, int relu, int accumulate,
                             int block_n, int num_sms, int device, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  // autograd worker threads start without a bound context; the driver-API tensor-map encode needs one
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return (int)e; }
  if ((lda % 8) || (ldb % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return -3;
  if (block_n == 0) {
    block_n = (N > 64) ? 128 : (N > 32 ? 64 : (b_mn ? 64 : 32));
    if (N >= 256 && N % 256 == 0) {
      // prefer 128x256 tiles unless the coarser tiling loses more to wave quantisation than it gains
      const long long mb = (M + BLOCK_M - 1) / BLOCK_M;
      const long long t128 = mb * ((N + 127) / 128), t256 = mb * (N / 256);
      const double eff128 = (double)t128 / (double)(((t128 + num_sms - 1) / num_sms) * num_sms);
      const double eff256 = (double)t256 / (double)(((t256 + num_sms - 1) / num_sms) * num_sms);
      if (eff256 * 1.12 >= eff128) block_n = 256;
    }
  }
  if (b_mn && block_n < 64) block_n = 64;
  CUtensorMap ta, tb;
  int r = make_tmap(&ta, A, M, K, lda, a_mn, BLOCK_M);
  if (r) return 1000 + r;
  r = make_tmap(&tb, B, N, K, ldb, b_mn, block_n);
  if (r) return 2000 + r;
  GemmArgs g;
  g.M = M; g.N = N; g.K = K; g.C = C; g.ldc = ldc; g.c_fp32 = c_fp32; g.bias_f32 = bias_f32;
  g.bias_bf16 = reinterpret_cast<const __nv_bfloat16*>(bias_bf16); g.relu = relu; g.accumulate = accumulate;
  switch (block_n) {
    case 32: return dispatch_major<32>(a_mn, b_mn, ta, tb, g, num_sms, stream);
    case 64: return dispatch_major<64>(a_mn, b_mn, ta, tb, g, num_sms, stream);
    case 128: return dispatch_major<128>(a_mn, b_mn, ta, tb, g, num_sms, stream);
    // 128x256 tiles: one UMMA reads 4 KB of A + 8 KB of B per 128 tensor cycles (96 B/clk) instead of 128 B/clk for
    // 128x128, i.e. it is no longer pinned at the shared-memory bandwidth limit (profiles/ncu_gemm.md: 72.6 % tensor pipe)
    case 256: return dispatch_major<256>(a_mn, b_mn, ta, tb, g, num_sms, stream);
    default: return -4;
  }
}
