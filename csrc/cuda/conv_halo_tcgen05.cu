// K10 (convolution part, continued): HALO-REUSE 3x3 / stride-1 convolution on tcgen05 for the 64 -> 64 channel layers.
//
// profiles/ncu_conv.md: on the layer1 shape (128 x 64 x 32 x 32) the TMA-patch kernel of conv_tcgen05.cu -- and cuDNN's --
// is bound by L2 -> SM traffic: every filter tap re-loads the shifted activation patch (9 x 16.9 MB = 147 MB at 7.5 TB/s)
// and re-loads its 8 KB weight slice for every tile.  With Cin = Cout = 64 everything a CTA needs fits in shared memory:
//
//   * the nine weight taps (9 x 8 KB) are loaded ONCE per CTA and stay resident (the CTA is persistent over tiles);
//   * an M tile is an 8 x 16 pixel block; its (8+2) x (16+2) HALO patch (180 pixel rows x 128 B = 23 KB) is loaded ONCE per
//     tile by one 4-D TMA box (out-of-image pixels zero-filled); alternatively the box is 16 pixels wide (`pw` = 16, 36 KB)
//     so that consecutive 8-row groups are 2048 B apart -- a multiple of the 1024-byte swizzle period;
//   * the A operand of tap (r, s) is that same shared-memory patch addressed through a row-shifted UMMA descriptor: start
//     address + (r*10 + s) * 128 B, stride between 8-row groups = 10 rows = 1280 B (an 8-pixel image row is one 8-row group,
//     the next image row starts 10 patch rows later).  The 128B swizzle is an XOR of address bits [4:6] with bits [7:9], the
//     patch base is 1024-byte aligned, so what TMA wrote and what the MMA unit reads agree for any row-aligned start.
//
// L2 traffic per tile drops from 9 x (16 + 8) KB to 23 KB; the 36 MMAs of a tile (9 taps x 4 K-steps) run back to back on
// one resident patch.  dgrad is the same kernel with mirrored taps and the weight tile read MN-major (same shared-memory image).
//
// `desc_mode` selects how the descriptor's 3-bit base-offset field is filled for the row-shifted starts (0: zero,
// 1: (start_address >> 7) & 7), `pw` the patch row pitch (10 or 16 pixels).  If the MMA unit derives the swizzle phase from
// absolute address bits, (pw = 10, mode 0) is right and cheapest; if it derives it from the row index inside an 8-row group
// plus the base offset, only (pw = 16, mode 1) can work.  The documentation available offline does not settle it; the
// hardware will (four combinations, one test run each).
//
// STATUS: compiled for sm_100a, SASS checked, NOT yet run on hardware: opt-in via DRACO_CONV3X3=halo, test gated by
// DRACO_EXPERIMENTAL=1.
//
// Reference counterpart: the 64 -> 64 BasicBlock convolutions of src/model_ops/resnet.py:14-36.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 64;          // Cout (fprop) / Cin (dgrad)
constexpr int BLOCK_K = 64;          // the 64 reduction channels: ONE K-slice per tap
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 256;
constexpr int TW = 8, TH = 16;       // tile = 8 x 16 pixels
constexpr int PH = TH + 2;                           // patch rows (image rows incl. halo)
constexpr int PW_MAX = 16;                           // patch row pitch in pixels: 10 (dense) or 16 (8-row groups 1024-B periodic)
constexpr int PATCH_STRIDE = PW_MAX * PH * 128;      // 36864 = 36 x 1024: stage pitch for either layout
constexpr int W_TAP_BYTES = BLOCK_N * 128;           // 8 KB
constexpr int W_BYTES = 9 * W_TAP_BYTES;             // 72 KB
constexpr int STAGES = 4;
constexpr int SMEM_BYTES = W_BYTES + STAGES * PATCH_STRIDE + 1024 + 256;

struct HaloArgs {
  int N, H, W;
  __nv_bfloat16* out;              // [N, H, W, 64]
  const float* bias_f32;
  const __nv_bfloat16* bias_bf16;
  int dgrad;
  int desc_mode;
  int pw;                          // patch row pitch in pixels (10 or 16)
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


// K-major (or MN-major for the dgrad weight tile) SWIZZLE_128B descriptor with an explicit stride between 8-row groups and
// an explicit base-offset field
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t lbo_field, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)lbo_field << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= 2ull << 61;
  return d;
}

template <bool B_MN>
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

template <bool DGRAD>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_halo_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const HaloArgs a) {
  constexpr int TMEM_COLS = 128;                    // 2 accumulators x 64 columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;                              // 9 taps x [64 rows][128 B]
  uint8_t* s_patch = smem + W_BYTES;                // STAGES x PATCH_STRIDE
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_patch + STAGES * PATCH_STRIDE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_bar = tmem_empty + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wt = a.W / TW, ht = a.H / TH;
  const int num_tiles = wt * ht * a.N;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
    mbar_init(w_bar, 1);
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
    // ===================== TMA producer: weights once, then one halo patch per tile =====================
    if (elect_one()) {
      mbar_expect_tx(w_bar, W_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(s_w + tap * W_TAP_BYTES, &tmap_w, tap * 64, 0, w_bar);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int w0 = (tile % wt) * TW, h0 = ((tile / wt) % ht) * TH, n0 = tile / (wt * ht);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], (uint32_t)(a.pw * PH * 128));
        tma_load_4d(s_patch + stage * PATCH_STRIDE, &tmap_x, 0, w0 - 1, h0 - 1, n0, &full_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 9 taps x 4 K-steps on one resident patch =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc<DGRAD>();
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
          const uint32_t a_addr = patch + (uint32_t)(pr * a.pw + ps) * 128u;
          const uint32_t b_addr = smem_u32(s_w + tap * W_TAP_BYTES);
          const uint32_t boff = a.desc_mode == 1 ? ((a_addr >> 7) & 7u) : 0u;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_desc(a_addr + k * UMMA_K * 2, (uint32_t)a.pw * 128u, 1, boff);
            // weights: K-major rows = Cout (fprop) or MN-major rows = Cout = K (dgrad); 64 rows -> 8 groups of 1024 B
            const uint64_t db = DGRAD ? make_desc(b_addr + k * UMMA_K * 128, 1024, (BLOCK_K * 128) >> 4, 0)
                                      : make_desc(b_addr + k * UMMA_K * 2, 1024, 1, 0);
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
    // ===================== epilogue =====================
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int w0 = (tile % wt) * TW, h0 = ((tile / wt) % ht) * TH, n0 = tile / (wt * ht);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const int m = q * 32 + lane;                               // tile row = pixel (x fastest)
      const int pw = w0 + (m % TW), ph = h0 + (m / TW);
      __nv_bfloat16* orow = a.out + (((long long)n0 * a.H + ph) * a.W + pw) * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c), v);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (a.bias_f32 || a.bias_bf16) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += a.bias_f32 ? a.bias_f32[c + j] : __bfloat162float(a.bias_bf16[c + j]);
        }
        __nv_bfloat16* dst = orow + c;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(f[j], f[j + 1]), p1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), p3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
          uint4 o;
          o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
          o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
          *reinterpret_cast<uint4*>(dst + j) = o;
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


}  // namespace

extern "C" int drc_conv_halo_supported(int H, int W, int Cin, int Cout) {
  return Cin == 64 && Cout == 64 && W % TW == 0 && H % TH == 0 && W >= TW && H >= TH;
}

// act: [N,H,W,64] bf16 (x for fprop, dy for dgrad); wgt: [64,3,3,64] bf16 (arena layout); out: [N,H,W,64] bf16.
extern "C" int drc_conv_halo(const void* act, const void* wgt, void* out, int N, int H, int W, int dgrad, const float* bias_f32,
                             const void* bias_bf16, int desc_mode, int patch_w, int num_sms, int device, cudaStream_t stream) {
  if (patch_w != 10 && patch_w != 16) return -3;
  if (!drc_conv_halo_supported(H, W, 64, 64)) return -1;
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return (int)e; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -2;
  HaloArgs a;
  a.N = N; a.H = H; a.W = W; a.out = (__nv_bfloat16*)out; a.bias_f32 = bias_f32; a.bias_bf16 = (const __nv_bfloat16*)bias_bf16;
  a.dgrad = dgrad; a.desc_mode = desc_mode; a.pw = patch_w;
  CUtensorMap tx, tw;
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    cuuint32_t box[4] = {64, (cuuint32_t)patch_w, PH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(act), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1000 + (int)r;
  }
  {
    cuuint64_t dims[2] = {9 * 64, 64};
    cuuint64_t strides[1] = {9 * 64 * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wgt), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 2000 + (int)r;
  }
  const int tiles = (W / TW) * (H / TH) * N;
  const int grid = tiles < num_sms ? tiles : num_sms;
  static bool cfg0 = false, cfg1 = false;
  if (!dgrad) {
    auto kern = conv_halo_tcgen05_kernel<false>;
    if (!cfg0) { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); if (e != cudaSuccess) return (int)e; cfg0 = true; }
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tx, tw, a);
  } else {
    auto kern = conv_halo_tcgen05_kernel<true>;
    if (!cfg1) { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); if (e != cudaSuccess) return (int)e; cfg1 = true; }
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tx, tw, a);
  }
  return (int)cudaGetLastError();
}
