// Epilogue shared by the tcgen05 convolution kernels (conv_tap_tcgen05.cu, conv_halo_tcgen05.cu).
//
// A finished 128-pixel x BLOCK_N accumulator tile leaves TMEM through the four epilogue warps (thread = pixel row) and is
//   1. converted to bf16 (+ bias) and written into a shared-memory staging tile in the 128B-swizzled layout TMA expects
//      (row = pixel, 16-byte chunk j of row r at chunk j ^ (r & 7): conflict-free 16-byte stores);
//   2. handed to ONE TMA tensor store per 64-channel half (cp.async.bulk.tensor...global.shared::cta, SASS UTMASTG) -- the
//      4-D box mirrors the load box, so partial batches are clipped by the TMA unit and every global write is a full line;
//   3. (optional) reduced per channel for the BatchNorm that follows the convolution: every epilogue thread owns one channel
//      (x one row range) of the staged bf16 tile and accumulates sum / sum-of-squares into shared-memory accumulators that live
//      for the whole persistent CTA.  At kernel end each CTA writes its partial [2][C] to a workspace slot and the LAST CTA
//      (atomic ticket) folds the slots in a fixed order with 16-byte loads and finalises mean / invstd / running statistics:
//      the statistics pass of the BatchNorm (one full read of the activation + one launch per layer) disappears.
//
// Determinism: tile -> CTA assignment, in-CTA accumulation order and the slot fold order depend only on the problem shape and
// the grid size, never on timing; replicas of a batch on different GPUs stay bit-identical (the exact-equality vote needs it).
//
// Reference counterpart: nn.Conv2d followed by nn.BatchNorm2d in src/model_ops/resnet.py:19-24, vgg.py:46-59.
#pragma once
#include "tcgen05_common.cuh"

namespace convepi {

constexpr int EPI_THREADS = 128;          // warps 4..7
constexpr int EPI_BAR_ID = 1;
constexpr int STAT_MAX_C = 512;

struct BnStatArgs {
  float* partial;                 // [slots][2][C] workspace, null: no statistics
  unsigned int* counter;          // zero on entry, reset by the last CTA
  float* mean;                    // [C] out
  float* invstd;                  // [C] out
  float* running_mean;            // [C] or null
  float* running_var;
  long long count;                // elements per channel (N * OH * OW)
  float eps, momentum;
};

// shared-memory bytes the epilogue needs: staging tile + statistics accumulators
__host__ __device__ constexpr int staging_bytes(int block_n) { return block_n * 128 * 2; }          // 128 rows x block_n bf16
__host__ __device__ constexpr int stat_bytes() { return 2 * 2 * STAT_MAX_C * 4; }                   // [parts <= 2][2][512] fp32

// Drain one accumulator tile.  Called by all 128 epilogue threads (et = 0..127 = tile row).
//   tmem_acc : TMEM address of column 0 of this accumulator (lane field 0)
//   sbuf     : staging tile (1024-byte aligned), s_stat: statistics accumulators or null
//   valid_rows: rows of the tile that are real pixels (partial batch tiles), col0: first output channel of the tile
template <int BLOCK_N>
__device__ __forceinline__ void drain_tile(uint32_t tmem_acc, uint8_t* sbuf, float* s_stat, int et, int valid_rows, int col0, int Cn,
                                           const float* bias_f32, const __nv_bfloat16* bias_bf16, uint64_t* tmem_empty_bar) {
  const int q = et >> 5, lane = et & 31;
  // the previous tile's TMA store must have finished READING the staging tile before it is overwritten
  if (et == 0) tc::tma_store_wait_read<0>();
  tc::named_bar_sync(EPI_BAR_ID, EPI_THREADS);
#pragma unroll 1
  for (int c = 0; c < BLOCK_N; c += 32) {
    uint32_t v[32];
    tc::tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (bias_f32 || bias_bf16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int cc = col0 + c + j;
        if (cc < Cn) f[j] += bias_f32 ? bias_f32[cc] : __bfloat162float(bias_bf16[cc]);
      }
    }
    uint8_t* row = sbuf + (c >> 6) * (128 * 128) + et * 128;
    const int lc0 = (c & 63) >> 3;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int phys = (lc0 + jj) ^ (et & 7);
      *reinterpret_cast<uint4*>(row + phys * 16) = tc::pack8(f + jj * 8);
    }
  }
  // TMEM fully read by this warp -> the MMA warp may reuse the accumulator
  tc::tcgen05_fence_before();
  __syncwarp();
  if (lane == 0) tc::mbar_arrive(tmem_empty_bar);
  tc::fence_async_smem();                          // generic-proxy writes -> visible to the TMA store
  tc::named_bar_sync(EPI_BAR_ID, EPI_THREADS);
  (void)valid_rows; (void)s_stat;
}

// per-channel sum / sum of squares of the staged (bf16-rounded) tile
template <int BLOCK_N>
__device__ __forceinline__ void accumulate_stats(const uint8_t* sbuf, float* s_stat, int et, int valid_rows, int col0, int Cn) {
  constexpr int PARTS = EPI_THREADS / BLOCK_N;                     // 2 (BLOCK_N = 64) or 1 (128)
  constexpr int ROWS = 128 / PARTS;
  const int ch = et % BLOCK_N, part = et / BLOCK_N;
  if (col0 + ch >= Cn) return;
  const uint8_t* base = sbuf + (ch >> 6) * (128 * 128) + (ch & 7) * 2;
  const int lc = (ch & 63) >> 3;
  float s = 0.f, qq = 0.f;
  const int r0 = part * ROWS;
  int r1 = r0 + ROWS; if (r1 > valid_rows) r1 = valid_rows;
#pragma unroll 8
  for (int r = r0; r < r1; ++r) {
    const float v = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(base + r * 128 + ((lc ^ (r & 7)) << 4)));
    s += v;
    qq = fmaf(v, v, qq);
  }
  float* acc = s_stat + (part * 2) * STAT_MAX_C + col0 + ch;       // slot owned by exactly this thread for this tile
  acc[0] += s;
  acc[STAT_MAX_C] += qq;
}

// Kernel tail, executed by ALL threads of the CTA after the role loops joined (__syncthreads before the call).
//   slot / nslots: workspace row of this CTA / number of rows the last CTA folds
//   c_lo, c_hi   : channel range this CTA writes into its slot (the others in the slot are written by sibling CTAs or are zero)
template <int NUM_THREADS>
__device__ __forceinline__ void finalize_stats(const BnStatArgs& st, float* s_stat, int parts, int Cn, int slot, int nslots,
                                               int c_lo, int c_hi, float* s_scratch /* >= NUM_THREADS * 4 floats */) {
  __shared__ int s_is_last;
  for (int i = threadIdx.x; i < 2 * (c_hi - c_lo); i += NUM_THREADS) {
    const int a = i / (c_hi - c_lo), c = c_lo + i % (c_hi - c_lo);
    float t = 0.f;
    for (int p = 0; p < parts; ++p) t += s_stat[(p * 2 + a) * STAT_MAX_C + c];
    st.partial[((long long)slot * 2 + a) * Cn + c] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(st.counter, 1u);
    s_is_last = (prev == gridDim.x - 1);
    if (s_is_last) *st.counter = 0;
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  // fold: items4 = 2C/4 float4 columns, SUB interleaved subsets of the slots; fixed order everywhere
  const int items4 = (2 * Cn) >> 2;
  const int SUB = NUM_THREADS / items4 > 0 ? NUM_THREADS / items4 : 1;
  const float4* p4 = reinterpret_cast<const float4*>(st.partial);
  for (int i0 = 0; i0 < items4; i0 += NUM_THREADS) {                 // one pass unless 2C/4 > NUM_THREADS
    const int item = i0 + (int)threadIdx.x % (items4 < NUM_THREADS ? items4 : NUM_THREADS);
    const int sub = items4 < NUM_THREADS ? (int)threadIdx.x / items4 : 0;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    if (sub < SUB && item < items4) {
      int g = sub;
      for (; g + 3 * SUB < nslots; g += 4 * SUB) {
        const float4 v0 = __ldcg(p4 + (long long)g * items4 + item);
        const float4 v1 = __ldcg(p4 + (long long)(g + SUB) * items4 + item);
        const float4 v2 = __ldcg(p4 + (long long)(g + 2 * SUB) * items4 + item);
        const float4 v3 = __ldcg(p4 + (long long)(g + 3 * SUB) * items4 + item);
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
        a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
        a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
      }
      for (; g < nslots; g += SUB) {
        const float4 v0 = __ldcg(p4 + (long long)g * items4 + item);
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
      }
    }
    float4 t;
    t.x = (a0.x + a1.x) + (a2.x + a3.x); t.y = (a0.y + a1.y) + (a2.y + a3.y);
    t.z = (a0.z + a1.z) + (a2.z + a3.z); t.w = (a0.w + a1.w) + (a2.w + a3.w);
    __syncthreads();
    reinterpret_cast<float4*>(s_scratch)[threadIdx.x] = t;
    __syncthreads();
    // subset 0 combines the subsets in order and leaves the totals in s_stat[0 .. 2C)
    if (sub == 0 && item < items4) {
      float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
      const int nsub = items4 < NUM_THREADS ? SUB : 1;
      for (int s2 = 0; s2 < nsub; ++s2) {
        const float4 v = reinterpret_cast<const float4*>(s_scratch)[s2 * items4 + (item - i0)];
        tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
      }
      reinterpret_cast<float4*>(s_stat)[item] = tot;                 // s_stat reused: [2][C] totals, a-major
    }
  }
  __syncthreads();
  const float inv_m = 1.0f / (float)st.count;
  for (int c = threadIdx.x; c < Cn; c += NUM_THREADS) {
    const float m = s_stat[c] * inv_m;
    float var = fmaf(-m, m, s_stat[Cn + c] * inv_m);
    var = var < 0.f ? 0.f : var;
    st.mean[c] = m;
    st.invstd[c] = rsqrtf(var + st.eps);
    if (st.running_mean) {
      const float unbiased = st.count > 1 ? var * ((float)st.count / (float)(st.count - 1)) : var;
      st.running_mean[c] = fmaf(st.momentum, m - st.running_mean[c], st.running_mean[c]);
      st.running_var[c] = fmaf(st.momentum, unbiased - st.running_var[c], st.running_var[c]);
    }
  }
}

}  // namespace convepi
