// K3: majority vote over repetition groups, per parameter tensor, with exact (whole-tensor) equality.
//
// Reference semantics (src/master/rep_master.py:154-168): for each group and each parameter tensor run
// Boyer-Moore over the group's members in order, comparing *whole tensors* with np.array_equal; the
// surviving candidate's tensor is that group's contribution.
//
// Device formulation: pass 1 (`vote_compare_kernel`) streams all P gradient slots once and produces, per
// (group, tensor), a bitmask of member pairs that differ anywhere (float `!=`, so NaN != NaN and +0 == -0
// exactly like np.array_equal).  Pass 2 (`vote_resolve_kernel`, one thread per (group, tensor)) replays
// Boyer-Moore on that pair table and writes the winning worker slot.  The winner table feeds the fused
// select + SGD + broadcast kernel (aggregate_update.cu).
#include "common.cuh"

struct VoteArgs {
  const float* grad_in;           // [P][slot_stride]
  long long slot_stride;
  const int* group_table;         // [G][max_r] worker slots, -1 padded
  int G, max_r;
  TileView tv;
  unsigned int* neq_mask;         // [G][ntensors] pair-mismatch bitmask, pair (i<j) -> bit i*max_r + j ... packed below
  int tile_begin, tile_end;       // bucket of tiles to compare (tile_end == 0: whole arena) -- the PS decodes a bucket as
                                  // soon as every worker has pushed it, while later buckets are still in flight
};

__device__ __forceinline__ int pair_bit(int i, int j) {   // i < j < 8  -> 0..27
  return j * (j - 1) / 2 + i;
}

// Pair mismatches of one group for this thread's float4 (members v[0..r)).
__device__ __forceinline__ unsigned int pair_mask(const float4* v, int r) {
  unsigned int m = 0u;
#pragma unroll
  for (int j = 1; j < DRC_MAX_R; ++j) {
#pragma unroll
    for (int i = 0; i < j; ++i) {
      if (j < r) {
        bool ne = (v[i].x != v[j].x) | (v[i].y != v[j].y) | (v[i].z != v[j].z) | (v[i].w != v[j].w);
        m |= ne ? (1u << pair_bit(i, j)) : 0u;
      }
    }
  }
  return m;
}

// FAST: G <= 4 groups of <= 4 members (the 7-worker / r=3 job: groups of 3 and 4).  Every member load of a tile -- all groups --
// is issued before the first compare, so a thread keeps G * r independent 16-byte loads in flight instead of r (the kernel is a
// pure stream over P gradient slabs: memory-level parallelism is the only thing that matters).
template <bool FAST>
__global__ void __launch_bounds__(DRC_THREADS) vote_compare_kernel(const __grid_constant__ VoteArgs a) {
  __shared__ int s_slot[DRC_MAX_WORKERS * DRC_MAX_R];
  for (int i = threadIdx.x; i < a.G * a.max_r; i += DRC_THREADS) s_slot[i] = a.group_table[i];
  __syncthreads();
  const int tile_end = a.tile_end > 0 ? a.tile_end : a.tv.ntiles;
  // Per-thread mismatch masks live in registers across the CTA's tiles and are flushed (warp OR + one global atomic per warp,
  // only when something differed) when the tensor changes: no barrier and no shared memory in the streaming loop, so the loads
  // of consecutive tiles overlap.
  constexpr int NG = FAST ? 4 : DRC_MAX_WORKERS;
  unsigned int acc[FAST ? 4 : 1];
#pragma unroll
  for (int g = 0; g < (FAST ? 4 : 1); ++g) acc[g] = 0u;
  int cur_tensor = -1;
  auto flush = [&](int tensor) {
    if (!FAST || tensor < 0) return;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (g < a.G) {
        const unsigned int m = __reduce_or_sync(0xffffffffu, acc[g]);
        if ((threadIdx.x & 31) == 0 && m) atomicOr(&a.neq_mask[g * a.tv.ntensors + tensor], m);
        acc[g] = 0u;
      }
    }
  };
  (void)NG;
  for (int tile = a.tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
    int tensor;
    const int valid = tile_valid(a.tv, tile, tensor);
    const long long idx = (long long)tile * DRC_TILE + threadIdx.x * 4;
    const bool active = (int)threadIdx.x * 4 < valid;     // padding is zero in every slot: skip whole float4s only
    if (FAST) {
      if (tensor != cur_tensor) { flush(cur_tensor); cur_tensor = tensor; }
      float4 v[4][4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int slot = (g < a.G && k < a.max_r) ? s_slot[g * a.max_r + k] : -1;
          if (slot >= 0 && active) v[g][k] = ld_f4(reinterpret_cast<const float4*>(a.grad_in + slot * a.slot_stride + idx));
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (g < a.G && active) {
          int r = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) if (k < a.max_r && s_slot[g * a.max_r + k] >= 0) r = k + 1;
          acc[g] |= pair_mask(v[g], r);
        }
      }
    } else {
      for (int g = 0; g < a.G; ++g) {
        float4 v[DRC_MAX_R];
        int r = 0;
#pragma unroll
        for (int k = 0; k < DRC_MAX_R; ++k) {
          int slot = k < a.max_r ? s_slot[g * a.max_r + k] : -1;
          if (slot >= 0) {
            r = k + 1;
            if (active) v[k] = ld_f4(reinterpret_cast<const float4*>(a.grad_in + slot * a.slot_stride + idx));
          }
        }
        unsigned int m = active ? pair_mask(v, r) : 0u;
        m = __reduce_or_sync(0xffffffffu, m);              // warp OR, then one global atomic per warp (mismatches are rare)
        if ((threadIdx.x & 31) == 0 && m) atomicOr(&a.neq_mask[g * a.tv.ntensors + tensor], m);
      }
    }
  }
  flush(cur_tensor);
}

struct ResolveArgs {
  const unsigned int* neq_mask;   // [G][T]
  const int* group_table;         // [G][max_r]
  int G, max_r, T;
  int* winner_slot;               // [G][T] winning worker slot
  int* winner_member;             // [G][T] winning member index (diagnostics), may be null
  unsigned int* clear_mask;       // same buffer as neq_mask: cleared for the next step after use
  int t_begin, t_end;             // tensor range to resolve (t_end == 0: all)
};

__global__ void vote_resolve_kernel(const __grid_constant__ ResolveArgs a) {
  const int t_end = a.t_end > 0 ? a.t_end : a.T;
  const int nt = t_end - a.t_begin;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.G * nt) return;
  int g = j / nt;
  const int i = g * a.T + a.t_begin + (j - g * nt);
  unsigned int mask = a.neq_mask[i];
  int cand = 0, count = 0;
  for (int k = 0; k < a.max_r; ++k) {
    if (a.group_table[g * a.max_r + k] < 0) break;
    if (count == 0) { cand = k; count = 1; }
    else {
      int lo = cand < k ? cand : k, hi = cand < k ? k : cand;
      bool equal = !((mask >> pair_bit(lo, hi)) & 1u);
      count += equal ? 1 : -1;
    }
  }
  a.winner_slot[i] = a.group_table[g * a.max_r + cand];
  if (a.winner_member) a.winner_member[i] = cand;
  if (a.clear_mask) a.clear_mask[i] = 0u;
}

extern "C" int drc_vote_compare(const VoteArgs* args, int grid, cudaStream_t stream) {
  if (args->max_r > DRC_MAX_R || args->G > DRC_MAX_WORKERS) return (int)cudaErrorInvalidValue;
  if (args->G <= 4 && args->max_r <= 4) vote_compare_kernel<true><<<grid, DRC_THREADS, 0, stream>>>(*args);
  else vote_compare_kernel<false><<<grid, DRC_THREADS, 0, stream>>>(*args);
  return (int)cudaGetLastError();
}

extern "C" int drc_vote_resolve(const ResolveArgs* args, cudaStream_t stream) {
  int n = args->G * ((args->t_end > 0 ? args->t_end : args->T) - args->t_begin);
  vote_resolve_kernel<<<(n + 127) / 128, 128, 0, stream>>>(*args);
  return (int)cudaGetLastError();
}

extern "C" int drc_sizeof_VoteArgs() { return (int)sizeof(VoteArgs); }
extern "C" int drc_sizeof_ResolveArgs() { return (int)sizeof(ResolveArgs); }
