// K12 / N4: lossless gradient codec on the device -- the GPU twin of csrc/host/codec.cpp (bit-identical stream, so a
// tensor compressed on the GPU can be decoded on the host and vice versa).
//
// Reference: every gradient tensor goes through `blosc.pack_array(..., cname='snappy')` on the worker and
// `blosc.unpack_array` on the PS (src/compress_gradient.py:7-15).  On NVLink the codec can only lose time against a
// 770 GB/s link, but it is a capability of the reference (--compress-grad compress is its default), so every transport
// honours the flag: on the fused transport the worker encodes into a local buffer, packs it, and stream_push_kernel below
// stores only the packed bytes into the PS's staging slot over NVLink (the byte count is read from device memory, so the whole
// path stays inside the captured graph); the PS unpacks straight into the worker's gradient slot ahead of the decode.
//
// Stream format "DRC2":
//   header  : 'DRC2' u32 | itemsize_flags u32 | raw_bytes u64 | block_elems u32 (=4096) | nblocks u32          (24 B)
//   index   : u32 block_bytes[nblocks]                       (makes decode embarrassingly parallel)
//   block   : plane[itemsize], plane := u8 mode, payload     mode 0 RAW n bytes | 1 CONST 1 byte | 2 PACK base,bits,ceil(n*bits/8) bytes
//   itemsize_flags bit 8: 32-bit words are rotated left by one first (sign bit leaves the exponent byte).
//
// One CTA per 4096-element block; the elements are staged once in shared memory, planes are analysed with a CTA-wide
// min/max, and bit-packing works on groups of 8 elements (= `bits` whole bytes), so no two threads share an output byte.
#include "common.cuh"

namespace {

constexpr int CB_ELEMS = 4096;
constexpr int CB_THREADS = 256;
constexpr int MAX_ITEM = 16;

__device__ __forceinline__ uint8_t plane_byte(const uint8_t* s_raw, int i, int p, int itemsize, bool rot) {
  if (!rot) return s_raw[i * itemsize + p];
  uint32_t v = reinterpret_cast<const uint32_t*>(s_raw)[i];
  v = (v << 1) | (v >> 31);
  return (uint8_t)(v >> (8 * p));
}

__device__ __forceinline__ int bits_for_range(int range) {
  int bits = 0;
  while ((1 << bits) <= range) ++bits;
  return bits;                                    // 8 means "not packable"
}

__device__ __forceinline__ uint32_t plane_size(int mode, int bits, int n) {
  if (mode == 1) return 2;
  if (mode == 2) return 3 + (uint32_t)((n * bits + 7) / 8);
  return 1 + (uint32_t)n;
}

// CTA-wide min / max of one byte plane
__device__ __forceinline__ void plane_minmax(const uint8_t* s_raw, int n, int p, int itemsize, bool rot, int& lo, int& hi, int* s_red) {
  int l = 255, h = 0;
  for (int i = threadIdx.x; i < n; i += CB_THREADS) {
    int b = plane_byte(s_raw, i, p, itemsize, rot);
    l = min(l, b); h = max(h, b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { l = min(l, __shfl_xor_sync(0xffffffffu, l, o)); h = max(h, __shfl_xor_sync(0xffffffffu, h, o)); }
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = l; s_red[8 + (threadIdx.x >> 5)] = h; }
  __syncthreads();
  l = 255; h = 0;
#pragma unroll
  for (int w = 0; w < CB_THREADS / 32; ++w) { l = min(l, s_red[w]); h = max(h, s_red[8 + w]); }
  __syncthreads();
  lo = l; hi = h;
}

__device__ __forceinline__ void stage_block(uint8_t* s_raw, const uint8_t* src, long long e0, int n, int itemsize) {
  const int nbytes = n * itemsize;
  const uint8_t* g = src + e0 * itemsize;
  if ((reinterpret_cast<uintptr_t>(g) & 3) == 0) {
    for (int i = threadIdx.x; i < nbytes / 4; i += CB_THREADS) reinterpret_cast<uint32_t*>(s_raw)[i] = reinterpret_cast<const uint32_t*>(g)[i];
    for (int i = (nbytes / 4) * 4 + threadIdx.x; i < nbytes; i += CB_THREADS) s_raw[i] = g[i];
  } else {
    for (int i = threadIdx.x; i < nbytes; i += CB_THREADS) s_raw[i] = g[i];
  }
  __syncthreads();
}

// pass 1: per-block plane decisions and compressed size
__global__ void __launch_bounds__(CB_THREADS) codec_plan_kernel(const uint8_t* src, long long elems, int itemsize, int rot,
                                                                uint32_t* plane_meta, uint32_t* block_bytes) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  __shared__ int s_red[16];
  const long long b = blockIdx.x;
  const long long e0 = b * CB_ELEMS;
  const int n = (int)((elems - e0) < CB_ELEMS ? (elems - e0) : CB_ELEMS);
  stage_block(s_raw, src, e0, n, itemsize);
  uint32_t total = 0;
  for (int p = 0; p < itemsize; ++p) {
    int lo, hi;
    plane_minmax(s_raw, n, p, itemsize, rot != 0, lo, hi, s_red);
    const int range = hi - lo;
    int mode, bits = 0;
    if (range == 0) mode = 1;
    else { bits = bits_for_range(range); mode = bits >= 8 ? 0 : 2; }
    if (threadIdx.x == 0) plane_meta[b * MAX_ITEM + p] = (uint32_t)mode | ((uint32_t)lo << 8) | ((uint32_t)bits << 16);
    total += plane_size(mode, bits, n);
  }
  if (threadIdx.x == 0) block_bytes[b] = total;
}

// pass 2: write the planes of each block at its offset
__global__ void __launch_bounds__(CB_THREADS) codec_pack_kernel(const uint8_t* src, long long elems, int itemsize, int rot,
                                                                const uint32_t* plane_meta, const long long* block_off, uint8_t* dst) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  const long long b = blockIdx.x;
  const long long e0 = b * CB_ELEMS;
  const int n = (int)((elems - e0) < CB_ELEMS ? (elems - e0) : CB_ELEMS);
  stage_block(s_raw, src, e0, n, itemsize);
  uint8_t* out = dst + block_off[b];
  for (int p = 0; p < itemsize; ++p) {
    const uint32_t m = plane_meta[b * MAX_ITEM + p];
    const int mode = m & 0xff, base = (m >> 8) & 0xff, bits = (m >> 16) & 0xff;
    if (threadIdx.x == 0) {
      out[0] = (uint8_t)mode;
      if (mode == 1) out[1] = (uint8_t)base;
      if (mode == 2) { out[1] = (uint8_t)base; out[2] = (uint8_t)bits; }
    }
    if (mode == 0) {
      for (int i = threadIdx.x; i < n; i += CB_THREADS) out[1 + i] = plane_byte(s_raw, i, p, itemsize, rot != 0);
    } else if (mode == 2) {
      const int groups = (n + 7) / 8;
      for (int g = threadIdx.x; g < groups; g += CB_THREADS) {
        unsigned long long acc = 0;
        const int cnt = min(8, n - g * 8);
        for (int j = 0; j < cnt; ++j)
          acc |= (unsigned long long)(plane_byte(s_raw, g * 8 + j, p, itemsize, rot != 0) - base) << (j * bits);
        const int nb = (cnt * bits + 7) / 8;
        for (int k = 0; k < nb; ++k) out[3 + g * bits + k] = (uint8_t)(acc >> (8 * k));
      }
    }
    out += plane_size(mode, bits, n);
  }
}

// decode: one CTA per block
__global__ void __launch_bounds__(CB_THREADS) codec_unpack_kernel(const uint8_t* stream, const long long* block_off, long long elems,
                                                                  int itemsize, int rot, uint8_t* dst, int* error) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  __shared__ uint32_t s_plane_off[MAX_ITEM + 1];
  __shared__ uint32_t s_plane_meta[MAX_ITEM];
  const long long b = blockIdx.x;
  const long long e0 = b * CB_ELEMS;
  const int n = (int)((elems - e0) < CB_ELEMS ? (elems - e0) : CB_ELEMS);
  const uint8_t* in = stream + block_off[b];
  if (threadIdx.x == 0) {
    uint32_t off = 0;
    for (int p = 0; p < itemsize; ++p) {
      const int mode = in[off];
      int base = 0, bits = 0;
      if (mode == 1) base = in[off + 1];
      else if (mode == 2) { base = in[off + 1]; bits = in[off + 2]; if (bits < 1 || bits > 7) atomicExch(error, 2); }
      else if (mode != 0) atomicExch(error, 1);
      s_plane_off[p] = off;
      s_plane_meta[p] = (uint32_t)mode | ((uint32_t)base << 8) | ((uint32_t)bits << 16);
      off += plane_size(mode > 2 ? 0 : mode, bits, n);
    }
    s_plane_off[itemsize] = off;
  }
  __syncthreads();
  for (int p = 0; p < itemsize; ++p) {
    const uint32_t m = s_plane_meta[p];
    const int mode = m & 0xff, base = (m >> 8) & 0xff, bits = (m >> 16) & 0xff;
    const uint8_t* pl = in + s_plane_off[p];
    for (int i = threadIdx.x; i < n; i += CB_THREADS) {
      uint8_t v;
      if (mode == 1) v = (uint8_t)base;
      else if (mode == 2) {
        const int bit = i * bits, byte = bit >> 3, sh = bit & 7;
        const int nb = (n * bits + 7) / 8;
        uint32_t word = pl[3 + byte] | (byte + 1 < nb ? (uint32_t)pl[3 + byte + 1] << 8 : 0u);
        v = (uint8_t)(base + ((word >> sh) & ((1u << bits) - 1)));
      } else v = pl[1 + i];
      s_raw[i * itemsize + p] = v;
    }
  }
  __syncthreads();
  uint8_t* g = dst + e0 * itemsize;
  if (rot) {
    for (int i = threadIdx.x; i < n; i += CB_THREADS) {
      uint32_t v = reinterpret_cast<const uint32_t*>(s_raw)[i];
      v = (v >> 1) | (v << 31);
      reinterpret_cast<uint32_t*>(g)[i] = v;                       // rot implies itemsize 4 and a 4-byte aligned dst
    }
  } else {
    for (int i = threadIdx.x; i < n * itemsize; i += CB_THREADS) g[i] = s_raw[i];
  }
}

}  // namespace

extern "C" {

// plane_meta: u32 [nblocks][16]; block_bytes: u32 [nblocks]
int drc_codec_plan(const void* src, long long raw_bytes, int itemsize_flags, uint32_t* plane_meta, uint32_t* block_bytes,
                   cudaStream_t stream) {
  const int itemsize = itemsize_flags & 0xff, rot = (itemsize_flags >> 8) & 1;
  if (itemsize < 1 || itemsize > MAX_ITEM || raw_bytes % itemsize || (rot && itemsize != 4)) return -1;
  const long long elems = raw_bytes / itemsize;
  const int nblocks = (int)((elems + CB_ELEMS - 1) / CB_ELEMS);
  if (nblocks == 0) return 0;
  auto k = codec_plan_kernel;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_ELEMS * MAX_ITEM);
  k<<<nblocks, CB_THREADS, CB_ELEMS * itemsize, stream>>>((const uint8_t*)src, elems, itemsize, rot, plane_meta, block_bytes);
  return (int)cudaGetLastError();
}

// block_off: i64 [nblocks] byte offset of each block inside `dst` (header and index included by the caller)
int drc_codec_pack(const void* src, long long raw_bytes, int itemsize_flags, const uint32_t* plane_meta, const long long* block_off,
                   void* dst, cudaStream_t stream) {
  const int itemsize = itemsize_flags & 0xff, rot = (itemsize_flags >> 8) & 1;
  const long long elems = raw_bytes / itemsize;
  const int nblocks = (int)((elems + CB_ELEMS - 1) / CB_ELEMS);
  if (nblocks == 0) return 0;
  auto k = codec_pack_kernel;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_ELEMS * MAX_ITEM);
  k<<<nblocks, CB_THREADS, CB_ELEMS * itemsize, stream>>>((const uint8_t*)src, elems, itemsize, rot, plane_meta, block_off, (uint8_t*)dst);
  return (int)cudaGetLastError();
}

int drc_codec_unpack(const void* stream_bytes, const long long* block_off, long long raw_bytes, int itemsize_flags, void* dst,
                     int* error, cudaStream_t stream) {
  const int itemsize = itemsize_flags & 0xff, rot = (itemsize_flags >> 8) & 1;
  if (itemsize < 1 || itemsize > MAX_ITEM || raw_bytes % itemsize || (rot && itemsize != 4)) return -1;
  const long long elems = raw_bytes / itemsize;
  const int nblocks = (int)((elems + CB_ELEMS - 1) / CB_ELEMS);
  if (nblocks == 0) return 0;
  auto k = codec_unpack_kernel;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_ELEMS * MAX_ITEM);
  k<<<nblocks, CB_THREADS, CB_ELEMS * itemsize, stream>>>((const uint8_t*)stream_bytes, block_off, elems, itemsize, rot, (uint8_t*)dst, error);
  return (int)cudaGetLastError();
}

// Push `*nbytes` bytes of a packed stream into a peer buffer (16-byte stores; both buffers are 16-byte aligned and padded) and raise
// the step-stamped flag from the last CTA -- the compressed twin of push_encode's store path.
struct StreamPushArgs {
  const uint4* src;
  uint4* dst;                          // peer pointer (PS staging slot of this worker)
  const long long* nbytes;             // device scalar: packed size
  long long* nbytes_out;               // optional peer word receiving the size (diagnostics / byte accounting)
  const unsigned long long* step_ptr;
  unsigned int* done_counter;
  unsigned long long* flag;
};

__global__ void __launch_bounds__(DRC_THREADS) stream_push_kernel(const __grid_constant__ StreamPushArgs a) {
  const long long n16 = (*a.nbytes + 15) >> 4;
  for (long long i = (long long)blockIdx.x * DRC_THREADS + threadIdx.x; i < n16; i += (long long)gridDim.x * DRC_THREADS) {
    const uint4 v = a.src[i];
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(a.dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.nbytes_out) *a.nbytes_out = *a.nbytes;
  if (grid_last_cta(a.done_counter)) {
    if (threadIdx.x == 0 && a.flag) st_release_sys(a.flag, *a.step_ptr);
  }
}

int drc_stream_push(const StreamPushArgs* args, int grid, cudaStream_t stream) {
  stream_push_kernel<<<grid, DRC_THREADS, 0, stream>>>(*args);
  return (int)cudaGetLastError();
}
int drc_sizeof_StreamPushArgs() { return (int)sizeof(StreamPushArgs); }

}  // extern "C"
