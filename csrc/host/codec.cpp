// Lossless gradient codec, host implementation (N4 replacement); csrc/cuda/codec.cu is the bit-identical device twin.
//
// Reference: `blosc.pack_array(grad, cname='snappy')` / `blosc.unpack_array` (src/compress_gradient.py:7-15), i.e.
// c-blosc's byte-shuffle filter followed by an LZ codec.  This codec keeps the idea that makes blosc work on
// floating-point gradients -- transpose to byte planes so that the (highly redundant) sign/exponent bytes sit
// together -- and replaces the LZ stage by per-plane frame-of-reference bit packing, which is branch-free and maps
// 1:1 onto a GPU thread block.
//
// Stream format "DRC2":
//   header  : 'DRC2' u32 | itemsize_flags u32 | raw_bytes u64 | block_elems u32 (=4096) | nblocks u32          (24 B)
//   index   : u32 block_bytes[nblocks]                       (compressed size of each block -> parallel decode)
//   block   : plane[itemsize]       (byte-plane p = byte p of every element of the block)
//   plane   : u8 mode, payload      mode 0 RAW n bytes | 1 CONST 1 byte | 2 PACK u8 base, u8 bits (1..7), ceil(n*bits/8) bytes
//   itemsize_flags bit 8 (0x100) = "float32 words": each 32-bit word is rotated left by one before the byte split so that
//   the sign bit leaves the exponent byte (the exponent plane of a gradient tensor then spans a few values only).
#include <stdint.h>
#include <string.h>

namespace {
const uint32_t MAGIC = 0x32435244u;  // 'DRC2'
const uint32_t BLOCK_ELEMS = 4096;
const uint32_t HEADER = 24;

struct Writer {
  uint8_t* p; uint8_t* end; bool ok;
  void put(const void* src, size_t n) { if (!ok || p + n > end) { ok = false; return; } memcpy(p, src, n); p += n; }
  void put8(uint8_t v) { put(&v, 1); }
};

void encode_plane(Writer& w, const uint8_t* plane, uint32_t n) {
  uint8_t lo = 255, hi = 0;
  for (uint32_t i = 0; i < n; ++i) { if (plane[i] < lo) lo = plane[i]; if (plane[i] > hi) hi = plane[i]; }
  const uint32_t range = (uint32_t)hi - lo;
  if (range == 0) { w.put8(1); w.put8(lo); return; }
  int bits = 0;
  while ((1u << bits) <= range) ++bits;            // smallest width with range < 2^bits
  if (bits >= 8) { w.put8(0); w.put(plane, n); return; }
  w.put8(2); w.put8(lo); w.put8((uint8_t)bits);
  uint32_t acc = 0; int fill = 0;                  // little-endian bitstream
  for (uint32_t i = 0; i < n; ++i) {
    acc |= (uint32_t)(plane[i] - lo) << fill; fill += bits;
    while (fill >= 8) { w.put8((uint8_t)(acc & 0xff)); acc >>= 8; fill -= 8; }
  }
  if (fill > 0) w.put8((uint8_t)(acc & 0xff));
}

bool parse_header(const uint8_t* src, uint64_t n, uint32_t& itemsize, bool& rot, uint64_t& raw, uint32_t& nblocks) {
  if (n < HEADER) return false;
  uint32_t magic, flags, be;
  memcpy(&magic, src, 4); memcpy(&flags, src + 4, 4); memcpy(&raw, src + 8, 8); memcpy(&be, src + 16, 4); memcpy(&nblocks, src + 20, 4);
  itemsize = flags & 0xff; rot = (flags & 0x100) != 0;
  if (magic != MAGIC || itemsize == 0 || itemsize > 16 || be != BLOCK_ELEMS || raw % itemsize || (rot && itemsize != 4)) return false;
  const uint64_t elems = raw / itemsize;
  if (nblocks != (uint32_t)((elems + BLOCK_ELEMS - 1) / BLOCK_ELEMS)) return false;
  return n >= HEADER + 4ull * nblocks;
}
}  // namespace

extern "C" {

// Worst-case size of the encoded stream for `raw_bytes` of payload.
uint64_t drc_codec_bound(uint64_t raw_bytes, uint32_t itemsize) {
  itemsize &= 0xff;
  if (itemsize == 0) itemsize = 1;
  uint64_t elems = raw_bytes / itemsize + 1;
  uint64_t blocks = elems / BLOCK_ELEMS + 1;
  return HEADER + 4 * blocks + raw_bytes + blocks * itemsize * 3 + 16;
}

// Returns encoded size, or 0 on failure (dst too small / bad arguments).
uint64_t drc_codec_encode(const uint8_t* src, uint64_t raw_bytes, uint32_t itemsize_flags, uint8_t* dst, uint64_t dst_cap) {
  const uint32_t itemsize = itemsize_flags & 0xff;
  const bool rot = (itemsize_flags & 0x100) != 0;
  if (itemsize == 0 || itemsize > 16 || raw_bytes % itemsize || (rot && itemsize != 4)) return 0;
  const uint64_t elems = raw_bytes / itemsize;
  const uint32_t nblocks = (uint32_t)((elems + BLOCK_ELEMS - 1) / BLOCK_ELEMS);
  if (dst_cap < HEADER + 4ull * nblocks) return 0;
  uint32_t be = BLOCK_ELEMS;
  memcpy(dst, &MAGIC, 4); memcpy(dst + 4, &itemsize_flags, 4); memcpy(dst + 8, &raw_bytes, 8); memcpy(dst + 16, &be, 4);
  memcpy(dst + 20, &nblocks, 4);
  uint8_t* index = dst + HEADER;
  Writer w{dst + HEADER + 4ull * nblocks, dst + dst_cap, true};
  uint8_t plane[BLOCK_ELEMS];
  for (uint32_t b = 0; b < nblocks; ++b) {
    const uint64_t e0 = (uint64_t)b * BLOCK_ELEMS;
    const uint32_t n = (uint32_t)((elems - e0) < BLOCK_ELEMS ? (elems - e0) : BLOCK_ELEMS);
    const uint8_t* start = w.p;
    for (uint32_t p = 0; p < itemsize; ++p) {
      if (!rot) {
        const uint8_t* s = src + e0 * itemsize + p;
        for (uint32_t i = 0; i < n; ++i) plane[i] = s[(uint64_t)i * itemsize];
      } else {
        for (uint32_t i = 0; i < n; ++i) {
          uint32_t v; memcpy(&v, src + (e0 + i) * 4, 4);
          v = (v << 1) | (v >> 31);
          plane[i] = (uint8_t)(v >> (8 * p));
        }
      }
      encode_plane(w, plane, n);
    }
    if (!w.ok) return 0;
    const uint32_t bytes = (uint32_t)(w.p - start);
    memcpy(index + 4ull * b, &bytes, 4);
  }
  return w.ok ? (uint64_t)(w.p - dst) : 0;
}

// Raw size recorded in a stream (0 if the header is invalid).
uint64_t drc_codec_raw_size(const uint8_t* src, uint64_t n) {
  uint32_t itemsize, nblocks; bool rot; uint64_t raw;
  return parse_header(src, n, itemsize, rot, raw, nblocks) ? raw : 0;
}

// itemsize_flags recorded in a stream (0 if invalid).
uint32_t drc_codec_itemsize_flags(const uint8_t* src, uint64_t n) {
  uint32_t itemsize, nblocks; bool rot; uint64_t raw;
  if (!parse_header(src, n, itemsize, rot, raw, nblocks)) return 0;
  return itemsize | (rot ? 0x100u : 0u);
}

// Returns 1 if the stream header is valid.
int drc_codec_valid(const uint8_t* src, uint64_t n) {
  uint32_t itemsize, nblocks; bool rot; uint64_t raw;
  return parse_header(src, n, itemsize, rot, raw, nblocks) ? 1 : 0;
}

// Returns decoded size; -1 (as u64 max) on failure.
uint64_t drc_codec_decode(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dst_cap) {
  uint32_t itemsize, nblocks; bool rot; uint64_t raw;
  if (!parse_header(src, n, itemsize, rot, raw, nblocks) || raw > dst_cap) return ~0ull;
  const uint8_t* p = src + HEADER + 4ull * nblocks; const uint8_t* end = src + n;
  const uint64_t elems = raw / itemsize;
  uint8_t plane[BLOCK_ELEMS];
  for (uint64_t e0 = 0; e0 < elems; e0 += BLOCK_ELEMS) {
    const uint32_t cnt = (uint32_t)((elems - e0) < BLOCK_ELEMS ? (elems - e0) : BLOCK_ELEMS);
    for (uint32_t pl = 0; pl < itemsize; ++pl) {
      if (p >= end) return ~0ull;
      const uint8_t mode = *p++;
      if (mode == 0) { if (p + cnt > end) return ~0ull; memcpy(plane, p, cnt); p += cnt; }
      else if (mode == 1) { if (p + 1 > end) return ~0ull; memset(plane, *p++, cnt); }
      else if (mode == 2) {
        if (p + 2 > end) return ~0ull;
        const uint8_t base = *p++; const uint32_t bits = *p++;
        if (bits < 1 || bits > 7) return ~0ull;
        const uint32_t nb = (cnt * bits + 7) / 8, mask = (1u << bits) - 1;
        if (p + nb > end) return ~0ull;
        for (uint32_t i = 0; i < cnt; ++i) {
          const uint32_t bit = i * bits, byte = bit >> 3, sh = bit & 7;
          uint32_t word = p[byte] | (byte + 1 < nb ? (uint32_t)p[byte + 1] << 8 : 0u);
          plane[i] = (uint8_t)(base + ((word >> sh) & mask));
        }
        p += nb;
      } else return ~0ull;
      uint8_t* d = dst + e0 * itemsize + pl;
      for (uint32_t i = 0; i < cnt; ++i) d[(uint64_t)i * itemsize] = plane[i];
    }
    if (rot) {
      for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t v; memcpy(&v, dst + (e0 + i) * 4, 4);
        v = (v >> 1) | (v << 31);
        memcpy(dst + (e0 + i) * 4, &v, 4);
      }
    }
  }
  return raw;
}

}  // extern "C"
