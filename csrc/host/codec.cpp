// Lossless gradient codec, host implementation (N4 replacement).
//
// Reference: `blosc.pack_array(grad, cname='snappy')` / `blosc.unpack_array` (src/compress_gradient.py:7-15), i.e.
// c-blosc's byte-shuffle filter followed by an LZ codec.  This codec keeps the idea that makes blosc work on
// floating-point gradients -- transpose to byte planes so that the (highly redundant) sign/exponent bytes sit
// together -- and replaces the LZ stage by per-plane frame-of-reference bit packing, which is branch-free and
// maps 1:1 onto a GPU block (the device twin in csrc/cuda/codec.cu produces the identical stream):
//
//   stream  := header  block*
//   header  := 'DRC1' u32 itemsize  u64 raw_bytes  u32 block_elems
//   block   := plane[itemsize]                  (elements of the block, byte-plane p = byte p of every element)
//   plane   := u8 mode  payload
//       mode 0 RAW    payload = n bytes
//       mode 1 CONST  payload = 1 byte                      (all bytes of the plane equal)
//       mode 2 PACK   payload = u8 base, u8 bits, ceil(n*bits/8) bytes   (byte - base  <  2^bits, bits in {1,2,4})
#include <stdint.h>
#include <string.h>

namespace {
const uint32_t MAGIC = 0x31435244u;  // 'DRC1'
const uint32_t BLOCK_ELEMS = 4096;

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
  int bits = range < 2 ? 1 : range < 4 ? 2 : range < 16 ? 4 : 0;
  if (!bits) { w.put8(0); w.put(plane, n); return; }
  w.put8(2); w.put8(lo); w.put8((uint8_t)bits);
  const uint32_t per = 8 / bits;
  for (uint32_t i = 0; i < n; i += per) {
    uint8_t b = 0;
    for (uint32_t j = 0; j < per && i + j < n; ++j) b |= (uint8_t)((plane[i + j] - lo) << (j * bits));
    w.put8(b);
  }
}
}  // namespace

extern "C" {

// Worst-case size of the encoded stream for `raw_bytes` of payload.
uint64_t drc_codec_bound(uint64_t raw_bytes, uint32_t itemsize) {
  uint64_t elems = raw_bytes / (itemsize ? itemsize : 1) + 1;
  uint64_t blocks = elems / BLOCK_ELEMS + 1;
  return 20 + raw_bytes + blocks * itemsize * 3 + 16;
}

// Returns encoded size, or 0 on failure (dst too small / bad arguments).
uint64_t drc_codec_encode(const uint8_t* src, uint64_t raw_bytes, uint32_t itemsize, uint8_t* dst, uint64_t dst_cap) {
  if (itemsize == 0 || itemsize > 16 || raw_bytes % itemsize) return 0;
  Writer w{dst, dst + dst_cap, true};
  uint32_t be = BLOCK_ELEMS;
  w.put(&MAGIC, 4); w.put(&itemsize, 4); w.put(&raw_bytes, 8); w.put(&be, 4);
  const uint64_t elems = raw_bytes / itemsize;
  uint8_t plane[BLOCK_ELEMS];
  for (uint64_t e0 = 0; e0 < elems; e0 += BLOCK_ELEMS) {
    const uint32_t n = (uint32_t)((elems - e0) < BLOCK_ELEMS ? (elems - e0) : BLOCK_ELEMS);
    for (uint32_t p = 0; p < itemsize; ++p) {
      const uint8_t* s = src + e0 * itemsize + p;
      for (uint32_t i = 0; i < n; ++i) plane[i] = s[(uint64_t)i * itemsize];
      encode_plane(w, plane, n);
    }
  }
  return w.ok ? (uint64_t)(w.p - dst) : 0;
}

// Raw size recorded in a stream (0 if the header is invalid).
uint64_t drc_codec_raw_size(const uint8_t* src, uint64_t n) {
  if (n < 20) return 0;
  uint32_t magic; memcpy(&magic, src, 4);
  if (magic != MAGIC) return 0;
  uint64_t raw; memcpy(&raw, src + 8, 8);
  return raw;
}

// Returns decoded size, or 0 on failure.
uint64_t drc_codec_decode(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dst_cap) {
  if (n < 20) return 0;
  uint32_t magic, itemsize, be; uint64_t raw;
  memcpy(&magic, src, 4); memcpy(&itemsize, src + 4, 4); memcpy(&raw, src + 8, 8); memcpy(&be, src + 16, 4);
  if (magic != MAGIC || itemsize == 0 || itemsize > 16 || be != BLOCK_ELEMS || raw > dst_cap || raw % itemsize) return 0;
  const uint8_t* p = src + 20; const uint8_t* end = src + n;
  const uint64_t elems = raw / itemsize;
  uint8_t plane[BLOCK_ELEMS];
  for (uint64_t e0 = 0; e0 < elems; e0 += BLOCK_ELEMS) {
    const uint32_t cnt = (uint32_t)((elems - e0) < BLOCK_ELEMS ? (elems - e0) : BLOCK_ELEMS);
    for (uint32_t pl = 0; pl < itemsize; ++pl) {
      if (p >= end) return 0;
      const uint8_t mode = *p++;
      if (mode == 0) { if (p + cnt > end) return 0; memcpy(plane, p, cnt); p += cnt; }
      else if (mode == 1) { if (p + 1 > end) return 0; memset(plane, *p++, cnt); }
      else if (mode == 2) {
        if (p + 2 > end) return 0;
        const uint8_t base = *p++; const uint32_t bits = *p++;
        if (bits != 1 && bits != 2 && bits != 4) return 0;
        const uint32_t per = 8 / bits, nb = (cnt + per - 1) / per, mask = (1u << bits) - 1;
        if (p + nb > end) return 0;
        for (uint32_t i = 0; i < cnt; ++i) plane[i] = (uint8_t)(base + ((p[i / per] >> ((i % per) * bits)) & mask));
        p += nb;
      } else return 0;
      uint8_t* d = dst + e0 * itemsize + pl;
      for (uint32_t i = 0; i < cnt; ++i) d[(uint64_t)i * itemsize] = plane[i];
    }
  }
  return raw;
}

}  // extern "C"
