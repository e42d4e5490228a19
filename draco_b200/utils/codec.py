"""Lossless gradient (de)compression -- API parity with the reference's ``compress_gradient.py``.

Reference: ``compress(ndarray) -> bytes`` = ``blosc.pack_array(grad, cname='snappy')`` and ``decompress(bytes) -> ndarray``
(src/compress_gradient.py:7-15).  Same two functions here; the payload codec is the in-tree byte-plane /
frame-of-reference codec (host: csrc/host/codec.cpp, device twin: csrc/cuda/codec.cu, identical "DRC2" stream) and a
small header carries dtype + shape the way ``pack_array`` does.  ``compress_tensor`` / ``decompress_tensor`` run the
codec on the GPU without leaving the device.
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from .. import _native as N

_HDR = struct.Struct("<4sB3xI")     # magic, ndim, dtype-string length
STREAM_HEADER = 24
BLOCK_ELEMS = 4096


def itemsize_flags(dtype: np.dtype) -> int:
    dtype = np.dtype(dtype)
    itemsize = dtype.itemsize if dtype.itemsize <= 16 else 1
    if dtype.kind == "c":
        itemsize = dtype.itemsize // 2          # complex: shuffle on the real/imag component width
    if dtype in (np.float32, np.complex64):
        itemsize |= 0x100                       # float32 words: rotate the sign bit out of the exponent byte
    return itemsize


def compress(grad: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(grad)
    lib = N.host()
    dt = arr.dtype.str.encode()
    header = _HDR.pack(b"DRCA", arr.ndim, len(dt)) + dt + struct.pack(f"<{arr.ndim}q", *arr.shape)
    flags = itemsize_flags(arr.dtype)
    cap = int(lib.drc_codec_bound(arr.nbytes, flags))
    out = (C.c_uint8 * cap)()
    n = int(lib.drc_codec_encode(arr.ctypes.data if arr.nbytes else None, arr.nbytes, flags, out, cap))
    if n == 0:
        raise RuntimeError("codec encode failed")
    return header + bytes(memoryview(out)[:n])


def decompress(msg: bytes) -> np.ndarray:
    lib = N.host()
    if len(msg) < _HDR.size:
        raise ValueError("not a draco_b200 compressed array")
    magic, ndim, dtlen = _HDR.unpack_from(msg, 0)
    if magic != b"DRCA":
        raise ValueError("not a draco_b200 compressed array")
    off = _HDR.size
    dtype = np.dtype(msg[off: off + dtlen].decode())
    off += dtlen
    shape = struct.unpack_from(f"<{ndim}q", msg, off)
    off += 8 * ndim
    payload = (C.c_uint8 * (len(msg) - off)).from_buffer_copy(msg, off)
    if not lib.drc_codec_valid(payload, len(msg) - off):
        raise ValueError("corrupt codec stream")
    raw = int(lib.drc_codec_raw_size(payload, len(msg) - off))
    out = np.empty(raw, dtype=np.uint8)
    n = int(lib.drc_codec_decode(payload, len(msg) - off, out.ctypes.data if raw else None, raw))
    if n != raw:
        raise RuntimeError("codec decode failed")
    return out.view(dtype).reshape(shape)


def ratio(grad: np.ndarray) -> float:
    """Compressed size / raw size."""
    return len(compress(grad)) / max(np.asarray(grad).nbytes, 1)


# ---------------------------------------------------------------------------------------------------------------------
# device codec
# ---------------------------------------------------------------------------------------------------------------------
def _cuda_lib():
    lib = N.cuda()
    if not getattr(lib, "_codec_ready", False):
        lib.drc_codec_plan.argtypes = [N.ptr, N.i64, C.c_int, N.ptr, N.ptr, N.ptr]
        lib.drc_codec_plan.restype = C.c_int
        lib.drc_codec_pack.argtypes = [N.ptr, N.i64, C.c_int, N.ptr, N.ptr, N.ptr, N.ptr]
        lib.drc_codec_pack.restype = C.c_int
        lib.drc_codec_unpack.argtypes = [N.ptr, N.ptr, N.i64, C.c_int, N.ptr, N.ptr, N.ptr]
        lib.drc_codec_unpack.restype = C.c_int
        lib._codec_ready = True
    return lib


def _torch_flags(t) -> int:
    import torch
    if t.dtype in (torch.float32, torch.complex64):
        return 4 | 0x100
    if t.dtype == torch.complex128:
        return 8
    return t.element_size()


def compress_tensor(t) -> "torch.Tensor":
    """Compress a contiguous CUDA tensor on the device.  Returns a uint8 CUDA tensor holding the DRC2 stream (its length
    is the compressed size; one host sync to learn it)."""
    import torch
    assert t.is_cuda and t.is_contiguous()
    lib = _cuda_lib()
    flags = _torch_flags(t)
    raw = t.numel() * t.element_size()
    itemsize = flags & 0xff
    elems = raw // itemsize
    nblocks = (elems + BLOCK_ELEMS - 1) // BLOCK_ELEMS
    st = torch.cuda.current_stream().cuda_stream
    meta = torch.empty(max(nblocks, 1) * 16, dtype=torch.int32, device=t.device)
    sizes = torch.empty(max(nblocks, 1), dtype=torch.int32, device=t.device)
    N.check(lib.drc_codec_plan(t.data_ptr(), raw, flags, meta.data_ptr(), sizes.data_ptr(), st), "codec_plan")
    csum = torch.cumsum(sizes[:nblocks].to(torch.int64), 0) if nblocks else torch.zeros(0, dtype=torch.int64, device=t.device)
    base = STREAM_HEADER + 4 * nblocks
    offs = (torch.cat([csum.new_zeros(1), csum[:-1]]) + base) if nblocks else csum
    total = base + (int(csum[-1].item()) if nblocks else 0)
    out = torch.empty(total, dtype=torch.uint8, device=t.device)
    hdr = struct.pack("<IIQII", 0x32435244, flags, raw, BLOCK_ELEMS, nblocks)
    out[:STREAM_HEADER].copy_(torch.frombuffer(bytearray(hdr), dtype=torch.uint8))
    if nblocks:
        out[STREAM_HEADER:base].copy_(sizes[:nblocks].view(torch.uint8))
        N.check(lib.drc_codec_pack(t.data_ptr(), raw, flags, meta.data_ptr(), offs.data_ptr(), out.data_ptr(), st), "codec_pack")
    return out


def decompress_tensor(stream, dtype, shape) -> "torch.Tensor":
    """Decode a DRC2 stream held in a uint8 CUDA tensor into a new tensor of ``dtype`` / ``shape`` on the same device."""
    import torch
    assert stream.is_cuda and stream.dtype == torch.uint8
    lib = _cuda_lib()
    hdr = bytes(stream[:STREAM_HEADER].cpu().numpy())
    magic, flags, raw, be, nblocks = struct.unpack("<IIQII", hdr)
    if magic != 0x32435244 or be != BLOCK_ELEMS:
        raise ValueError("corrupt codec stream")
    out = torch.empty(shape, dtype=dtype, device=stream.device)
    if out.numel() * out.element_size() != raw:
        raise ValueError("shape/dtype do not match the stream")
    if nblocks == 0:
        return out
    base = STREAM_HEADER + 4 * nblocks
    sizes = stream[STREAM_HEADER:base].view(torch.int32).to(torch.int64)
    csum = torch.cumsum(sizes, 0)
    offs = torch.cat([csum.new_zeros(1), csum[:-1]]) + base
    err = torch.zeros(1, dtype=torch.int32, device=stream.device)
    N.check(lib.drc_codec_unpack(stream.data_ptr(), offs.data_ptr(), raw, flags, out.data_ptr(), err.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream), "codec_unpack")
    if int(err.item()):
        raise ValueError("corrupt codec stream")
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Sync-free device path (fused transport): fixed worst-case buffers, every size stays on the device -> capturable in a graph
# ---------------------------------------------------------------------------------------------------------------------
class DeviceStreamCodec:
    """Packs / unpacks a flat fp32 buffer of ``nfloats`` elements without any host synchronisation.

    ``pack(src, out)`` writes a DRC2 stream into ``out`` (uint8, ``capacity`` bytes) and returns the device scalar holding its
    length; ``unpack(stream, dst)`` decodes a stream (its index is read on the device).  Header and block count are static."""

    def __init__(self, nfloats: int, device):
        import torch
        self.lib = _cuda_lib()
        self.device = device
        self.raw = nfloats * 4
        self.flags = 4 | 0x100                                   # 32-bit words, rotated (sign bit leaves the exponent byte)
        self.nblocks = (nfloats + BLOCK_ELEMS - 1) // BLOCK_ELEMS
        self.base = STREAM_HEADER + 4 * self.nblocks
        # worst case: every plane RAW = 1 mode byte + n bytes per plane, 4 planes per block
        self.capacity = ((self.base + self.raw + 4 * self.nblocks + 15) // 16 + 1) * 16
        hdr = struct.pack("<IIQII", 0x32435244, self.flags, self.raw, BLOCK_ELEMS, self.nblocks)
        self.header = torch.frombuffer(bytearray(hdr), dtype=torch.uint8).to(device)
        self.meta = torch.empty(max(self.nblocks, 1) * 16, dtype=torch.int32, device=device)
        self.sizes = torch.empty(max(self.nblocks, 1), dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)

    def pack(self, src, out):
        import torch
        st = torch.cuda.current_stream().cuda_stream
        N.check(self.lib.drc_codec_plan(src.data_ptr(), self.raw, self.flags, self.meta.data_ptr(), self.sizes.data_ptr(), st), "codec_plan")
        csum = torch.cumsum(self.sizes.to(torch.int64), 0)
        offs = torch.cat([csum.new_zeros(1), csum[:-1]]) + self.base
        out[:STREAM_HEADER].copy_(self.header)
        out[STREAM_HEADER:self.base].copy_(self.sizes.view(torch.uint8))
        N.check(self.lib.drc_codec_pack(src.data_ptr(), self.raw, self.flags, self.meta.data_ptr(), offs.data_ptr(), out.data_ptr(), st),
                "codec_pack")
        return (csum[-1:] + self.base).contiguous()              # int64 [1] on the device: packed bytes

    def unpack(self, stream, dst) -> None:
        import torch
        sizes = stream[STREAM_HEADER:self.base].view(torch.int32).to(torch.int64)
        csum = torch.cumsum(sizes, 0)
        offs = torch.cat([csum.new_zeros(1), csum[:-1]]) + self.base
        N.check(self.lib.drc_codec_unpack(stream.data_ptr(), offs.data_ptr(), self.raw, self.flags, dst.data_ptr(), self.err.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "codec_unpack")
