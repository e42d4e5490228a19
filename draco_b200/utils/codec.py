"""Lossless gradient (de)compression -- API parity with the reference's ``compress_gradient.py``.

Reference: ``compress(ndarray) -> bytes`` = ``blosc.pack_array(grad, cname='snappy')`` and ``decompress(bytes) -> ndarray``
(src/compress_gradient.py:7-15).  Same two functions here; the payload codec is the in-tree C++ byte-plane /
frame-of-reference codec (csrc/host/codec.cpp) and a small header carries dtype + shape the way ``pack_array`` does.
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from .. import _native as N

_HDR = struct.Struct("<4sB3xI")     # magic, ndim, dtype-string length


def compress(grad: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(grad)
    lib = N.host()
    dt = arr.dtype.str.encode()
    header = _HDR.pack(b"DRCA", arr.ndim, len(dt)) + dt + struct.pack(f"<{arr.ndim}q", *arr.shape)
    itemsize = arr.dtype.itemsize if arr.dtype.itemsize <= 16 else 1
    if arr.dtype.kind == "c":
        itemsize = arr.dtype.itemsize // 2          # complex: shuffle on the real/imag component width
    if arr.dtype in (np.float32, np.complex64):
        itemsize |= 0x100                           # float32 words: rotate the sign bit out of the exponent byte
    cap = int(lib.drc_codec_bound(arr.nbytes, itemsize))
    out = (C.c_uint8 * cap)()
    n = int(lib.drc_codec_encode(arr.ctypes.data if arr.nbytes else None, arr.nbytes, itemsize, out, cap))
    if n == 0:
        raise RuntimeError("codec encode failed")
    return header + bytes(memoryview(out)[:n])


def decompress(msg: bytes) -> np.ndarray:
    lib = N.host()
    magic, ndim, dtlen = _HDR.unpack_from(msg, 0)
    if magic != b"DRCA":
        raise ValueError("not a draco_b200 compressed array")
    off = _HDR.size
    dtype = np.dtype(msg[off: off + dtlen].decode())
    off += dtlen
    shape = struct.unpack_from(f"<{ndim}q", msg, off)
    off += 8 * ndim
    payload = (C.c_uint8 * (len(msg) - off)).from_buffer_copy(msg, off)
    raw = int(lib.drc_codec_raw_size(payload, len(msg) - off))
    out = np.empty(raw, dtype=np.uint8)
    n = int(lib.drc_codec_decode(payload, len(msg) - off, out.ctypes.data if raw else None, raw))
    if n != raw:
        raise RuntimeError("codec decode failed")
    return out.view(dtype).reshape(shape)


def ratio(grad: np.ndarray) -> float:
    """Compressed size / raw size."""
    return len(compress(grad)) / max(np.asarray(grad).nbytes, 1)
