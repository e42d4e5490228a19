"""Python face of the native input pipeline (``csrc/host/dataload.cpp``).

``NativeLoader`` owns a pool of sleeping C++ worker threads.  ``submit`` enqueues one sub-batch -- dataset indices plus,
optionally, the CIFAR augmentation draws (reflect-pad-4 random crop + horizontal flip) -- whose pixels and labels are
written straight into caller-provided staging memory (the pinned buffers of ``parallel/worker.py``); ``wait`` blocks until
everything submitted so far has landed.  The Python thread holds no GIL-bound work in between, so the gather of step k+1
overlaps the enqueue of step k.

Reference counterpart: src/data_loader_ops/my_data_loader.py (torch-0.3 DataLoader copy with ``next_batch``) and the
torchvision transform stack of src/util.py:37-52.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from .. import _native as N


def augment_draws(n: int, seed: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Crop offsets in [0, 8] and flip bits for ``n`` samples; the same stream as ``data.augment_cifar`` so that the Python
    and the native paths produce identical pixels for a given seed."""
    g = torch.Generator().manual_seed(int(seed) & 0x7FFFFFFF)
    dx = torch.randint(0, 9, (n,), generator=g)
    dy = torch.randint(0, 9, (n,), generator=g)
    flip = torch.rand(n, generator=g) < 0.5
    return (dx.to(torch.int32).numpy(), dy.to(torch.int32).numpy(), flip.to(torch.uint8).numpy())


def _lib():
    lib = N.host()
    if not getattr(lib, "_loader_ready", False):
        p, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        lib.drc_host_gather_augment.argtypes = [p, p, i64, i32, i32, i32, p, i32, p, p, p, i32, p, p]
        lib.drc_host_gather_augment.restype = i32
        lib.drc_loader_create.argtypes = [p, p, i64, i32, i32, i32, i32]
        lib.drc_loader_create.restype = p
        lib.drc_loader_submit.argtypes = [p, p, i32, p, p, p, i32, p, p]
        lib.drc_loader_submit.restype = i32
        lib.drc_loader_wait.argtypes = [p, i32]
        lib.drc_loader_wait.restype = i32
        lib.drc_loader_destroy.argtypes = [p]
        lib.drc_loader_destroy.restype = None
        lib._loader_ready = True
    return lib


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def gather_augment(images: torch.Tensor, labels: Optional[torch.Tensor], idx: np.ndarray, out_images: torch.Tensor,
                   out_labels: Optional[torch.Tensor] = None, seed: Optional[int] = None, pad: int = 4) -> None:
    """Synchronous one-call form: ``out_images[i] = augment(images[idx[i]])`` (plain gather when ``seed`` is None)."""
    assert images.dtype == torch.uint8 and images.is_contiguous() and out_images.dtype == torch.uint8 and out_images.is_contiguous()
    n_items, c, h, w = images.shape
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    n = len(idx)
    assert out_images.shape == (n, c, h, w)
    dx = dy = fl = None
    if seed is not None:
        dx, dy, fl = augment_draws(n, seed)
    rc = _lib().drc_host_gather_augment(images.data_ptr(), labels.data_ptr() if labels is not None else None, n_items, c, h, w,
                                        idx.ctypes.data, n, _np_ptr(dx), _np_ptr(dy), _np_ptr(fl), pad if seed is not None else -1,
                                        out_images.data_ptr(), out_labels.data_ptr() if out_labels is not None else None)
    if rc:
        raise IndexError(f"native gather failed (code {rc})")


class NativeLoader:
    """Thread-pool loader over an in-memory uint8 image dataset ([N, C, H, W]) and int64 labels."""

    def __init__(self, images: torch.Tensor, labels: torch.Tensor, threads: int = 2):
        assert images.dtype == torch.uint8 and images.is_contiguous() and images.dim() == 4
        assert labels.dtype == torch.int64 and labels.is_contiguous()
        self.images, self.labels = images, labels            # keep the storage alive
        n, c, h, w = images.shape
        self.shape = (c, h, w)
        self._lib = _lib()
        self._h = self._lib.drc_loader_create(images.data_ptr(), labels.data_ptr(), n, c, h, w, int(threads))
        self._last = -1

    def submit(self, idx: np.ndarray, out_images: torch.Tensor, out_labels: torch.Tensor, seed: Optional[int] = None,
               pad: int = 4) -> int:
        """Enqueue one sub-batch (``seed`` given => augmented).  The output tensors must outlive ``wait``."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        n = len(idx)
        assert out_images.dtype == torch.uint8 and out_images.is_contiguous() and tuple(out_images.shape) == (n,) + self.shape
        assert out_labels.dtype == torch.int64 and out_labels.is_contiguous() and out_labels.numel() == n
        dx = dy = fl = None
        if seed is not None:
            dx, dy, fl = augment_draws(n, seed)
        self._last = self._lib.drc_loader_submit(self._h, idx.ctypes.data, n, _np_ptr(dx), _np_ptr(dy), _np_ptr(fl),
                                                 pad if seed is not None else -1, out_images.data_ptr(), out_labels.data_ptr())
        return self._last

    def wait(self, ticket: Optional[int] = None) -> None:
        t = self._last if ticket is None else ticket
        if t < 0:
            return
        rc = self._lib.drc_loader_wait(self._h, t)
        if rc:
            raise IndexError(f"native loader job failed (code {rc}: 1 = index out of range, 2 = image too wide)")

    def close(self) -> None:
        if self._h:
            self._lib.drc_loader_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
