"""Replica-deterministic dropout (``ReplicaDropout``), safe under CUDA-graph replay.

The repetition and cyclic codes compare gradients of *replicas* of a batch bit for bit (reference: src/master/rep_master.py:154-168,
cyclic_master.py:167-213), so every holder of a batch must apply the same dropout mask.  ``nn.Dropout`` cannot guarantee that
inside a captured graph (the Philox offset advances per replay and per call site).  ``ReplicaDropout`` derives the mask from
``(seed, step, batch id, layer salt, element index)`` with a counter-based hash; on CUDA the step is read from a device tensor by
the kernel (csrc/cuda/dropout.cu), so a replayed graph advances by itself, and the backward pass recomputes the mask instead of
storing it.  The worker runtime sets the key with ``set_context`` before each sub-batch (parallel/worker.py); without a context
the layer behaves like ``nn.Dropout`` (single-process training, evaluation).

Reference counterpart: ``nn.Dropout()`` in the VGG classifier, src/model_ops/vgg.py:24-31.
"""
from __future__ import annotations

import ctypes as C
import torch
from torch import nn

_ctx = {"step": None, "seed": 0, "batch": 0}      # step: device int64 tensor [1] (CUDA) or python int (CPU)
_MASK64 = (1 << 64) - 1


def set_context(step, seed: int, batch_id: int) -> None:
    """``step``: device int64 tensor read by the kernel at execution time, or an int on CPU.  ``None`` disables the context."""
    _ctx["step"], _ctx["seed"], _ctx["batch"] = step, int(seed), int(batch_id)


def clear_context() -> None:
    _ctx["step"] = None


def _mix64(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & _MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
    return z ^ (z >> 31)


def _key(salt: int) -> int:
    return _mix64((_ctx["seed"] * 0x100000001B3 + _ctx["batch"] * 0x10001 + salt * 0x9E3779B1) & _MASK64)


def _lib():
    from .. import _native as N
    lib = N.cuda()
    if not getattr(lib, "_dropout_ready", False):
        lib.drc_dropout.argtypes = [N.ptr, N.ptr, N.ptr, N.i64, C.c_uint64, C.c_float, C.c_int, N.ptr]
        lib.drc_dropout.restype = C.c_int
        lib._dropout_ready = True
    return lib


def _launch(x: torch.Tensor, step: torch.Tensor, key: int, p: float) -> torch.Tensor:
    from .. import _native as N
    y = torch.empty_like(x)
    N.check(_lib().drc_dropout(x.data_ptr(), y.data_ptr(), step.data_ptr(), x.numel(), key, float(p), int(x.dtype == torch.bfloat16),
                               torch.cuda.current_stream().cuda_stream), "dropout")
    return y


class _KeyedDropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, step, key, p):
        ctx.step, ctx.key, ctx.p = step, key, p
        return _launch(x.contiguous(), step, key, p)

    @staticmethod
    def backward(ctx, gy):
        return _launch(gy.contiguous(), ctx.step, ctx.key, ctx.p), None, None, None      # same (step, key) -> same mask


def _cpu_mask(shape, step: int, key: int, p: float) -> torch.Tensor:
    g = torch.Generator()
    g.manual_seed(_mix64(key ^ _mix64(step & _MASK64)) & 0x7FFFFFFFFFFFFFFF)
    return (torch.rand(shape, generator=g) >= p).float() / (1.0 - p)


class ReplicaDropout(nn.Dropout):
    """``nn.Dropout`` whose mask is a function of the context key when one is set (see module docstring)."""

    def __init__(self, p: float = 0.5, inplace: bool = False, salt: int = 0):
        super().__init__(p, False)
        self.salt = int(salt)             # position of the layer in its model (set by the model: same in every process)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        step = _ctx["step"]
        if not self.training or self.p == 0.0 or step is None:
            return super().forward(x)
        key = _key(self.salt)
        if x.is_cuda and x.dtype in (torch.bfloat16, torch.float32) and isinstance(step, torch.Tensor):
            return _KeyedDropoutFn.apply(x, step, key, self.p)
        s = int(step.item()) if isinstance(step, torch.Tensor) else int(step)
        return x * _cpu_mask(x.shape, s, key, self.p).to(x.dtype).to(x.device)
