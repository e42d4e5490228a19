"""Global average pooling on channels-last activations (``csrc/cuda/pool_head.cu``).

The ResNet heads end in ``F.avg_pool2d(out, 4)`` over a 4x4 map (reference: src/model_ops/resnet.py:102-104), i.e. a global
average; on CUDA / bf16 / channels-last this module runs it -- and its backward, a 1/HW broadcast -- as one 16-byte-vectorised
kernel each.  Other inputs take ``F.adaptive_avg_pool2d``.  ``backend_counters`` records which path served a call.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn.functional as F

backend_counters = {"native": 0, "aten": 0}


def _lib():
    from .. import _native as N
    lib = N.cuda()
    if not getattr(lib, "_pool_ready", False):
        lib.drc_gap_fwd.argtypes = [N.ptr, N.ptr, C.c_int, C.c_int, C.c_int, N.ptr]
        lib.drc_gap_fwd.restype = C.c_int
        lib.drc_gap_bwd.argtypes = [N.ptr, N.ptr, C.c_int, C.c_int, C.c_int, N.ptr]
        lib.drc_gap_bwd.restype = C.c_int
        lib.drc_head_prep.argtypes = [N.ptr, N.ptr, N.ptr, C.c_int, C.c_int, C.c_int, C.c_int, N.ptr]
        lib.drc_head_prep.restype = C.c_int
        lib._pool_ready = True
    return lib


class _GapFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        from .. import _native as N
        n, c, h, w = x.shape
        ctx.shape = (n, c, h, w)
        y = torch.empty((n, c), dtype=x.dtype, device=x.device)
        N.check(_lib().drc_gap_fwd(x.data_ptr(), y.data_ptr(), n, h * w, c, torch.cuda.current_stream().cuda_stream), "gap_fwd")
        return y

    @staticmethod
    def backward(ctx, dy):
        from .. import _native as N
        n, c, h, w = ctx.shape
        dy = dy.contiguous()
        dx = torch.empty((n, c, h, w), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        N.check(_lib().drc_gap_bwd(dy.data_ptr(), dx.data_ptr(), n, h * w, c, torch.cuda.current_stream().cuda_stream), "gap_bwd")
        return dx


def global_avg_pool(x: torch.Tensor) -> torch.Tensor:
    """[N, C, H, W] -> [N, C] mean over H x W."""
    if (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] % 8 == 0
            and x.is_contiguous(memory_format=torch.channels_last) and os.environ.get("DRACO_POOL", "native") == "native"):
        backend_counters["native"] += 1
        return _GapFn.apply(x)
    backend_counters["aten"] += 1
    return F.adaptive_avg_pool2d(x, 1).flatten(1)


def head_prep(dy: torch.Tensor, npad: int, want_bias_grad: bool):
    """Backward helper of a narrow Linear: (dy zero-padded to ``npad`` columns, column sums of dy or None) in ONE launch."""
    from .. import _native as N
    b, n = dy.shape
    dyp = torch.empty((b, npad), dtype=dy.dtype, device=dy.device)
    db = torch.empty((n,), dtype=dy.dtype, device=dy.device) if want_bias_grad else None
    N.check(_lib().drc_head_prep(dy.data_ptr(), dyp.data_ptr(), db.data_ptr() if db is not None else None, 1, b, n, npad,
                                 torch.cuda.current_stream().cuda_stream), "head_prep")
    return dyp, db
