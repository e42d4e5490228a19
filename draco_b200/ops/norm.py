"""BatchNorm2d with fused residual-add + ReLU on the sm_100a kernels of ``csrc/cuda/bn_fused.cu``.

``FusedBatchNorm2d`` *is* an ``nn.BatchNorm2d`` (same parameters, buffers and state_dict keys as the reference's models,
src/model_ops/resnet.py:19-24) whose ``forward(x, residual=None, relu=False)`` computes
``relu?(batch_norm(x) + residual)``.  On CUDA, in training mode, for bf16 channels-last activations with a power-of-two
channel count it runs as streaming kernels with deterministic reductions (forward: apply only when the producing convolution's epilogue
delivered the statistics, else statistics + apply; backward: reduce + apply); everywhere else it is exactly ``F.batch_norm`` (+ add, + relu) so CPU runs and evaluation are unchanged.
``backend_counters`` records which path served each call.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .. import _native as N

backend_counters = {"fused": 0, "aten": 0}
_counter_cache = {}


def _lib():
    lib = N.cuda()
    if not getattr(lib, "_bn_ready", False):
        st = N.ptr
        lib.drc_bn_supported.argtypes = [C.c_int]
        lib.drc_bn_supported.restype = C.c_int
        lib.drc_bn_workspace.argtypes = [N.i64, C.c_int, C.c_int]
        lib.drc_bn_workspace.restype = N.i64
        lib.drc_bn_fwd.argtypes = [N.ptr] * 11 + [N.i64, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, st]
        lib.drc_bn_fwd.restype = C.c_int
        lib.drc_bn_bwd.argtypes = [N.ptr] * 14 + [N.i64, C.c_int, C.c_int, C.c_int, st]
        lib.drc_bn_bwd.restype = C.c_int
        lib.drc_bn_bwd_apply.argtypes = [N.ptr] * 7 + [N.i64, C.c_int, C.c_int, st]
        lib.drc_bn_bwd_apply.restype = C.c_int
        lib._bn_ready = True
    return lib


# Set by the engine when logical workers run on concurrent streams: only one worker per process may update the shared
# running statistics.
UPDATE_RUNNING_STATS = True


# ``num_batches_tracked += 1`` is a kernel launch per BatchNorm layer per forward (20 launches, 35 us of a 2.0 ms ResNet-18
# step, profiles/worker_profile_ResNet18_fused.txt).  Inside ``deferred_batch_counts()`` the fused path only remembers the
# counters and the context exit bumps all of them with ONE multi-tensor kernel.
_defer_counts = False
_pending_counts: list = []


class deferred_batch_counts:
    def __enter__(self):
        global _defer_counts
        self._prev = _defer_counts
        _defer_counts = True
        return self

    def __exit__(self, *exc):
        global _defer_counts
        _defer_counts = self._prev
        if not _defer_counts and _pending_counts:
            pend = list(_pending_counts)
            _pending_counts.clear()
            torch._foreach_add_(pend, 1)
        return False


def _counter(device: torch.device) -> torch.Tensor:
    """Grid-barrier / last-CTA counters; one set per (device, stream) so that concurrent streams never share them."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    if key not in _counter_cache:
        _counter_cache[key] = torch.zeros(32, dtype=torch.int32, device=device)   # zero-filled on this very stream
    return _counter_cache[key]


def _sms(device: torch.device) -> int:
    return torch.cuda.get_device_properties(device).multi_processor_count


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def fused_supported(x: torch.Tensor, training: bool) -> bool:
    return (training and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and
            x.is_contiguous(memory_format=torch.channels_last) and os.environ.get("DRACO_BN", "fused") == "fused" and
            bool(_lib().drc_bn_supported(x.shape[1])))


class _BnActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, gamma, beta, running_mean, running_var, eps, momentum, relu, mean, invstd, link_box=None):
        """``mean`` / ``invstd`` given: statistics came out of the producing convolution's epilogue, only the apply kernel runs."""
        lib = _lib()
        n, c, h, w = x.shape
        M = n * h * w
        dev = x.device
        y = torch.empty_like(x, memory_format=torch.channels_last)
        have_stats = mean is not None
        if residual is not None and not residual.is_contiguous(memory_format=torch.channels_last):
            residual = residual.contiguous(memory_format=torch.channels_last)
        if have_stats:
            ws_ptr = None
        else:
            mean = torch.empty(c, dtype=torch.float32, device=dev)
            invstd = torch.empty(c, dtype=torch.float32, device=dev)
            ws = torch.empty(int(lib.drc_bn_workspace(M, c, _sms(dev))), dtype=torch.float32, device=dev)
            ws_ptr = ws.data_ptr()
        N.check(lib.drc_bn_fwd(x.data_ptr(), _p(residual), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _p(running_mean),
                               _p(running_var), mean.data_ptr(), invstd.data_ptr(), ws_ptr, _counter(dev).data_ptr(),
                               M, c, float(eps), float(momentum), int(relu), _sms(dev), int(have_stats),
                               torch.cuda.current_stream().cuda_stream), "bn_fwd")
        # ReLU mask in backward: the saved output y.  DRACO_BN_MASK=x recomputes it from x with the forward's own arithmetic for
        # layers without a residual input (one saved activation less per layer: memory for large models); measured A/B in one
        # process it is ~1.5 % slower on the ResNet-18/CIFAR step (the activations are L2-resident there), so it is opt-in.
        mask_from_x = residual is None and os.environ.get("DRACO_BN_MASK", "y") == "x"
        ctx.save_for_backward(x, y if (relu and not mask_from_x) else None, gamma, beta, mean, invstd)
        ctx.relu, ctx.has_res = bool(relu), residual is not None
        ctx.link = None
        if link_box is not None:                       # see ops.conv.BnBwdLink: the consumer convolution may reduce for us
            from .conv import BnBwdLink
            ctx.link = link_box[0] = BnBwdLink(x, mean, invstd, relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib()
        x, y, gamma, beta, mean, invstd = ctx.saved_tensors
        n, c, h, w = x.shape
        M = n * h * w
        dev = x.device
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        link, ctx.link = ctx.link, None
        if link is not None and link.sums is not None and link.dz is not None and link.dz.data_ptr() == dy.data_ptr():
            # the consuming convolution's dgrad epilogue already masked dy (dz) and reduced sum(dz), sum(dz * xhat): apply only
            backend_counters["bwd_in_conv"] = backend_counters.get("bwd_in_conv", 0) + 1
            torch.cuda.set_device(dev)
            N.check(lib.drc_bn_bwd_apply(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                         link.sums.data_ptr(), dx.data_ptr(), M, c, _sms(dev),
                                         torch.cuda.current_stream().cuda_stream), "bn_bwd_apply")
            dgamma, dbeta = link.dgamma, link.dbeta
            link.dz = link.sums = None
            return dx, (dy if ctx.has_res else None), dgamma, dbeta, None, None, None, None, None, None, None, None
        dres = torch.empty_like(x, memory_format=torch.channels_last) if ctx.has_res else None
        dgamma = torch.empty(c, dtype=torch.float32, device=dev)
        dbeta = torch.empty(c, dtype=torch.float32, device=dev)
        torch.cuda.set_device(dev)
        sums = torch.empty(2 * c, dtype=torch.float32, device=dev)
        ws = torch.empty(int(lib.drc_bn_workspace(M, c, _sms(dev))), dtype=torch.float32, device=dev)
        N.check(lib.drc_bn_bwd(dy.data_ptr(), _p(y), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                               dx.data_ptr(), _p(dres), dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), sums.data_ptr(),
                               _counter(dev)[8:].data_ptr(), M, c, int(ctx.relu), _sms(dev),
                               torch.cuda.current_stream().cuda_stream), "bn_bwd")
        return dx, dres, dgamma, dbeta, None, None, None, None, None, None, None, None


class FusedBatchNorm2d(nn.BatchNorm2d):
    """``nn.BatchNorm2d`` + optional fused residual add and ReLU.  When the producing ``ops.conv.Conv2d`` was called with
    ``bn=self`` its epilogue already computed this layer's batch statistics (``pending_stats``) and only the apply kernel runs."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1, relu: bool = False):
        super().__init__(num_features, eps=eps, momentum=momentum)
        self.fuse_relu = relu
        self.pending_stats = None

    def _fused_ok(self, x: torch.Tensor) -> bool:
        return self.weight.dtype == torch.float32 and self.track_running_stats and fused_supported(x, self.training)

    def stat_request(self, x_in: torch.Tensor, channels: int):
        """Called by the producing convolution: a BnStatRequest if this layer will take the fused path for its output."""
        from .conv import BnStatRequest
        if not (self.training and self.weight.dtype == torch.float32 and self.track_running_stats and channels == self.num_features
                and channels <= 512 and os.environ.get("DRACO_BN", "fused") == "fused"
                and os.environ.get("DRACO_BN_STATS", "conv") == "conv" and bool(_lib().drc_bn_supported(channels))):
            return None
        upd = UPDATE_RUNNING_STATS
        return BnStatRequest(self.eps, self.momentum if self.momentum is not None else 0.1,
                             self.running_mean if upd else None, self.running_var if upd else None)

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None, relu: Optional[bool] = None) -> torch.Tensor:
        relu = self.fuse_relu if relu is None else relu
        pend, self.pending_stats = self.pending_stats, None
        if self._fused_ok(x) and (residual is None or residual.shape == x.shape):
            backend_counters["fused"] += 1
            upd = UPDATE_RUNNING_STATS
            if upd and self.num_batches_tracked is not None:
                if _defer_counts:
                    _pending_counts.append(self.num_batches_tracked)
                else:
                    self.num_batches_tracked.add_(1)
            mean = invstd = None
            if pend is not None and pend[0] is x:
                mean, invstd = pend[1], pend[2]                   # running statistics were updated by the convolution
                backend_counters["conv_stats"] = backend_counters.get("conv_stats", 0) + 1
            box = [None] if x.requires_grad else None
            y = _BnActFn.apply(x, residual, self.weight, self.bias, self.running_mean if upd else None,
                               self.running_var if upd else None, self.eps,
                               self.momentum if self.momentum is not None else 0.1, relu, mean, invstd, box)
            if box is not None and box[0] is not None:
                y._draco_bn_link = box[0]              # picked up by the ops.conv.Conv2d that consumes y (BatchNorm backward in its dgrad)
            return y
        assert pend is None, "a convolution produced fused statistics for a BatchNorm call that cannot use them"
        backend_counters["aten"] += 1
        if self.training and not UPDATE_RUNNING_STATS:
            y = F.batch_norm(x, None, None, self.weight, self.bias, True, 0.0, self.eps)
        else:
            y = super().forward(x)
        if residual is not None:
            y = y + residual
        return F.relu(y) if relu else y
