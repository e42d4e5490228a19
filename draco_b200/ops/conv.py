"""Convolution layer whose pointwise (1x1, stride 1) case runs on the hand-written tcgen05 GEMM.

In NHWC a 1x1 convolution *is* a GEMM: ``y[N*H*W, Cout] = x[N*H*W, Cin] @ W[Cout, Cin]^T`` -- both operands K-major, no
im2col, no layout change (weights live in the arena as ``[Cout, kH, kW, Cin]``).  The three products of the layer map onto
``csrc/cuda/gemm_tcgen05.cu`` exactly like a Linear layer (ops/linear.py): fprop K-major x K-major, dgrad with an
MN-major B, wgrad with MN-major A and B (K = N*H*W).  This covers two thirds of the convolutions of the bottleneck
ResNets (50/101/152: conv1, conv3 and the stride-1 projection shortcuts).  Everything else -- 3x3, strided -- goes to
cuDNN's implicit-GEMM kernels in deterministic mode unless ``DRACO_CONV3X3=tcgen05`` selects the TMA-patch implicit-GEMM
kernels of ``csrc/cuda/conv_tcgen05.cu`` (fprop, dgrad and split-K wgrad) for the 3x3 / stride-1 layers.  ``backend_counters`` records which path served each call.

Reference counterpart: ``nn.Conv2d`` inside src/model_ops/resnet.py:14-64 / vgg.py:46-59 (PyTorch-0.3 CPU THNN).
"""
from __future__ import annotations

import os

import torch
from torch import nn

backend_counters = {"tcgen05": 0, "cudnn": 0}


class _Conv1x1Fn(torch.autograd.Function):
    """x2: [M, Cin] (NHWC rows), w2: [Cout, Cin] -> y2: [M, Cout]."""

    @staticmethod
    def forward(ctx, x2, w2, bias):
        from . import kernels as K
        ctx.save_for_backward(x2, w2)
        ctx.has_bias = bias is not None
        return K.gemm_bf16(x2, w2, bias=bias)

    @staticmethod
    def backward(ctx, dy2):
        from . import kernels as K
        x2, w2 = ctx.saved_tensors
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = K.gemm_bf16(dy2, w2, b_mn=True) if ctx.needs_input_grad[0] else None
        dw = K.gemm_bf16(dy2, x2, a_mn=True, b_mn=True) if ctx.needs_input_grad[1] else None
        db = dy2.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db


def _lib():
    import ctypes as C
    from .. import _native as N
    lib = N.cuda()
    if not getattr(lib, "_conv_ready", False):
        lib.drc_conv3x3_supported.argtypes = [C.c_int] * 5
        lib.drc_conv3x3_supported.restype = C.c_int
        lib.drc_conv3x3.argtypes = [N.ptr, N.ptr, N.ptr] + [C.c_int] * 6 + [N.ptr, N.ptr, C.c_int, C.c_int, N.ptr]
        lib.drc_conv3x3.restype = C.c_int
        lib.drc_conv3x3_wgrad_supported.argtypes = [C.c_int] * 4
        lib.drc_conv3x3_wgrad_supported.restype = C.c_int
        lib.drc_conv3x3_wgrad_plan.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_longlong)]
        lib.drc_conv3x3_wgrad_plan.restype = C.c_int
        lib.drc_conv3x3_wgrad.argtypes = [N.ptr] * 4 + [C.c_int] * 7 + [N.ptr]
        lib.drc_conv3x3_wgrad.restype = C.c_int
        lib.drc_convg_supported.argtypes = [C.c_int] * 6
        lib.drc_convg_supported.restype = C.c_int
        lib.drc_convg.argtypes = [N.ptr, N.ptr, N.ptr] + [C.c_int] * 8 + [N.ptr, N.ptr, C.c_int, C.c_int, N.ptr]
        lib.drc_convg.restype = C.c_int
        lib.drc_convg_wgrad_supported.argtypes = [C.c_int] * 6
        lib.drc_convg_wgrad_supported.restype = C.c_int
        lib.drc_convg_wgrad_plan.argtypes = [C.c_int] * 8 + [C.POINTER(C.c_longlong)]
        lib.drc_convg_wgrad_plan.restype = C.c_int
        lib.drc_convg_wgrad.argtypes = [N.ptr] * 4 + [C.c_int] * 9 + [N.ptr]
        lib.drc_convg_wgrad.restype = C.c_int
        lib.drc_conv_halo_supported.argtypes = [C.c_int] * 4
        lib.drc_conv_halo_supported.restype = C.c_int
        lib.drc_conv_halo.argtypes = [N.ptr, N.ptr, N.ptr] + [C.c_int] * 4 + [N.ptr, N.ptr, C.c_int, C.c_int, C.c_int, C.c_int, N.ptr]
        lib.drc_conv_halo.restype = C.c_int
        lib.drc_conv_stem_supported.argtypes = [C.c_int] * 4
        lib.drc_conv_stem_supported.restype = C.c_int
        lib.drc_conv_stem_wgrad_parts.argtypes = [C.c_int] * 4
        lib.drc_conv_stem_wgrad_parts.restype = C.c_int
        lib.drc_conv_stem_fprop.argtypes = [N.ptr] * 4 + [C.c_int] * 4 + [N.ptr]
        lib.drc_conv_stem_fprop.restype = C.c_int
        lib.drc_conv_stem_wgrad.argtypes = [N.ptr] * 4 + [C.c_int] * 5 + [N.ptr]
        lib.drc_conv_stem_wgrad.restype = C.c_int
        lib._conv_ready = True
    return lib


def conv3x3_wgrad_tcgen05(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """Weight gradient of the 3x3 / stride 1 / pad 1 convolution on the split-K tcgen05 kernel: both operands are read
    MN-major straight from the NHWC activations (K = pixels), fp32 partials are folded in a fixed order.  Returns
    [Cout, Cin, 3, 3] in channels-last storage (the arena layout)."""
    import ctypes as C
    from .. import _native as N
    from . import kernels as K
    lib = _lib()
    n, cout, h, w = dy.shape
    cin = x.shape[1]
    assert dy.is_contiguous(memory_format=torch.channels_last) and x.is_contiguous(memory_format=torch.channels_last)
    sms = K.sm_count(dy.device)
    ws_elems = C.c_longlong(0)
    lib.drc_conv3x3_wgrad_plan(n, h, w, cin, cout, sms, C.byref(ws_elems))
    ws = torch.empty(ws_elems.value, dtype=torch.float32, device=dy.device)
    dw = torch.empty((cout, cin, 3, 3), dtype=torch.bfloat16, device=dy.device, memory_format=torch.channels_last)
    N.check(lib.drc_conv3x3_wgrad(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), ws.data_ptr(), n, h, w, cin, cout, sms,
                                  dy.device.index, torch.cuda.current_stream().cuda_stream), "conv3x3_wgrad")
    return dw


def conv3x3_tcgen05(act: torch.Tensor, weight: torch.Tensor, dgrad: bool = False, bias: torch.Tensor = None) -> torch.Tensor:
    """3x3 / stride 1 / pad 1 convolution (``dgrad=False``) or its backward-data pass (``dgrad=True``) on the tcgen05
    implicit-GEMM kernel (csrc/cuda/conv_tcgen05.cu).  ``act``: channels-last bf16 [N, C, H, W]; ``weight``:
    [Cout, Cin, 3, 3] in channels-last storage ([Cout, 3, 3, Cin] in memory, the arena layout)."""
    from .. import _native as N
    from . import kernels as K
    lib = _lib()
    n, _, h, w = act.shape
    cout, cin = weight.shape[0], weight.shape[1]
    assert act.is_contiguous(memory_format=torch.channels_last) and weight.is_contiguous(memory_format=torch.channels_last)
    out = torch.empty((n, cin if dgrad else cout, h, w), dtype=torch.bfloat16, device=act.device,
                      memory_format=torch.channels_last)
    bf32 = bias.data_ptr() if bias is not None and bias.dtype == torch.float32 else None
    bb16 = bias.data_ptr() if bias is not None and bias.dtype == torch.bfloat16 else None
    N.check(lib.drc_conv3x3(act.data_ptr(), weight.data_ptr(), out.data_ptr(), n, h, w, cin, cout, int(dgrad), bf32, bb16,
                            K.sm_count(act.device), act.device.index, torch.cuda.current_stream().cuda_stream), "conv3x3")
    return out


# ---------------------------------------------------------------------------------------------------------------------
# generalised tap-table kernels (csrc/cuda/conv_strided_tcgen05.cu): stride 1 or 2, 1x1 or 3x3.  Numerics validated on
# hardware, not yet timed against cuDNN: opt-in with DRACO_CONV_STRIDED=tcgen05.
# ---------------------------------------------------------------------------------------------------------------------
def convg_tcgen05(act: torch.Tensor, weight: torch.Tensor, in_hw, stride: int, dgrad: bool = False,
                  bias: torch.Tensor = None) -> torch.Tensor:
    """``dgrad=False``: ``act`` = x [N, Cin, H, W] -> y [N, Cout, H/stride, W/stride];  ``dgrad=True``: ``act`` = dy -> dx.
    ``in_hw`` is the spatial size of the forward input x; ``weight`` [Cout, Cin, ks, ks] in channels-last storage."""
    from .. import _native as N
    from . import kernels as K
    lib = _lib()
    n = act.shape[0]
    h, w = in_hw
    cout, cin, ks = weight.shape[0], weight.shape[1], weight.shape[2]
    assert act.is_contiguous(memory_format=torch.channels_last)
    assert weight.permute(0, 2, 3, 1).is_contiguous(), "weights must be stored [Cout, ks, ks, Cin]"
    oshape = (n, cin, h, w) if dgrad else (n, cout, h // stride, w // stride)
    out = torch.empty(oshape, dtype=torch.bfloat16, device=act.device, memory_format=torch.channels_last)
    bf32 = bias.data_ptr() if bias is not None and bias.dtype == torch.float32 else None
    bb16 = bias.data_ptr() if bias is not None and bias.dtype == torch.bfloat16 else None
    N.check(lib.drc_convg(act.data_ptr(), weight.data_ptr(), out.data_ptr(), n, h, w, cin, cout, ks, stride, int(dgrad), bf32, bb16,
                          K.sm_count(act.device), act.device.index, torch.cuda.current_stream().cuda_stream), "convg")
    return out


def convg_wgrad_tcgen05(dy: torch.Tensor, x: torch.Tensor, ks: int, stride: int) -> torch.Tensor:
    """Weight gradient [Cout, Cin, ks, ks] (channels-last storage) of the strided / 1x1 / 3x3 convolution."""
    import ctypes as C
    from .. import _native as N
    from . import kernels as K
    lib = _lib()
    n, cout = dy.shape[0], dy.shape[1]
    cin, h, w = x.shape[1], x.shape[2], x.shape[3]
    assert dy.is_contiguous(memory_format=torch.channels_last) and x.is_contiguous(memory_format=torch.channels_last)
    sms = K.sm_count(dy.device)
    ws_elems = C.c_longlong(0)
    lib.drc_convg_wgrad_plan(n, h, w, cin, cout, ks, stride, sms, C.byref(ws_elems))
    ws = torch.empty(ws_elems.value, dtype=torch.float32, device=dy.device)
    dw = torch.empty((cout, ks, ks, cin), dtype=torch.bfloat16, device=dy.device).permute(0, 3, 1, 2)   # [Cout,Cin,ks,ks] view
    N.check(lib.drc_convg_wgrad(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), ws.data_ptr(), n, h, w, cin, cout, ks, stride, sms,
                                dy.device.index, torch.cuda.current_stream().cuda_stream), "convg_wgrad")
    return dw


# ---------------------------------------------------------------------------------------------------------------------
# 3-channel stem (csrc/cuda/conv_stem.cu), CUDA cores.  Numerics validated on hardware, not yet timed: opt-in with DRACO_CONV_STEM=native.
# ---------------------------------------------------------------------------------------------------------------------
def conv_stem_fprop(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor = None) -> torch.Tensor:
    """x: channels-last bf16 [N, 3, H, W]; weight [64, 3, 3, 3] stored [Cout, 3, 3, Cin] -> y channels-last [N, 64, H, W]."""
    from .. import _native as N
    n, _, h, w = x.shape
    assert x.is_contiguous(memory_format=torch.channels_last) and weight.permute(0, 2, 3, 1).is_contiguous()
    y = torch.empty((n, 64, h, w), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    b = bias.float().contiguous() if bias is not None else None
    N.check(_lib().drc_conv_stem_fprop(x.data_ptr(), weight.data_ptr(), y.data_ptr(), b.data_ptr() if b is not None else None, n, h, w,
                                       x.device.index, torch.cuda.current_stream().cuda_stream), "conv_stem_fprop")
    return y


def conv_stem_wgrad(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    from .. import _native as N
    from . import kernels as K
    lib = _lib()
    n, _, h, w = x.shape
    assert dy.is_contiguous(memory_format=torch.channels_last) and x.is_contiguous(memory_format=torch.channels_last)
    sms = K.sm_count(dy.device)
    parts = lib.drc_conv_stem_wgrad_parts(n, h, w, sms)
    ws = torch.empty(parts * 64 * 27, dtype=torch.float32, device=dy.device)
    dw = torch.empty((64, 3, 3, 3), dtype=torch.bfloat16, device=dy.device).permute(0, 3, 1, 2)       # [Cout,Cin,3,3] view
    N.check(lib.drc_conv_stem_wgrad(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), ws.data_ptr(), n, h, w, sms, dy.device.index,
                                    torch.cuda.current_stream().cuda_stream), "conv_stem_wgrad")
    return dw


class _ConvStemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x)
        ctx.has_bias = bias is not None
        return conv_stem_fprop(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        assert not ctx.needs_input_grad[0], "the stem kernel has no dgrad (network inputs carry no gradient)"
        dw = conv_stem_wgrad(dy, x) if ctx.needs_input_grad[1] else None
        db = dy.sum((0, 2, 3)) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return None, dw, db


class _ConvGFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride):
        ctx.save_for_backward(x, weight)
        ctx.stride, ctx.has_bias = stride, bias is not None
        return convg_tcgen05(x, weight, x.shape[2:], stride, False, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = dw = db = None
        backend_counters["tcgen05"] += 1
        if ctx.needs_input_grad[0]:
            dx = convg_tcgen05(dy, weight, x.shape[2:], ctx.stride, True)
        if ctx.needs_input_grad[1]:
            dw = convg_wgrad_tcgen05(dy, x, weight.shape[2], ctx.stride)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum((0, 2, 3))
        return dx, dw, db, None


def _halo_ok(h: int, w: int, cin: int, cout: int) -> bool:
    """Halo-reuse kernels (csrc/cuda/conv_halo_tcgen05.cu) for the 64 -> 64 layers: opt-in with DRACO_CONV3X3=halo (never run
    on hardware yet; DRACO_HALO_DESC=0|1 and DRACO_HALO_PW=10|16 select the descriptor convention / patch pitch to try)."""
    return os.environ.get("DRACO_CONV3X3", "cudnn") == "halo" and bool(_lib().drc_conv_halo_supported(h, w, cin, cout))


def conv3x3_halo(act: torch.Tensor, weight: torch.Tensor, dgrad: bool = False, bias: torch.Tensor = None) -> torch.Tensor:
    from .. import _native as N
    from . import kernels as K
    n, _, h, w = act.shape
    assert act.is_contiguous(memory_format=torch.channels_last) and weight.is_contiguous(memory_format=torch.channels_last)
    out = torch.empty((n, 64, h, w), dtype=torch.bfloat16, device=act.device, memory_format=torch.channels_last)
    bf32 = bias.data_ptr() if bias is not None and bias.dtype == torch.float32 else None
    bb16 = bias.data_ptr() if bias is not None and bias.dtype == torch.bfloat16 else None
    N.check(_lib().drc_conv_halo(act.data_ptr(), weight.data_ptr(), out.data_ptr(), n, h, w, int(dgrad), bf32, bb16,
                                 int(os.environ.get("DRACO_HALO_DESC", "0")), int(os.environ.get("DRACO_HALO_PW", "10")),
                                 K.sm_count(act.device), act.device.index,
                                 torch.cuda.current_stream().cuda_stream), "conv_halo")
    return out


class _Conv3x3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        if _halo_ok(x.shape[2], x.shape[3], weight.shape[1], weight.shape[0]):
            return conv3x3_halo(x, weight, False, bias)
        return conv3x3_tcgen05(x, weight, False, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        cout, cin = weight.shape[0], weight.shape[1]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if _halo_ok(x.shape[2], x.shape[3], cin, cout):
                backend_counters["tcgen05"] += 1
                dx = conv3x3_halo(dy, weight, True)
            elif _lib().drc_conv3x3_supported(x.shape[2], x.shape[3], cin, cout, 1):
                backend_counters["tcgen05"] += 1
                dx = conv3x3_tcgen05(dy, weight, True)
            else:
                backend_counters["cudnn"] += 1
                dx = torch.ops.aten.convolution_backward(dy, x, weight, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                                         [True, False, False])[0]
        if ctx.needs_input_grad[1]:
            if (os.environ.get("DRACO_CONV3X3_WGRAD", "tcgen05") == "tcgen05"
                    and _lib().drc_conv3x3_wgrad_supported(x.shape[2], x.shape[3], cin, cout)):
                backend_counters["tcgen05"] += 1
                dw = conv3x3_wgrad_tcgen05(dy, x)
            else:
                backend_counters["cudnn"] += 1
                dw = torch.ops.aten.convolution_backward(dy, x, weight, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                                         [False, True, False])[1]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum((0, 2, 3))
        return dx, dw, db


class Conv2d(nn.Conv2d):
    """``nn.Conv2d`` (same parameters / state_dict) with the pointwise fast path described above."""

    def _pointwise_ok(self, x: torch.Tensor) -> bool:
        return (self.kernel_size == (1, 1) and self.stride == (1, 1) and self.padding == (0, 0) and self.dilation == (1, 1)
                and self.groups == 1 and x.is_cuda and x.dtype == torch.bfloat16 and self.weight.dtype == torch.bfloat16
                and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
                and self.in_channels % 8 == 0 and self.out_channels % 8 == 0 and self.in_channels >= 64 and self.out_channels >= 64
                and os.environ.get("DRACO_CONV1X1", "tcgen05") == "tcgen05")

    def _conv3x3_ok(self, x: torch.Tensor) -> bool:
        return (self.kernel_size == (3, 3) and self.stride == (1, 1) and self.padding == (1, 1) and self.dilation == (1, 1)
                and self.groups == 1 and x.is_cuda and x.dtype == torch.bfloat16 and self.weight.dtype == torch.bfloat16
                and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
                and self.weight.is_contiguous(memory_format=torch.channels_last)
                and os.environ.get("DRACO_CONV3X3", "cudnn") in ("tcgen05", "halo")
                and bool(_lib().drc_conv3x3_supported(x.shape[2], x.shape[3], self.in_channels, self.out_channels, 0)))

    def _strided_ok(self, x: torch.Tensor) -> bool:
        """Opt-in path (DRACO_CONV_STRIDED=tcgen05): stride-2 3x3 and 1x1 layers on the tap-table kernels."""
        if os.environ.get("DRACO_CONV_STRIDED", "cudnn") != "tcgen05":
            return False
        ks = self.kernel_size[0]
        return (self.kernel_size in ((1, 1), (3, 3)) and self.stride == (2, 2) and self.padding == (ks // 2, ks // 2)
                and self.dilation == (1, 1) and self.groups == 1 and x.is_cuda and x.dtype == torch.bfloat16
                and self.weight.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
                and self.weight.permute(0, 2, 3, 1).is_contiguous()
                and bool(_lib().drc_convg_supported(x.shape[2], x.shape[3], self.in_channels, self.out_channels, ks, 2))
                and bool(_lib().drc_convg_wgrad_supported(x.shape[2], x.shape[3], self.in_channels, self.out_channels, ks, 2)))

    def _stem_ok(self, x: torch.Tensor) -> bool:
        return (os.environ.get("DRACO_CONV_STEM", "cudnn") == "native" and self.kernel_size == (3, 3) and self.stride == (1, 1)
                and self.padding == (1, 1) and self.dilation == (1, 1) and self.groups == 1 and self.in_channels == 3
                and self.out_channels == 64 and x.is_cuda and x.dtype == torch.bfloat16 and self.weight.dtype == torch.bfloat16
                and x.dim() == 4 and not x.requires_grad and x.is_contiguous(memory_format=torch.channels_last)
                and self.weight.permute(0, 2, 3, 1).is_contiguous()
                and bool(_lib().drc_conv_stem_supported(x.shape[2], x.shape[3], 3, 64)))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._stem_ok(x):
            backend_counters["native_stem"] = backend_counters.get("native_stem", 0) + 1
            return _ConvStemFn.apply(x, self.weight, self.bias)
        if self._strided_ok(x):
            backend_counters["tcgen05"] += 1
            return _ConvGFn.apply(x, self.weight, self.bias, 2)
        if self._conv3x3_ok(x):
            backend_counters["tcgen05"] += 1
            return _Conv3x3Fn.apply(x, self.weight, self.bias)
        if not self._pointwise_ok(x):
            backend_counters["cudnn"] += 1
            return super().forward(x)
        backend_counters["tcgen05"] += 1
        n, c, h, w = x.shape
        x2 = x.permute(0, 2, 3, 1).reshape(n * h * w, c)                    # view: NHWC rows
        w2 = self.weight.permute(0, 2, 3, 1).reshape(self.out_channels, c)  # view of the [Cout,1,1,Cin] arena storage
        if not w2.is_contiguous():
            w2 = w2.contiguous()
        y2 = _Conv1x1Fn.apply(x2, w2, self.bias)
        return y2.view(n, h, w, self.out_channels).permute(0, 3, 1, 2)      # channels-last [N, Cout, H, W]
