"""Flat parameter / gradient arenas.

The reference ships one message per parameter tensor per worker per step (62 up + 62 down for ResNet-18;
src/master/baseline_master.py:180-200).  Here every rank keeps parameters, gradients and momentum in flat arenas that
share one element layout, so a step is one push and one broadcast; per-tensor semantics (the reference votes /
selects / decodes *per tensor*) are preserved through a tile table: each tensor starts on a ``TILE``-element boundary
and a kernel CTA looks up the tensor its tile belongs to.

Dtype policy (bf16 compute): parameters of BatchNorm layers stay fp32; everything else is computed in bf16 from a bf16
copy that a cast kernel refreshes from the fp32 arena the PS broadcasts.  The *wire* and the PS master copy are fp32.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import nn

from .. import _native as N

TILE = N.TILE


@dataclass(frozen=True)
class TensorSpec:
    name: str
    shape: Tuple[int, ...]
    numel: int
    offset: int            # element offset in every arena, multiple of TILE
    is_bf16: bool          # computed in bf16 on workers (when the job runs in bf16)


class ArenaLayout:
    """Element layout shared by all arenas of a model."""

    def __init__(self, specs: List[TensorSpec], channels_last: bool = False):
        self.specs = specs
        # 4-d (conv) tensors are stored [Cout, kH, kW, Cin] in the arena and exposed as channels_last views, so
        # cuDNN / our NHWC kernels never re-layout weights or weight gradients
        self.channels_last = channels_last
        self.total = 0 if not specs else specs[-1].offset + _round_up(specs[-1].numel, TILE)
        self.ntensors = len(specs)
        self.ntiles = self.total // TILE
        tile_tensor = np.empty(self.ntiles, dtype=np.int32)
        for i, s in enumerate(specs):
            t0 = s.offset // TILE
            tile_tensor[t0: t0 + _round_up(s.numel, TILE) // TILE] = i
        self.tile_tensor_np = tile_tensor
        self.meta_np = np.zeros(self.ntensors, dtype=[("offset", "<i8"), ("numel", "<i8"), ("is_bf16", "<i4"), ("pad", "<i4")])
        for i, s in enumerate(specs):
            self.meta_np[i] = (s.offset, s.numel, int(s.is_bf16), 0)
        self._dev: Dict[torch.device, Tuple[torch.Tensor, torch.Tensor]] = {}

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_model(cls, model: nn.Module, bf16: bool = True, channels_last: bool = False) -> "ArenaLayout":
        bn_params = set()
        for m in model.modules():
            if isinstance(m, nn.modules.batchnorm._BatchNorm):
                bn_params.update(id(p) for p in m.parameters(recurse=False))
        specs, off = [], 0
        for name, p in model.named_parameters():
            if not p.requires_grad:
                continue
            specs.append(TensorSpec(name, tuple(p.shape), p.numel(), off, bf16 and id(p) not in bn_params))
            off += _round_up(p.numel(), TILE)
        return cls(specs, channels_last)

    @property
    def num_params(self) -> int:
        return sum(s.numel for s in self.specs)

    def valid_mask(self) -> np.ndarray:
        """bool[total]: True on real elements, False on padding."""
        m = np.zeros(self.total, dtype=bool)
        for s in self.specs:
            m[s.offset: s.offset + s.numel] = True
        return m

    # ------------------------------------------------------------------ views
    def view(self, arena: torch.Tensor, i: int) -> torch.Tensor:
        s = self.specs[i]
        flat = arena[s.offset: s.offset + s.numel]
        if self.channels_last and len(s.shape) == 4:
            return flat.view(s.shape[0], s.shape[2], s.shape[3], s.shape[1]).permute(0, 3, 1, 2)
        return flat.view(s.shape)

    def views(self, arena: torch.Tensor) -> List[torch.Tensor]:
        return [self.view(arena, i) for i in range(self.ntensors)]

    def flatten_into(self, arena: torch.Tensor, tensors: List[torch.Tensor]) -> None:
        for i, t in enumerate(tensors):
            self.view(arena, i).copy_(t)

    def unflatten(self, arena: torch.Tensor) -> List[torch.Tensor]:
        return [self.view(arena, i).clone() for i in range(self.ntensors)]

    def new_arena(self, device, dtype=torch.float32, rows: int = 0) -> torch.Tensor:
        shape = (self.total,) if rows == 0 else (rows, self.total)
        return torch.zeros(shape, dtype=dtype, device=device)

    # ------------------------------------------------------------------ device tables for the kernels
    def device_tables(self, device: torch.device) -> Tuple[torch.Tensor, torch.Tensor]:
        device = torch.device(device)
        if device not in self._dev:
            tt = torch.from_numpy(self.tile_tensor_np).to(device)
            meta = torch.from_numpy(self.meta_np.view(np.uint8).reshape(-1).copy()).to(device)
            self._dev[device] = (tt, meta)
        return self._dev[device]

    def tile_view(self, device: torch.device) -> N.TileView:
        tt, meta = self.device_tables(device)
        return N.TileView(tt.data_ptr(), meta.data_ptr(), self.ntiles, self.ntensors)


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class ModelBinder:
    """Re-points a model's parameters and gradients at arena views (no per-step gather/scatter copies).

    * ``params_f32``  : the arena the PS broadcast lands in (fp32, all tensors)
    * ``params_c``    : bf16 compute copy (only tensors flagged ``is_bf16`` are meaningful)
    * ``grads_f32`` / ``grads_c`` : gradient arenas; ``p.grad`` is a view so autograd accumulates in place
    """

    def __init__(self, model: nn.Module, layout: ArenaLayout, device, bf16: bool,
                 params_f32: Optional[torch.Tensor] = None):
        self.model, self.layout, self.bf16 = model, layout, bf16
        self.device = torch.device(device)
        self.params = [p for p in model.parameters() if p.requires_grad]
        assert len(self.params) == layout.ntensors
        # `params_f32` may be a view into a peer-mapped region so that the PS broadcast lands in the model itself
        self.params_f32 = params_f32 if params_f32 is not None else layout.new_arena(self.device)
        assert self.params_f32.numel() == layout.total and self.params_f32.dtype == torch.float32
        self.params_c = layout.new_arena(self.device, torch.bfloat16) if bf16 else None
        # initial values -> fp32 arena
        with torch.no_grad():
            layout.flatten_into(self.params_f32, [p.detach().to(self.device, torch.float32) for p in self.params])
        model.to(self.device)
        if bf16:
            # floating buffers of non-BN modules follow the compute dtype; BN statistics stay fp32
            for m in model.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    continue
                for k, b in list(m._buffers.items()):
                    if b is not None and b.is_floating_point():
                        m._buffers[k] = b.to(torch.bfloat16)
        for i, p in enumerate(self.params):
            spec = layout.specs[i]
            src = self.params_c if (bf16 and spec.is_bf16) else self.params_f32
            p.data = layout.view(src, i)
        if bf16:
            self.refresh_compute_copy()

    def refresh_compute_copy(self) -> None:
        """Torch fallback of the cast kernel (CPU path / tests)."""
        if self.params_c is None:
            return
        with torch.no_grad():
            for i, s in enumerate(self.layout.specs):
                if s.is_bf16:
                    self.layout.view(self.params_c, i).copy_(self.layout.view(self.params_f32, i))

    def new_grad_arenas(self) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        g32 = self.layout.new_arena(self.device)
        g16 = self.layout.new_arena(self.device, torch.bfloat16) if self.bf16 else None
        return g32, g16

    def bind_grads(self, g32: torch.Tensor, g16: Optional[torch.Tensor]) -> None:
        """Make ``p.grad`` views of the given arenas (autograd then accumulates in place)."""
        for i, p in enumerate(self.params):
            spec = self.layout.specs[i]
            src = g16 if (self.bf16 and spec.is_bf16) else g32
            p.grad = self.layout.view(src, i)

    def flat_grad_f32(self, g32: torch.Tensor, g16: Optional[torch.Tensor]) -> torch.Tensor:
        """Torch fallback of the push kernel's gather: one fp32 flat gradient (CPU path / baseline transport)."""
        if g16 is None:
            return g32
        out = g32.clone()
        for i, s in enumerate(self.layout.specs):
            if s.is_bf16:
                self.layout.view(out, i).copy_(self.layout.view(g16, i))
        return out
