"""Symmetric (peer-mapped) memory and NVLS multicast for one-process-per-GPU jobs.

Policy layer over ``csrc/cuda/rt_symm.cpp``.  A rank allocates exportable HBM regions, ships their file descriptors
to peers over ``AF_UNIX`` sockets (``SCM_RIGHTS``), and maps what it receives; kernels then read/write peer memory with
plain loads/stores.  A multicast object bound to every rank's parameter region gives the PS a single address whose
stores the NVSwitch replicates to all GPUs (``multimem.st``).

This replaces the reference's mpi4py transport (every ``comm.isend/Isend/irecv/Bcast`` in
src/master/baseline_master.py:156-200 and src/worker/baseline_worker.py:163-273): after setup there is no message
passing at all on the step path -- only memory traffic and flag words.

torch.distributed (NCCL or Gloo) is used for *bootstrapping only* (exchanging sizes / a job id / barriers).
"""
from __future__ import annotations

import ctypes as C
import os
import socket
import struct
import uuid
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .. import _native as N


class _CAI:
    """Expose a raw device pointer through ``__cuda_array_interface__`` so torch can alias it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def tensor_from_ptr(ptr: int, nbytes: int, device: torch.device, dtype=torch.uint8) -> torch.Tensor:
    t = torch.as_tensor(_CAI(ptr, nbytes), device=device)
    return t.view(dtype)


@dataclass
class Region:
    name: str
    ptr: int
    size: int
    handle: int = 0
    fd: int = -1
    tensor: Optional[torch.Tensor] = None     # uint8 view (local regions only)


class SymmContext:
    """Per-process manager of exported / imported regions."""

    def __init__(self, device: torch.device, rank: int = 0, world: int = 1, group=None):
        self.device = torch.device(device)
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.rank, self.world, self.group = rank, world, group
        self.lib = N.cuda()
        N.check(self.lib.drc_rt_init(self.dev_index), "rt_init")
        g = N.u64()
        N.check(self.lib.drc_rt_granularity(self.dev_index, C.byref(g)), "rt_granularity")
        self.granularity = int(g.value)
        self.local: Dict[str, Region] = {}
        self.remote: Dict[str, Dict[int, Region]] = {}
        self.mc_ptr: Dict[str, int] = {}
        self._sock = None
        self._job = None
        if world > 1:
            self._open_listener()

    # ------------------------------------------------------------------ allocation
    def _round(self, nbytes: int, gran: Optional[int] = None) -> int:
        g = gran or self.granularity
        return (nbytes + g - 1) // g * g

    def alloc(self, name: str, nbytes: int, gran: Optional[int] = None) -> Region:
        """Allocate an exportable, zero-filled region on this rank."""
        size = self._round(max(nbytes, 1), gran)
        if self.world == 1:
            t = torch.zeros(size, dtype=torch.uint8, device=self.device)
            reg = Region(name, t.data_ptr(), size, tensor=t)
        else:
            p, h, fd = N.ptr(), N.u64(), C.c_int(-1)
            N.check(self.lib.drc_rt_alloc(self.dev_index, size, C.byref(p), C.byref(h), C.byref(fd)), f"rt_alloc({name})")
            t = tensor_from_ptr(p.value, size, self.device)
            t.zero_()
            reg = Region(name, p.value, size, h.value, fd.value, t)
        self.local[name] = reg
        return reg

    # ------------------------------------------------------------------ fd exchange
    def _sock_path(self, rank: int) -> str:
        return f"/tmp/draco_b200_{self._job}_{rank}.sock"

    def _open_listener(self) -> None:
        obj = [uuid.uuid4().hex[:12] if self.rank == 0 else None]
        dist.broadcast_object_list(obj, src=0, group=self.group)
        self._job = obj[0]
        path = self._sock_path(self.rank)
        if os.path.exists(path):
            os.unlink(path)
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        s.bind(path)
        s.listen(256)
        self._sock = s
        dist.barrier(group=self.group)

    def _send_fd(self, dst: int, fd: int, tag: str) -> None:
        with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
            c.connect(self._sock_path(dst))
            hdr = struct.pack("<i", self.rank) + tag.encode()[:60].ljust(60, b"\0")
            socket.send_fds(c, [hdr], [fd])

    def _recv_fds(self, count: int, tag: str) -> Dict[int, int]:
        out: Dict[int, int] = {}
        while len(out) < count:
            conn, _ = self._sock.accept()
            with conn:
                msg, fds, _, _ = socket.recv_fds(conn, 64, 1)
            src = struct.unpack("<i", msg[:4])[0]
            got = msg[4:].rstrip(b"\0").decode()
            assert got == tag[:60], f"fd exchange out of order: expected {tag}, got {got}"
            out[src] = fds[0]
        return out

    def share(self, name: str, exporters: List[int], importers: List[int]) -> Dict[int, Region]:
        """Make region ``name`` of every rank in ``exporters`` addressable on every rank in ``importers``.
        Collective over the whole group.  Returns {exporter rank: Region mapped here}."""
        mapped: Dict[int, Region] = {}
        if self.rank in exporters:
            mapped[self.rank] = self.local[name]
        if self.world > 1:
            sizes = [None] * self.world
            dist.all_gather_object(sizes, self.local[name].size if self.rank in exporters else 0, group=self.group)
            if self.rank in exporters:
                for dst in importers:
                    if dst != self.rank:
                        self._send_fd(dst, self.local[name].fd, name)
            if self.rank in importers:
                expect = [e for e in exporters if e != self.rank]
                fds = self._recv_fds(len(expect), name)
                for src, fd in fds.items():
                    p, h = N.ptr(), N.u64()
                    N.check(self.lib.drc_rt_import(self.dev_index, fd, sizes[src], C.byref(p), C.byref(h)), f"rt_import({name})")
                    self.lib.drc_rt_close_fd(fd)
                    mapped[src] = Region(name, p.value, sizes[src], h.value)
            dist.barrier(group=self.group)
        self.remote[name] = mapped
        return mapped

    # ------------------------------------------------------------------ multicast
    def multicast_supported(self) -> bool:
        if self.world == 1:
            return False
        v = C.c_int(0)
        if self.lib.drc_rt_mc_supported(self.dev_index, C.byref(v)) != 0:
            return False
        flags = [None] * self.world
        dist.all_gather_object(flags, int(v.value), group=self.group)
        return all(flags)

    def bind_multicast(self, name: str) -> Optional[int]:
        """Bind region ``name`` of *every* rank to one multicast object; returns the multicast device pointer here
        (stores to it land in all ranks' regions), or None when NVLS is unavailable.  Collective."""
        if not self.multicast_supported():
            return None
        size = self.local[name].size
        ok = 1
        mc, fd = N.u64(), C.c_int(-1)
        try:
            if self.rank == 0:
                N.check(self.lib.drc_rt_mc_create(self.world, size, C.byref(mc), C.byref(fd)), "mc_create")
                for dst in range(1, self.world):
                    self._send_fd(dst, fd.value, "mc:" + name)
            else:
                got = self._recv_fds(1, "mc:" + name)
                N.check(self.lib.drc_rt_mc_import(got[0], C.byref(mc)), "mc_import")
                self.lib.drc_rt_close_fd(got[0])
            N.check(self.lib.drc_rt_mc_add_device(mc.value, self.dev_index), "mc_add_device")
        except RuntimeError:
            ok = 0
        oks = [None] * self.world
        dist.all_gather_object(oks, ok, group=self.group)       # also: every device added before anyone binds
        if not all(oks):
            return None
        p = N.ptr()
        try:
            N.check(self.lib.drc_rt_mc_bind(mc.value, 0, self.local[name].handle, 0, size), "mc_bind")
            N.check(self.lib.drc_rt_mc_map(self.dev_index, mc.value, size, C.byref(p)), "mc_map")
        except RuntimeError:
            ok = 0
        dist.all_gather_object(oks, ok, group=self.group)
        if not all(oks):
            return None
        self.mc_ptr[name] = p.value
        return p.value

    def close(self) -> None:
        if self._sock is not None:
            try:
                self._sock.close()
                os.unlink(self._sock_path(self.rank))
            except OSError:
                pass
            self._sock = None
