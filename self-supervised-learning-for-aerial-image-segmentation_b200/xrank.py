"""Cross-rank exchange over NVLink / NVSwitch peer memory: symmetric buffers + libdinomc's own all-reduce kernel.

`torch.distributed._symmetric_memory` is used as plumbing only: it allocates a buffer at the same virtual address
layout on every rank, maps every rank's buffer into every process and (on NVSwitch systems) binds them to one
multicast address.  The all-reduce itself is `dmc_xrank_allreduce` (csrc/xrank.cu): one kernel, two cross-rank barriers
on signal pads, `multimem.ld_reduce` / `multimem.st` through the switch -- or peer loads / stores when there is no
multicast address.  It replaces the NCCL all-reduce behind DDP's gradient averaging (main_dino_mc.py:260).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops

_PAD_OFFSET = 16384          # bytes: torch's own barrier channels live at the start of a signal pad; ours start here
_PAD_BYTES = 1 << 17


def available() -> bool:
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
        return dist.is_available() and dist.is_initialized()
    except Exception:       # noqa: BLE001
        return False


class SymmetricBuffer:
    """A flat buffer that exists on every rank of `group` in symmetric memory, with an in-place all-reduce."""

    def __init__(self, numel: int, dtype: torch.dtype, group=None, ctas: int = 148):
        import os
        import sys
        import torch.distributed._symmetric_memory as symm
        if os.environ.get("DMC_XRANK_DEBUG"):
            print(f"[xrank] allocating SymmetricBuffer({numel}, {dtype}) capturing={torch.cuda.is_current_stream_capturing()}", file=sys.stderr, flush=True)
        if dtype not in (torch.bfloat16, torch.float32):
            raise TypeError("SymmetricBuffer: float32 or bfloat16")
        group = group if group is not None else dist.group.WORLD
        lib = L.load()
        self.ctas = int(ctas)
        world = dist.get_world_size(group)
        need = _PAD_OFFSET + lib.dmc_xrank_signal_bytes(world, self.ctas)
        if need > _PAD_BYTES:
            raise ValueError("SymmetricBuffer: too many CTAs for the signal pad")
        try:
            if symm.get_signal_pad_size() < _PAD_BYTES:
                symm.set_signal_pad_size(_PAD_BYTES)          # must precede the first symmetric allocation
        except Exception:       # noqa: BLE001 -- older API: the default pad is checked below
            pass
        try:
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:       # noqa: BLE001 -- not needed (or deprecated) on newer torch
            pass
        esz = 2 if dtype == torch.bfloat16 else 4
        per16 = 16 // esz
        self.numel = (int(numel) + per16 - 1) // per16 * per16      # whole 16-byte vectors
        dev = torch.device("cuda", torch.cuda.current_device())
        self.tensor = symm.empty(self.numel, dtype=dtype, device=dev)
        self.tensor.zero_()
        self.handle = symm.rendezvous(self.tensor, group=group)
        self.rank, self.world = int(self.handle.rank), int(self.handle.world_size)
        if int(self.handle.signal_pad_size) < need:
            raise RuntimeError(f"symmetric-memory signal pad too small: {self.handle.signal_pad_size} < {need}")
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        self.multicast = mc != 0
        self._mc = C.c_void_p(mc if mc else None)
        self._peers = (L.vp * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        self._pads = (L.vp * self.world)(*[int(p) + _PAD_OFFSET for p in self.handle.signal_pad_ptrs])
        self.handle.barrier()                                   # every rank has zeroed its buffer and mapped the others
        torch.cuda.synchronize()

    def allreduce_(self, scale: float = 1.0, widen_to=None, widen_offsets=None):
        """In place on every rank: buffer <- scale * sum over ranks (enqueued on the current stream).  `widen_to`: up to
        eight fp32 tensors that receive element ranges of the finished bf16 buffer (starting at `widen_offsets`)."""
        lib = L.load()
        n_out = 0
        outs = offs = ns = None
        if widen_to:
            n_out = len(widen_to)
            for t in widen_to:
                ops._need_cuda(t)
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise TypeError("allreduce_: widen_to takes contiguous float32 tensors")
            outs = (L.vp * n_out)(*[t.data_ptr() for t in widen_to])
            offs = (L.i64 * n_out)(*[int(o) for o in widen_offsets])
            ns = (L.i64 * n_out)(*[t.numel() for t in widen_to])
        with ops._timed("xrank_allreduce"):
            L.check(lib.dmc_xrank_allreduce(self._mc, self._peers, self._pads, self.numel, ops._dt(self.tensor), self.rank,
                                            self.world, float(scale), self.ctas, n_out, outs, offs, ns, ops._stream()),
                    "dmc_xrank_allreduce")
        ops._count()
        return self.tensor
