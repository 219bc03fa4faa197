"""Gradient all-reduce for the data-parallel step (the exchange DDP performs at main_dino_mc.py:260).

`GradAllReduce` averages each parameter's gradient over the ranks on a dedicated communication stream so the
NCCL transfers (89 MB of head gradients at K = 65536) overlap the rest of the backward pass and the EMA update:

  * a large gradient (the weight-normed last layer's dv, 64 MB) is reduced the moment the kernel that produced
    it has finished (`ops.mark_ready` event), not when its autograd node returns;
  * the small gradients are batched into one coalesced NCCL launch once all of them exist;
  * while a reducer is active the persistent GEMM grids of the backward pass leave `reserve_sms` SMs free, so the
    NCCL kernel neither queues behind a one-CTA-per-SM GEMM nor starves its tail CTAs.

Unlike `DistributedDataParallel` it keeps no Python-side reducer state between steps, which makes the whole step
(collectives included) capturable in one CUDA graph (`StepGraph`).  `torch.distributed` (NCCL over NVLink /
NVSwitch) does the transport.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import ops

_BIG = 32 << 20      # bytes: gradients at least this large get their own all-reduce
_FLUSH = 1 << 20     # bytes: pending small gradients are sent as one coalesced all-reduce once they reach this


class GradAllReduce:
    def __init__(self, params, group=None, reserve_sms: int = 16):
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        self.comm = torch.cuda.Stream(priority=-1)
        self._pending = []
        self._pending_bytes = 0
        self._seen = 0
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.bytes_per_step = sum(p.numel() * p.element_size() for p in self.params)
        if reserve_sms > 0:
            ops.backward_max_ctas = 148 - reserve_sms
        from . import functional
        functional.grad_exchange_active = os.environ.get("DMC_REDUCER_AUXWN", "") != "1"      # env: timing experiments only
        # early-launched (PDL) GEMM CTAs hold SMs while they wait for their predecessor, which delays the NCCL kernels
        # that share the machine: measured 0.992 -> 0.969 ms per step at 2 GPUs without it
        from . import _lib
        keep_pdl = os.environ.get("DMC_REDUCER_PDL", "") == "1"                               # env: timing experiments only
        self._pdl_prev = _lib.load().dmc_set_pdl(1 if keep_pdl else 0)

    def _hook(self, p):
        g = p.grad
        self._seen += 1
        skip = os.environ.get("DMC_REDUCER_SKIP", "")       # timing experiments only: "big" / "small"
        big = g.numel() * g.element_size() >= _BIG
        if (skip == "big" and big) or (skip == "small" and not big) or skip == "all":
            if self._seen == len(self.params):
                self._flush()
            return
        if big:
            ev = ops.ready_events.pop(g.data_ptr(), None)
            if ev is not None:
                self.comm.wait_event(ev)                 # start as soon as the producing kernel is done
            else:
                self.comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm):
                dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
            g.record_stream(self.comm)
        else:
            self._pending.append((g, ops.ready_events.pop(g.data_ptr(), None)))
            self._pending_bytes += g.numel() * g.element_size()
        if self._seen == len(self.params) or self._pending_bytes >= _FLUSH:
            self._flush()

    def _flush(self):
        if self._pending:
            if all(ev is not None for _, ev in self._pending):
                for _, ev in self._pending:              # start as soon as the producing kernels are done
                    self.comm.wait_event(ev)
            else:
                self.comm.wait_stream(torch.cuda.current_stream())
            grads = [g for g, _ in self._pending]
            with torch.cuda.stream(self.comm):
                try:
                    with dist._coalescing_manager(group=self.group, device=grads[0].device, async_ops=False):
                        for g in grads:
                            dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
                except (AttributeError, TypeError, RuntimeError):
                    for g in grads:
                        dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
            for g in grads:
                g.record_stream(self.comm)
            self._pending = []
            self._pending_bytes = 0
        if self._seen >= len(self.params):
            self._seen = 0

    def wait(self):
        """Join: later work on the current stream sees the averaged gradients."""
        self._flush()
        self._seen = 0
        torch.cuda.current_stream().wait_stream(self.comm)

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
        ops.backward_max_ctas = 0
        from . import functional, _lib
        functional.grad_exchange_active = False
        _lib.load().dmc_set_pdl(self._pdl_prev)
