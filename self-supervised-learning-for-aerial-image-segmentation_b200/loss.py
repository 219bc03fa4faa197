"""DINOLoss -- drop-in for the reference's `main_dino_mc.DINOLoss` (main_dino_mc.py:419-473): same
constructor, attributes, `center` buffer / state_dict, forward(student_output, teacher_output, epoch)
returning a 0-dim loss with autograd to student_output, and the center update (with its cross-rank
all-reduce) as a side effect AFTER the loss, on the libdinomc kernels.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import functional as Fn
from . import ops
from .head import wait_ready


import weakref

_async_center = False
_instances = weakref.WeakSet()          # live DINOLoss modules (StepGraph drops their pending events before capture)


def drop_pending_events():
    """After a device-wide synchronize: forget events of completed asynchronous center exchanges.  An event
    recorded inside one CUDA-graph capture must not be waited on from eager code or from another capture."""
    for m in list(_instances):
        m._center_event = None


# (module, center tensor at the start of the captured step): see commit_captured_centers
_capture_commits = []


def commit_captured_centers():
    """Called by StepGraph at the END of the captured step (still inside the capture).  The reference rebinds
    `self.center` to a new tensor every step (main_dino_mc.py:473), which a replayed graph cannot express: its kernels
    read and write fixed addresses.  So inside a capture the step's new center is copied back, after the backward pass
    (which may still read the old values), into the buffer the step READ its center from, and the module is re-bound
    to that buffer: every replay then sees the center the previous replay left, exactly like the eager loop."""
    done = set()
    for mod, persistent in _capture_commits:
        if id(mod) in done:
            continue
        done.add(id(mod))
        mod.sync_center()                       # an asynchronous exchange must have produced the new center
        new = mod.center
        if new.data_ptr() != persistent.data_ptr():
            with torch.no_grad():
                persistent.copy_(new)
            mod.center = persistent
    _capture_commits.clear()


def set_async_center(enabled: bool):
    """Multi-GPU only.  When enabled, the center exchange (column-sum all-reduce + EMA, main_dino_mc.py:468-473)
    runs on a side stream and the next `DINOLoss.forward` waits for it, taking the latency-bound 256 KiB
    all-reduce off the step's critical path (its result is first needed by the NEXT step's teacher softmax).
    Code that reads `loss.center` directly in between must call `loss.sync_center()` first."""
    global _async_center
    _async_center = bool(enabled)


class DINOLoss(nn.Module):
    def __init__(self, out_dim, ncrops, warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs, nepochs,
                 teacher_crops_number=2, student_temp=0.1, center_momentum=0.9):
        super().__init__()
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.ncrops = ncrops
        self.teacher_crops_number = teacher_crops_number
        self.register_buffer("center", torch.zeros(1, out_dim))
        self._center_event = None
        self._comm_stream = None
        self.center_exchange = None     # optional xrank.SymmetricBuffer (fp32, >= out_dim): the column-sum all-reduce (:469) then
                                        # runs on libdinomc's own NVLink / NVSwitch kernel instead of NCCL
        _instances.add(self)
        # same schedule construction as main_dino_mc.py:431-435
        self.teacher_temp_schedule = np.concatenate((
            np.linspace(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs),
            np.ones(nepochs - warmup_teacher_temp_epochs) * teacher_temp,
        ))
        # temperature the teacher head's fused statistics assume until the first forward tells otherwise
        self._last_inv_tt = 1.0 / float(self.teacher_temp_schedule[0]) if len(self.teacher_temp_schedule) else 25.0
        Fn.register_loss(self)

    @staticmethod
    def _common(student_output, teacher_output):
        """Bring both logit tensors to one kernel dtype (fp32 or bf16) without touching values needlessly."""
        if student_output.dtype == torch.bfloat16 and teacher_output.dtype == torch.bfloat16:
            return student_output, teacher_output
        return student_output.float(), teacher_output.float()

    def forward(self, student_output, teacher_output, epoch):
        if not student_output.is_cuda:
            raise RuntimeError("dinomc_b200.DINOLoss runs on CUDA (sm_100a) only; there is no CPU fallback")
        C, G = self.ncrops, self.teacher_crops_number
        if student_output.shape[0] % C or teacher_output.shape[0] % G:
            raise ValueError("student/teacher rows must be ncrops*B / teacher_crops_number*B (crop-major)")
        B = student_output.shape[0] // C
        if teacher_output.shape[0] // G != B:
            raise ValueError("student and teacher batches differ")
        temp = float(self.teacher_temp_schedule[epoch])
        if ops.capture_notes is not None and torch.cuda.is_current_stream_capturing():
            ops.capture_notes.append(("loss", self, epoch, temp))       # StepGraph re-captures when the temperature changes
        self.sync_center()
        wait_ready(teacher_output)              # no-op unless the teacher head ran on the overlap side stream
        s, t = self._common(student_output, teacher_output.detach())
        inv_ts, inv_tt = 1.0 / self.student_temp, 1.0 / temp
        # statistics the heads' GEMM epilogues may have attached to the logits; re-validated here
        s_pre = getattr(student_output, "_dmc_stats", None) if s is student_output else None
        if s_pre is not None and not (s_pre.get("kind") == "student" and s_pre["scale"] == inv_ts
                                      and s_pre["row_partials"].shape[0] == s.shape[0]):
            s_pre = None
        raw_pre = getattr(teacher_output, "_dmc_stats", None)
        if raw_pre is not None and raw_pre.get("event") is not None:
            # statistics a teacher head launched on a side stream: ALWAYS join that stream here, used or not (inside a CUDA
            # graph capture an unjoined side stream is an error; outside it keeps the allocator's stream bookkeeping simple)
            torch.cuda.current_stream().wait_event(raw_pre["event"])
        t_pre = raw_pre if t.data_ptr() == teacher_output.data_ptr() else None
        if t_pre is not None and not (t_pre.get("kind") in ("teacher", "teacher_final") and t_pre["scale"] == inv_tt
                                      and t_pre["center_ptr"] == self.center.data_ptr()
                                      and t_pre["center_version"] == self.center._version
                                      and (t_pre["row_partials"].shape[0] if t_pre["kind"] == "teacher" else t_pre["rows"]) == t.shape[0]):
            t_pre = None
        if t_pre is not None:
            # the statistics may have been allocated on another stream (side-stream statistics pass, or a teacher head that
            # ran on the overlap stream): tell the caching allocator that this stream reads them too
            cur = torch.cuda.current_stream()
            for key in ("t_stats", "colsum", "row_partials", "colsum_partials"):
                buf = t_pre.get(key)
                if buf is not None:
                    buf.record_stream(cur)
        self._last_inv_tt = inv_tt
        Fn.register_loss(self)
        loss, colsum = Fn.DinoLossFn.apply(s, t, self.center, inv_ts, inv_tt, B, C, G, s_pre, t_pre)
        self._update_center_from_colsum(colsum, teacher_output.shape[0])
        return loss

    def _allreduce_colsum(self, colsum):
        """SUM of the per-GPU column sums over the ranks (main_dino_mc.py:469), on the current stream."""
        buf = self.center_exchange
        if buf is None:
            dist.all_reduce(colsum)
            return colsum
        k = colsum.numel()
        view = buf.tensor[:k]
        view.copy_(colsum.reshape(-1))
        buf.allreduce_(1.0)
        return view

    def sync_center(self):
        """Make the current stream wait for an in-flight asynchronous center exchange (no-op otherwise)."""
        if self._center_event is not None:
            torch.cuda.current_stream().wait_event(self._center_event)
            self._center_event = None

    @torch.no_grad()
    def _update_center_from_colsum(self, colsum, n_rows):
        if ops.preserve_state:
            # StepGraph warm-up: run the same kernels (and collectives) but keep the center as it is
            keep = self.center
            try:
                ops.preserve_state = False
                self._update_center_from_colsum(colsum, n_rows)
                self.sync_center()
            finally:
                ops.preserve_state = True
                self.center = keep
            return
        if torch.cuda.is_current_stream_capturing() and not any(m is self for m, _ in _capture_commits):
            _capture_commits.append((self, self.center))
        world = 1
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size()
        if world > 1 and _async_center:
            cur = torch.cuda.current_stream()
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream()
            comm = self._comm_stream
            comm.wait_stream(cur)                       # colsum and the old center were produced on `cur`
            with torch.cuda.stream(comm):
                colsum = self._allreduce_colsum(colsum)     # main_dino_mc.py:469
                with ops.no_pdl():                          # behind a cross-rank kernel: do not sit resident while it runs
                    new_center = ops.center_update(self.center, colsum, n_rows * world, self.center_momentum)
            colsum.record_stream(comm)
            self.center.record_stream(comm)
            new_center.record_stream(cur)
            ev = torch.cuda.Event()
            ev.record(comm)
            self._center_event = ev
            self.center = new_center
            return
        if world > 1:
            colsum = self._allreduce_colsum(colsum)     # main_dino_mc.py:469 (65536 fp32 = 256 KiB)
        # rebinding the buffer (like the reference, :473) keeps the old tensor alive for backward
        self.center = ops.center_update(self.center, colsum, n_rows * world, self.center_momentum)

    @torch.no_grad()
    def update_center(self, teacher_output):
        """Public API of the reference (main_dino_mc.py:463-473) for callers that use it directly."""
        wait_ready(teacher_output)
        t = teacher_output.detach()
        t = t if t.dtype == torch.bfloat16 else t.float()
        _, colsum = ops.teacher_stats_colsum(t, self.center.reshape(-1), 1.0)
        self._update_center_from_colsum(colsum, teacher_output.shape[0])
