"""CUDA-graph capture of a whole head/loss/EMA step.

Every libdinomc entry point only enqueues work on the caller's stream, never synchronises and never
allocates (scratch comes from torch's pool through `ops.workspace`), so a step built from the drop-in
modules -- forward, loss, backward, `ema_update_` -- is capturable as is.  State that the eager modules carry by
re-binding (DINOLoss.center, main_dino_mc.py:473) is written back in place at the end of the captured step, so
consecutive replays evolve it exactly like consecutive eager steps (tests/test_gpu_modules.py).  Replaying the graph removes the
per-launch host cost (~50 launches per step), which is what bounds the eager path.

    step = dinomc_b200.StepGraph(lambda: run_one_step(static_inputs))   # warm-up + capture
    static_inputs.copy_(new_batch, non_blocking=True)                    # refill the static buffers
    loss = step.replay(momentum=momentum_schedule[it])                   # same tensors, new values

Host scalars and a captured graph (the reference changes them while it trains):
  * EMA momentum (`momentum_schedule[it]`, main_dino_mc.py:404, a new value EVERY iteration): not baked in -- the EMA
    kernel reads (m, 1-m) from device memory and `replay(momentum=m)` refills it before the launch;
  * teacher temperature (`teacher_temp_schedule[epoch]`, main_dino_mc.py:445, changes once per epoch during warm-up):
    baked in as a kernel argument; `replay(epoch=e)` compares the temperature the captured DINOLoss modules would use
    at epoch e with the captured one and RE-CAPTURES when it differs (needs `fn` to take an `epoch` keyword), or
    raises when `fn` cannot be re-run with another epoch;
  * optimizer steps (bias correction depends on the step count) are refused during capture.

The warm-up calls of `fn` run with `ops.preserve_state` set: all kernels run, but the EMA uses m = 1 (exact identity),
DINOLoss leaves its center alone and optimizers refuse to step -- teacher, center and optimizer state are the same
before and after constructing a StepGraph.
"""
from __future__ import annotations

import inspect

import torch

from . import ops


class StepGraph:
    def __init__(self, fn, warmup: int = 3, capture_error_mode: str = "global", epoch=None, stream=None):
        """`fn()` (or `fn(epoch=...)`) must read its inputs from fixed (static) device tensors; its return value
        (tensor or tuple of tensors) is kept as the static output of the graph."""
        self.fn = fn
        self.capture_error_mode = capture_error_mode
        self.stream = stream                    # capture stream (e.g. a high-priority one); None = torch's default side stream
        try:
            self._takes_epoch = "epoch" in inspect.signature(fn).parameters
        except (TypeError, ValueError):
            self._takes_epoch = False
        self.epoch = epoch if epoch is not None else (0 if self._takes_epoch else None)
        side = stream if stream is not None else torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        ops.preserve_state = True
        try:
            with torch.cuda.stream(side):       # warm-up off the default stream: allocator pools, EMA plan, workspaces
                for _ in range(max(warmup, 1)):
                    self._call()
        finally:
            ops.preserve_state = False
        torch.cuda.current_stream().wait_stream(side)
        self._capture()

    def _call(self):
        return self.fn(epoch=self.epoch) if self._takes_epoch else self.fn()

    def _capture(self):
        from .loss import commit_captured_centers, drop_pending_events
        torch.cuda.synchronize()
        drop_pending_events()                   # everything has completed: no stale cross-stream events into the capture
        self.graph = torch.cuda.CUDAGraph()
        ops.capture_notes = []
        try:
            with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode=self.capture_error_mode):
                self.out = self._call()
                commit_captured_centers()       # DINOLoss.center: new value back into the buffer the step reads
            notes = ops.capture_notes
        finally:
            ops.capture_notes = None
        torch.cuda.synchronize()
        drop_pending_events()                   # events recorded inside the capture are not usable outside it
        self._ema = [(n[1], n[2]) for n in notes if n[0] == "ema"]
        self._losses = [(n[1], n[2], n[3]) for n in notes if n[0] == "loss"]
        for plan, m in self._ema:               # the device scalars still hold the warm-up's m = 1
            plan.set_momentum(m)

    def replay(self, momentum=None, epoch=None):
        """Launch the captured step.  `momentum`: this iteration's EMA momentum (refills the device scalars when it
        changed); `epoch`: this iteration's epoch (re-captures when the teacher temperature differs from the captured one)."""
        if epoch is not None and epoch != self.epoch:
            stale = any(float(mod.teacher_temp_schedule[epoch]) != temp for mod, _, temp in self._losses)
            if stale and not self._takes_epoch:
                raise RuntimeError("StepGraph: the captured step uses the teacher temperature of epoch "
                                   f"{self._losses[0][1]}; epoch {epoch} needs another one.  Build the StepGraph from a "
                                   "function with an `epoch` keyword so it can be re-captured.")
            self.epoch = epoch
            if stale:
                ms = [plan.m_on_device for plan, _ in self._ema]
                self._capture()
                for (plan, _), m in zip(self._ema, ms):
                    if m is not None:
                        plan.set_momentum(m)
        if momentum is not None:
            for plan, _ in self._ema:
                if plan.m_on_device != float(momentum):
                    plan.set_momentum(momentum)
        self.graph.replay()
        return self.out
