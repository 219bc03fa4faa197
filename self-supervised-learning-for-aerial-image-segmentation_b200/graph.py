"""CUDA-graph capture of a whole head/loss/EMA step.

Every libdinomc entry point only enqueues work on the caller's stream, never synchronises and never
allocates (scratch comes from torch's pool through `ops.workspace`), so a step built from the drop-in
modules -- forward, loss, backward, `ema_update_` -- is capturable as is.  State that the eager modules carry by
re-binding (DINOLoss.center, main_dino_mc.py:473) is written back in place at the end of the captured step, so
consecutive replays evolve it exactly like consecutive eager steps (tests/test_gpu_modules.py).  Replaying the graph removes the
per-launch host cost (~50 launches per step), which is what bounds the eager path.

    step = dinomc_b200.StepGraph(lambda: run_one_step(static_inputs))   # warm-up + capture
    static_inputs.copy_(new_batch, non_blocking=True)                    # refill the static buffers
    loss = step.replay()                                                  # same tensors, new values
"""
from __future__ import annotations

import torch


class StepGraph:
    def __init__(self, fn, warmup: int = 3, capture_error_mode: str = "global"):
        """`fn()` must read its inputs from fixed (static) device tensors; its return value (tensor or tuple of
        tensors) is kept as the static output of the graph."""
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):           # warm-up off the default stream: allocator pools, EMA plan, workspaces
            for _ in range(max(warmup, 1)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from .loss import drop_pending_events
        drop_pending_events()                   # everything has completed: no stale cross-stream events into the capture
        self.graph = torch.cuda.CUDAGraph()
        from .loss import commit_captured_centers
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.out = fn()
            commit_captured_centers()           # DINOLoss.center: new value back into the buffer the step reads
        torch.cuda.synchronize()
        drop_pending_events()                   # events recorded inside the capture are not usable outside it

    def replay(self):
        self.graph.replay()
        return self.out
