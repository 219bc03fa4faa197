"""Between backward and the optimizer: the two gradient utilities main_dino_mc.py calls every step
(:383-400), SURVEY 8(f) rank 1.

    clip_gradients(model, clip)                                utils/utils.py:145-154
    cancel_gradients_last_layer(epoch, model, freeze_last_layer)   utils/utils.py:157-162

`clip_gradients` keeps the reference's per-PARAMETER semantics (each gradient is scaled by clip / (||g|| + 1e-6) when
that is below 1) but runs as two multi-tensor kernel launches over all gradients and returns the norms as ONE device
tensor: no `.item()` per parameter, hence no host synchronisation (the reference's list of Python floats is
`norms.tolist()`).  `cancel_gradients_last_layer` needs no kernel: it drops the gradients, exactly like the reference.
"""
from __future__ import annotations

import torch

from . import ops

_plans = {}


@torch.no_grad()
def clip_gradients(model, clip) -> torch.Tensor:
    """In-place per-parameter clipping of every existing gradient of `model` (an nn.Module or an iterable of
    parameters), in `named_parameters()` order like the reference.  Returns the pre-clip L2 norms (fp32, on the device)."""
    params = model.parameters() if hasattr(model, "parameters") else model
    grads = [p.grad.data for p in params if p.grad is not None]
    if not grads:
        return torch.zeros(0)
    for i, g in enumerate(grads):
        if not g.is_cuda:
            raise RuntimeError("dinomc_b200 has no CPU path: gradients must be CUDA tensors")
        if g.dtype != torch.float32 or not g.is_contiguous():
            raise TypeError("clip_gradients: gradients must be contiguous float32 tensors")
    key = tuple((g.data_ptr(), g.numel()) for g in grads)
    plan = _plans.get(key)
    if plan is None:
        if len(_plans) > 16:
            _plans.clear()
        plan = ops.ClipPlan(grads)
        _plans[key] = plan
    return plan.run(float(clip))


def cancel_gradients_last_layer(epoch, model, freeze_last_layer):
    """utils/utils.py:157-162: during the first `freeze_last_layer` epochs the last layer receives no update."""
    if epoch >= freeze_last_layer:
        return
    for n, p in model.named_parameters():
        if "last_layer" in n:
            p.grad = None
