"""EMA teacher update -- replacement for the inline loop at main_dino_mc.py:403-406:

    for param_q, param_k in zip(student.module.parameters(), teacher_without_ddp.parameters()):
        param_k.data.mul_(m).add_((1 - m) * param_q.detach().data)

as one multi-tensor kernel launch.  The loop has no seam in the reference, so the drop-in launcher
(dropin.py) patches `train_one_epoch`; standalone callers use `ema_update_` directly.
"""
from __future__ import annotations

import torch

from . import ops

_plans = {}


@torch.no_grad()
def ema_update_(teacher_params, student_params, m):
    """In-place p_k <- p_k*m + (1-m)*p_q over the zipped lists (teacher first, like `param_k`), bit-exact
    with the reference's three fp32 roundings.  The chunk table is cached while storages do not move."""
    tp = [p.data for p in teacher_params]
    sp = [p.data for p in student_params]
    n = min(len(tp), len(sp))
    key = tuple((a.data_ptr(), b.data_ptr(), a.numel()) for a, b in zip(tp[:n], sp[:n]))
    plan = _plans.get(key)
    if plan is None:
        if len(_plans) > 16:
            _plans.clear()
        plan = ops.EmaPlan(tp, sp)
        _plans[key] = plan
    plan.run(float(m))
    return plan
