"""EMA teacher update -- replacement for the inline loop at main_dino_mc.py:403-406:

    for param_q, param_k in zip(student.module.parameters(), teacher_without_ddp.parameters()):
        param_k.data.mul_(m).add_((1 - m) * param_q.detach().data)

as one multi-tensor kernel launch.  The loop has no seam in the reference, so the drop-in launcher
(dropin.py) patches `train_one_epoch`; standalone callers use `ema_update_` directly.

The same pass also refreshes the GEMM operands of teacher heads that asked for it (`DINOHead` in no-grad mode with
frozen parameters, bf16-GEMM mode): bf16 copies of the new MLP weights and the weight-normed last layer's
W = g v/||v|| -- the update streams every teacher parameter anyway, and the next teacher forward then starts straight
with its GEMMs.
"""
from __future__ import annotations

import weakref

import torch

from . import ops

_plans = {}
_shadow_heads = weakref.WeakSet()       # DINOHead modules holding operand shadows (head.py registers them)


def register_shadow_head(head):
    _shadow_heads.add(head)


def _shadow_spec(tp):
    """(shadows {index: bf16 tensor}, wn tuple or None, heads) for the registered heads whose parameters are in `tp`."""
    if not _shadow_heads:
        return {}, None, []
    index = {p.data_ptr(): i for i, p in enumerate(tp)}
    shadows, wn, heads = {}, None, []
    for head in list(_shadow_heads):
        sh = head._shadow
        if sh is None:
            continue
        lin_idx = [index.get(w.data_ptr()) for w in sh["weights"]]
        iv, ig = index.get(head.last_layer.weight_v.data_ptr()), index.get(head.last_layer.weight_g.data_ptr())
        if any(i is None for i in lin_idx) or iv is None or ig is None or wn is not None:
            continue                    # not (entirely) part of this update, or a second weight-normed layer: no shadows
        for i, t in zip(lin_idx, sh["mlp"]):
            shadows[i] = t
        wn = (iv, ig, head.last_layer.in_features, sh["what"], sh["scale"], sh["inv_norm"])
        heads.append(head)
    return shadows, wn, heads


@torch.no_grad()
def ema_update_(teacher_params, student_params, m):
    """In-place p_k <- p_k*m + (1-m)*p_q over the zipped lists (teacher first, like `param_k`), bit-exact
    with the reference's three fp32 roundings.  The chunk table is cached while storages do not move."""
    tparams = list(teacher_params)
    tp = [p.data for p in tparams]
    sp = [p.data for p in student_params]
    n = min(len(tp), len(sp))
    shadows, wn, heads = _shadow_spec(tp[:n])
    key = (tuple((a.data_ptr(), b.data_ptr(), a.numel()) for a, b in zip(tp[:n], sp[:n])),
           tuple(sorted((i, t.data_ptr()) for i, t in shadows.items())), None if wn is None else (wn[0], wn[1], wn[3].data_ptr()))
    plan = _plans.get(key)
    if plan is None:
        if len(_plans) > 16:
            _plans.clear()
        plan = ops.EmaPlan(tp, sp, shadows=shadows, wn=wn)
        _plans[key] = plan
    plan.run(float(m))
    for head in heads:                  # the shadows now describe exactly the parameter values the kernel wrote
        head._mark_shadow_fresh()
    return plan
