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
    copies = []                       # (original, contiguous fp32 stand-in): e.g. channels_last conv gradients of a ResNet
    for i, g in enumerate(grads):
        if not g.is_cuda:
            raise RuntimeError("dinomc_b200 has no CPU path: gradients must be CUDA tensors")
        if g.dtype != torch.float32 or not g.is_contiguous():
            c = g.to(dtype=torch.float32, memory_format=torch.contiguous_format, copy=True)
            copies.append((g, c))
            grads[i] = c
    if copies:                        # stand-ins move every call: no plan caching, clip, copy back
        norms = ops.ClipPlan(grads).run(float(clip))
        for g, c in copies:
            g.copy_(c)
        return norms
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


def _refuse_inside_graph(name):
    if ops.preserve_state or torch.cuda.is_current_stream_capturing():
        raise RuntimeError(f"{name}.step() cannot be part of a StepGraph: its update depends on host-side state (step count, "
                           "scheduled lr / weight decay) that a captured graph would freeze.  Call it outside the captured step.")


class FusedAdamW(torch.optim.Optimizer):
    """Drop-in for `torch.optim.AdamW(params_groups)` as main_dino_mc.py:281-282 builds it (lr / weight decay rewritten
    per iteration by the schedules at :363-367): same constructor defaults, same `param_groups` keys, same per-parameter
    state (`step`, `exp_avg`, `exp_avg_sq`), so `state_dict()` / `load_state_dict()` interchange with the torch
    optimizer and with the reference's checkpoints (`main_dino_mc.py:336`).  `step()` is one multi-tensor launch per
    parameter group.  fp32 CUDA parameters only; `amsgrad` / `maximize` / `capturable` are not supported (the reference
    uses none of them)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdamW: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._plans = {}

    @torch.no_grad()
    def step(self, closure=None):
        _refuse_inside_graph("FusedAdamW")
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                if not p.is_cuda:
                    raise RuntimeError("dinomc_b200 has no CPU path: FusedAdamW parameters must be CUDA tensors")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdamW does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)          # torch.optim.AdamW's layout
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            # parameters that have stepped the same number of times share a launch (normally: the whole group)
            by_step = {}
            for p in params:
                st = self.state[p]
                st["step"] += 1
                by_step.setdefault(int(st["step"].item()), []).append(p)
            if group.get("amsgrad") or group.get("maximize"):
                raise RuntimeError("FusedAdamW does not implement amsgrad / maximize (the reference uses neither)")
            beta1, beta2 = group["betas"]
            for t, ps in by_step.items():
                grads = [p.grad.data if p.grad.is_contiguous() else p.grad.data.contiguous() for p in ps]
                key = (gi, tuple((p.data_ptr(), g.data_ptr(), self.state[p]["exp_avg"].data_ptr()) for p, g in zip(ps, grads)))
                plan = self._plans.get(key)
                if plan is None:
                    if len(self._plans) > 8:
                        self._plans.clear()
                    plan = ops.AdamWPlan([p.data for p in ps], grads, [self.state[p]["exp_avg"] for p in ps],
                                         [self.state[p]["exp_avg_sq"] for p in ps])
                    self._plans[key] = plan
                plan.run(group["lr"], beta1, beta2, group["eps"], group["weight_decay"], t)
        return loss


class FusedLARS(torch.optim.Optimizer):
    """Drop-in for the reference's `utils.LARS(params_groups)` (utils/utils.py:570-608, main_dino_mc.py:285-286): same
    constructor and defaults, same `param_groups` keys (lr / weight_decay rewritten per iteration by the schedules at
    main_dino_mc.py:363-367), same per-parameter state (`mu`), so `state_dict()` / `load_state_dict()` interchange with
    the reference optimizer and its checkpoints.  `step()` is two multi-tensor launches per parameter group (norms, then
    update) with the trust ratio formed on the device: no per-parameter norm kernels, `where`s or temporaries.  Like the
    reference, parameters with `ndim == 1` (biases, norm scales) get neither weight decay nor the LARS adaptation, and
    the two filter arguments are accepted and ignored.  fp32 CUDA parameters only."""

    def __init__(self, params, lr=0, weight_decay=0, momentum=0.9, eta=0.001, weight_decay_filter=None,
                 lars_adaptation_filter=None):
        defaults = dict(lr=lr, weight_decay=weight_decay, momentum=momentum, eta=eta, weight_decay_filter=weight_decay_filter,
                        lars_adaptation_filter=lars_adaptation_filter)
        super().__init__(params, defaults)
        self._plans = {}

    @torch.no_grad()
    def step(self):
        _refuse_inside_graph("FusedLARS")
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                if not p.is_cuda:
                    raise RuntimeError("dinomc_b200 has no CPU path: FusedLARS parameters must be CUDA tensors")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedLARS does not support sparse gradients")
                st = self.state[p]
                if "mu" not in st:
                    st["mu"] = torch.zeros_like(p)
            grads = [p.grad.data if p.grad.is_contiguous() else p.grad.data.contiguous() for p in params]
            key = (gi, tuple((p.data_ptr(), g.data_ptr(), self.state[p]["mu"].data_ptr()) for p, g in zip(params, grads)))
            plan = self._plans.get(key)
            if plan is None:
                if len(self._plans) > 8:
                    self._plans.clear()
                plan = ops.LarsPlan([p.data for p in params], grads, [self.state[p]["mu"] for p in params])
                self._plans[key] = plan
            plan.run(group["lr"], group["weight_decay"], group["momentum"], group["eta"])
