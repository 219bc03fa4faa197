"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- torch-CPU port of the reference path.

The reference's hot path is eager PyTorch; this file restates it with the same torch ops,
in the same order and with the same number of passes over the logits, on CPU tensors of any
dtype, so that (a) it can be the loop-faithful second oracle that travels to the GPU box and
(b) timing it on the box's host cores is representative of "the reference's CPU path"
(bench.py `cpu_baseline` with kind="port", and `bench.py --impl reference`).

Reference lines followed (relative to /root/reference):
  head ......... utils/vision_transformer.py:260-294
  loss ......... main_dino_mc.py:419-473
  EMA .......... main_dino_mc.py:403-406

Parity pin: tests/golden/*.npz (outputs of the real reference, see oracle/gen_golden.py).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


def make_head_params(in_dim, out_dim, nlayers=3, hidden_dim=2048, bottleneck_dim=256,
                     norm_last_layer=True, seed=0, dtype=torch.float32):
    """Parameters with the reference's names/order/shapes (utils/vision_transformer.py:261-282).
    Init statistics follow the reference (MLP weights ~N(0,.02^2), biases 0, weight_v ~ default
    nn.Linear init U(-1/sqrt(fan_in), 1/sqrt(fan_in)), weight_g = 1) but NOT its RNG stream --
    for bit-identical weights load a reference state_dict instead."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    nlayers = max(nlayers, 1)
    if nlayers == 1:
        dims = [(in_dim, bottleneck_dim)]
        names = ["mlp"]
    else:
        dims = [(in_dim, hidden_dim)] + [(hidden_dim, hidden_dim)] * (nlayers - 2) + [(hidden_dim, bottleneck_dim)]
        names = [f"mlp.{2 * i}" for i in range(nlayers)]
    for n, (fi, fo) in zip(names, dims):
        p[n + ".weight"] = (torch.randn(fo, fi, generator=g, dtype=torch.float32) * 0.02).to(dtype)
        p[n + ".bias"] = torch.zeros(fo, dtype=dtype)
    p["last_layer.weight_g"] = torch.ones(out_dim, 1, dtype=dtype)
    bound = 1.0 / bottleneck_dim ** 0.5
    p["last_layer.weight_v"] = ((torch.rand(out_dim, bottleneck_dim, generator=g, dtype=torch.float32) * 2 - 1) * bound).to(dtype)
    for k, v in p.items():
        v.requires_grad_(not (k.endswith("weight_g") and norm_last_layer))
    return p


def head_forward(x, p):
    """utils/vision_transformer.py:290-294 (use_bn=False)."""
    if "mlp.weight" in p:
        x = F.linear(x, p["mlp.weight"], p["mlp.bias"])
    else:
        idx = sorted(int(k.split(".")[1]) for k in p if k.startswith("mlp.") and k.endswith(".weight"))
        for j, i in enumerate(idx):
            x = F.linear(x, p[f"mlp.{i}.weight"], p[f"mlp.{i}.bias"])
            if j < len(idx) - 1:
                x = F.gelu(x)
    x = F.normalize(x, dim=-1, p=2)
    v, g = p["last_layer.weight_v"], p["last_layer.weight_g"]
    w = v * (g / v.norm(dim=1, keepdim=True))          # weight_norm pre-hook, recomputed every forward
    return F.linear(x, w)


class LossState:
    """center buffer + temperature schedule of DINOLoss (main_dino_mc.py:420-435)."""

    def __init__(self, out_dim, ncrops, warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs,
                 nepochs, teacher_crops_number=2, student_temp=0.1, center_momentum=0.9, dtype=torch.float32):
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.ncrops = ncrops
        self.teacher_crops_number = teacher_crops_number
        self.center = torch.zeros(1, out_dim, dtype=dtype)
        self.teacher_temp_schedule = np.concatenate((
            np.linspace(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs),
            np.ones(nepochs - warmup_teacher_temp_epochs) * teacher_temp))


def loss_forward(state: LossState, student_output, teacher_output, epoch, world_size=1, all_reduce=None):
    """main_dino_mc.py:437-473: the literal pair loop, then the center update (old center in the loss)."""
    student_out = (student_output / state.student_temp).chunk(state.ncrops)
    temp = state.teacher_temp_schedule[epoch]
    teacher_out = F.softmax((teacher_output - state.center) / temp, dim=-1).detach().chunk(state.teacher_crops_number)
    total, n = 0, 0
    for iq, q in enumerate(teacher_out):
        for v in range(len(student_out)):
            if v == iq:
                continue
            loss = torch.sum(-q * F.log_softmax(student_out[v], dim=-1), dim=-1)
            total = total + loss.mean()
            n += 1
    total = total / n
    with torch.no_grad():
        bc = torch.sum(teacher_output, dim=0, keepdim=True)
        if all_reduce is not None:
            all_reduce(bc)
        bc = bc / (len(teacher_output) * world_size)
        state.center = state.center * state.center_momentum + bc * (1 - state.center_momentum)
    return total


@torch.no_grad()
def ema_update(teacher_params, student_params, m):
    """main_dino_mc.py:403-406."""
    for pq, pk in zip(student_params, teacher_params):
        pk.data.mul_(m).add_((1 - m) * pq.detach().data)


@torch.no_grad()
def clip_gradients(grads, clip):
    """utils/utils.py:145-154 on a list of gradient tensors (the reference walks `model.named_parameters()` and
    skips parameters without a gradient).  Clips in place; returns the pre-clip norms as Python floats."""
    norms = []
    for g in grads:
        param_norm = g.data.norm(2)
        norms.append(param_norm.item())
        clip_coef = clip / (param_norm + 1e-6)
        if clip_coef < 1:
            g.data.mul_(clip_coef)
    return norms


@torch.no_grad()
def lars_step(params, grads, mus, lr, weight_decay, momentum=0.9, eta=0.001):
    """utils/utils.py:584-608 (`LARS.step`) for one parameter group, on plain tensor lists: `params` and the momentum
    buffers `mus` are updated in place.  Parameters with ndim == 1 get neither weight decay nor the trust ratio."""
    for p, dp, mu in zip(params, grads, mus):
        if dp is None:
            continue
        if p.ndim != 1:
            dp = dp.add(p, alpha=weight_decay)
            param_norm = torch.norm(p)
            update_norm = torch.norm(dp)
            one = torch.ones_like(param_norm)
            q = torch.where(param_norm > 0., torch.where(update_norm > 0, (eta * param_norm / update_norm), one), one)
            dp = dp.mul(q)
        mu.mul_(momentum).add_(dp)
        p.add_(mu, alpha=-lr)


def step(x_student, x_teacher, student_p, teacher_p, state: LossState, epoch, ema_m, ema_extra=None):
    """One whole step of the path as SURVEY.md section 8d defines it: teacher head fwd (no grad),
    student head fwd, loss (+center), backward to head params and features, EMA over head (+extra
    backbone-shaped) params.  Returns (loss, grads dict)."""
    with torch.no_grad():
        t_out = head_forward(x_teacher, teacher_p)
    x_student = x_student.detach().requires_grad_(True)
    s_out = head_forward(x_student, student_p)
    loss = loss_forward(state, s_out, t_out, epoch)
    wrt = [x_student] + [v for v in student_p.values() if v.requires_grad]
    names = ["x"] + [k for k, v in student_p.items() if v.requires_grad]
    grads = dict(zip(names, torch.autograd.grad(loss, wrt)))
    ema_update(list(teacher_p.values()), list(student_p.values()), ema_m)
    if ema_extra is not None:
        ema_update(ema_extra[0], ema_extra[1], ema_m)
    return loss.detach(), grads
