"""GPU: the drop-in training step EXECUTED -- `dropin.train_one_epoch` (the seam for main_dino_mc.py:356-416) drives a toy
backbone + the drop-in DINOHead / DINOLoss / clip / EMA over a list-of-crops loader for a few iterations, and the result
is compared with the same loop written against the oracle (oracle/torch_port.py, float64, the reference's op order:
schedules -> teacher / student forward -> loss (+ center) -> backward -> per-parameter clip -> optimizer -> EMA)."""
import types

import numpy as np
import pytest
import torch
from torch import nn

from conftest import rel_err

pytestmark = pytest.mark.gpu

D_FEAT, K, HID, BOT = 64, 512, 128, 64
B, G, C = 4, 2, 6
STEPS = 3


class ToyBackbone(nn.Module):
    """Stands in for a ViT / ResNet: any crop resolution -> [n, D_FEAT] (global average pool + Linear)."""

    def __init__(self):
        super().__init__()
        self.proj = nn.Linear(3, D_FEAT)
        self.fc, self.head = nn.Linear(1, 1), nn.Linear(1, 1)       # replaced by Identity in MultiCropWrapper

    def forward(self, x):
        return self.proj(x.mean((2, 3)))


def _loader(seed):
    g = torch.Generator().manual_seed(seed)
    batches = []
    for _ in range(STEPS):
        crops = [torch.randn(B, 3, 32, 32, generator=g) for _ in range(G)] + [torch.randn(B, 3, 16, 16, generator=g) for _ in range(C - G)]
        batches.append((crops, None))
    return batches


def _schedules():
    lr = np.linspace(0.05, 0.02, STEPS)
    wd = np.linspace(0.0, 0.0, STEPS)
    mom = np.linspace(0.9, 0.95, STEPS)
    return lr, wd, mom


def _reference_loop(student_sd, teacher_sd, batches, clip, dev="cuda"):
    """float64 on the device, oracle functions, the reference's order of operations."""
    from oracle import torch_port as T
    sp = {k: v.detach().double().to(dev).clone().requires_grad_(True) for k, v in student_sd.items()}
    tp = {k: v.detach().double().to(dev).clone() for k, v in teacher_sd.items()}
    head_s = {k[len("head."):]: v for k, v in sp.items() if k.startswith("head.")}
    head_t = {k[len("head."):]: v for k, v in tp.items() if k.startswith("head.")}
    head_s["last_layer.weight_g"].requires_grad_(False)
    st = T.LossState(K, C, 0.04, 0.04, 0, 10, teacher_crops_number=G, dtype=torch.float64)
    st.center = st.center.to(dev)
    lr, wd, mom = _schedules()
    losses = []

    def features(p, crops):
        outs = []
        for grp in (crops[:G], crops[G:]):
            if grp:
                x = torch.cat(grp).double().to(dev)
                outs.append(torch.nn.functional.linear(x.mean((2, 3)), p["backbone.proj.weight"], p["backbone.proj.bias"]))
        return torch.cat(outs)

    for it, (crops, _) in enumerate(batches):
        with torch.no_grad():
            t_out = T.head_forward(features(tp, crops[:G]), head_t)
        s_out = T.head_forward(features(sp, crops), head_s)
        loss = T.loss_forward(st, s_out, t_out, 0)
        losses.append(float(loss))
        train = [(k, v) for k, v in sp.items() if v.requires_grad]
        grads = torch.autograd.grad(loss, [v for _, v in train])
        grads = [g.clone() for g in grads]
        if clip:
            T.clip_gradients(grads, clip)
        with torch.no_grad():
            for (k, v), g in zip(train, grads):
                v.add_(g, alpha=-float(lr[it]))                     # SGD, no momentum, no weight decay
        T.ema_update(list(tp.values()), list(sp.values()), float(mom[it]))
    return losses, sp, tp, st.center


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("host_sync", ["reference", "deferred"])
def test_train_one_epoch_executes_and_matches_oracle(mode, tol, host_sync):
    import dinomc_b200 as D
    from dinomc_b200 import dropin
    torch.manual_seed(0)
    student = D.MultiCropWrapper(ToyBackbone(), D.DINOHead(D_FEAT, K, hidden_dim=HID, bottleneck_dim=BOT)).cuda()
    teacher = D.MultiCropWrapper(ToyBackbone(), D.DINOHead(D_FEAT, K, hidden_dim=HID, bottleneck_dim=BOT)).cuda()
    teacher.load_state_dict(student.state_dict())                  # main_dino_mc.py:262
    for p in teacher.parameters():
        p.requires_grad = False
    student.head.precision = teacher.head.precision = mode
    student_sd = {k: v.detach().clone() for k, v in student.named_parameters()}
    teacher_sd = {k: v.detach().clone() for k, v in teacher.named_parameters()}
    loss_mod = D.DINOLoss(K, C, 0.04, 0.04, 0, 10, teacher_crops_number=G).cuda()
    opt = torch.optim.SGD([p for p in student.parameters() if p.requires_grad], lr=0.0)
    lr, wd, mom = _schedules()
    args = types.SimpleNamespace(epochs=1, global_crops_number=G, clip_grad=0.3, freeze_last_layer=0)
    batches = _loader(5)
    meters = dropin._PlainMeters()
    stats = dropin.train_one_epoch(student, teacher, teacher, loss_mod, batches, opt, lr, wd, mom, 0, None, args,
                                   meters=meters, host_sync=host_sync)
    torch.cuda.synchronize()
    ref_losses, sp, tp, center = _reference_loop(student_sd, teacher_sd, batches, 0.3)
    assert abs(stats["loss"] - float(np.mean(ref_losses))) / abs(float(np.mean(ref_losses))) < tol
    assert stats["lr"] == pytest.approx(float(np.mean(lr)))
    for k, v in student.named_parameters():
        assert rel_err(v.detach().cpu().numpy(), sp[k].detach().cpu().numpy()) < tol, k
    for k, v in teacher.named_parameters():
        assert rel_err(v.detach().cpu().numpy(), tp[k].detach().cpu().numpy()) < tol, k
    assert rel_err(loss_mod.center.cpu().numpy(), center.cpu().numpy()) < (1e-5 if mode == "fp32" else 3e-3)
