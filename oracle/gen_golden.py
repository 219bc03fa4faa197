"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the REAL reference modules.

Run in the build container (needs /root/reference):

    python -m oracle.gen_golden

For every case the reference's own `DINOHead` (utils/vision_transformer.py:260),
`DINOLoss` (main_dino_mc.py:419) and EMA loop (main_dino_mc.py:403-406) are executed on seeded
synthetic inputs, once in float32 (the reference's numerics) and once in float64 (ground truth),
and inputs + outputs are stored.  The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_loader  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> config.  Small on purpose: the whole directory stays a few MB.
CASES = {
    # DINO-MC layout: 2 global + 6 local crops
    "mc_small": dict(in_dim=48, out_dim=512, hidden_dim=64, bottleneck_dim=32, nlayers=3, norm_last_layer=True,
                     B=4, ncrops=8, G=2, warmup_tt=0.04, tt=0.04, warmup_epochs=0, nepochs=10, epoch=0),
    # DINO-TP layout: 3 global + 6 local crops, trainable weight_g, warm-up temperature mid-ramp,
    # out_dim not a power of two
    "tp_small": dict(in_dim=40, out_dim=384, hidden_dim=96, bottleneck_dim=64, nlayers=3, norm_last_layer=False,
                     B=3, ncrops=9, G=3, warmup_tt=0.04, tt=0.07, warmup_epochs=5, nepochs=10, epoch=2),
    # nlayers == 1 (bare Linear named "mlp"), global crops only
    "one_layer": dict(in_dim=24, out_dim=256, hidden_dim=32, bottleneck_dim=16, nlayers=1, norm_last_layer=True,
                      B=5, ncrops=2, G=2, warmup_tt=0.04, tt=0.04, warmup_epochs=0, nepochs=4, epoch=1),
    # tile-sized dims for the tensor-core kernel (bottleneck a multiple of 64, out_dim of 128)
    "mc_wide": dict(in_dim=64, out_dim=768, hidden_dim=128, bottleneck_dim=128, nlayers=3, norm_last_layer=True,
                    B=4, ncrops=8, G=2, warmup_tt=0.04, tt=0.04, warmup_epochs=0, nepochs=10, epoch=3),
}


def run_case(cfg, dtype):
    main_dino_mc, vits, ref_utils = reference_loader.load()
    reference_loader.ensure_process_group()
    B, C, G, K = cfg["B"], cfg["ncrops"], cfg["G"], cfg["out_dim"]
    gen = torch.Generator().manual_seed(1234)
    x_s = torch.randn(C * B, cfg["in_dim"], generator=gen)
    x_t = torch.randn(G * B, cfg["in_dim"], generator=gen)
    center0 = torch.randn(1, K, generator=gen) * 0.3

    def make_head(seed):
        torch.manual_seed(seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            h = vits.DINOHead(cfg["in_dim"], K, use_bn=False, norm_last_layer=cfg["norm_last_layer"],
                              nlayers=cfg["nlayers"], hidden_dim=cfg["hidden_dim"],
                              bottleneck_dim=cfg["bottleneck_dim"])
        # exercise the weight-norm gain: the reference fills weight_g with 1; perturb it so g matters
        with torch.no_grad():
            h.last_layer.weight_g.mul_(1.0 + 0.25 * torch.rand(K, 1, generator=gen))
            for n, p in h.named_parameters():
                if n.endswith(".bias"):
                    p.add_(0.05 * torch.randn(p.shape, generator=gen))
        return h

    student, teacher = make_head(0), make_head(1)
    student_sd32 = {k: v.detach().clone() for k, v in student.state_dict().items()}
    teacher_sd32 = {k: v.detach().clone() for k, v in teacher.state_dict().items()}
    student, teacher = student.to(dtype), teacher.to(dtype)
    for p in teacher.parameters():
        p.requires_grad = False
    loss_mod = main_dino_mc.DINOLoss(K, C, cfg["warmup_tt"], cfg["tt"], cfg["warmup_epochs"], cfg["nepochs"],
                                     teacher_crops_number=G).to(dtype)
    loss_mod.center.copy_(center0.to(dtype))

    xs = x_s.to(dtype).requires_grad_(True)
    with torch.no_grad():
        t_out = teacher(x_t.to(dtype))
    s_out = student(xs)
    s_out.retain_grad()
    loss1 = loss_mod(s_out, t_out, cfg["epoch"])
    loss1.backward()
    center1 = loss_mod.center.detach().clone()
    # second evaluation on the same logits: uses the UPDATED center (main_dino_mc.py:459-460 ordering)
    with torch.no_grad():
        loss2 = loss_mod(s_out.detach(), t_out, cfg["epoch"])
    center2 = loss_mod.center.detach().clone()

    out = {
        "student_logits": s_out.detach(), "teacher_logits": t_out, "loss1": loss1.detach(), "loss2": loss2,
        "center1": center1, "center2": center2, "dlogits": s_out.grad, "grad.x": xs.grad,
    }
    for n, p in student.named_parameters():
        if p.grad is not None:
            out["grad." + n] = p.grad
    # EMA exactly as main_dino_mc.py:403-406, momentum from the reference's cosine schedule
    sched = ref_utils.cosine_scheduler(0.996, 1, 10, 7)
    m = sched[5]
    with torch.no_grad():
        for param_q, param_k in zip(student.parameters(), teacher.parameters()):
            param_k.data.mul_(m).add_((1 - m) * param_q.detach().data)
    for n, p in teacher.named_parameters():
        out["ema." + n] = p.detach()
    ins = {"x_student": x_s, "x_teacher": x_t, "center0": center0, "ema_m": torch.tensor(m, dtype=torch.float64),
           "temp": torch.tensor(loss_mod.teacher_temp_schedule[cfg["epoch"]], dtype=torch.float64)}
    for k, v in student_sd32.items():
        ins["student." + k] = v
    for k, v in teacher_sd32.items():
        ins["teacher." + k] = v
    return ins, out


def main():
    if not reference_loader.available():
        raise SystemExit("reference not available; fixtures can only be generated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, cfg in CASES.items():
        ins, out32 = run_case(cfg, torch.float32)
        _, out64 = run_case(cfg, torch.float64)
        blob = {"cfg_keys": np.array(list(cfg.keys())), "cfg_vals": np.array([float(v) for v in cfg.values()])}
        for k, v in ins.items():
            blob["in." + k] = v.detach().numpy()
        # float32 run: what the reference itself produces (EMA is an fp32-exact contract);
        # float64 run: ground truth for everything differentiable.  Large per-weight arrays are
        # kept once (ema.* only in fp32, grad.* only in fp64) to keep the fixtures small.
        for k, v in out32.items():
            if not k.startswith("grad.last_layer") and not k.startswith("grad.mlp"):
                blob["ref32." + k] = v.detach().numpy().astype(np.float32)
        for k, v in out64.items():
            if not k.startswith("ema."):
                blob["ref64." + k] = v.detach().numpy().astype(np.float64)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: loss32={float(out32['loss1']):.9f} loss64={float(out64['loss1']):.12f} -> {path} "
              f"({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
