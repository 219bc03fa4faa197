"""Host logic above the C ABI, on CPU: the real DINOHead / DINOLoss / autograd Functions run with the kernels replaced by the
torch restatements of tests/_ops_double.py, and the whole step is compared with the fp64 numpy oracle.  Catches wiring
mistakes (operand layouts, epilogue requests, saved tensors, gradient arity, the head -> loss statistics hand-over) in the
container that has no GPU; the kernels themselves are covered by the `-m gpu` suite."""
import numpy as np
import pytest
import torch

import dinomc_b200 as D
from oracle import np_oracle as O

import _ops_double as dbl  # noqa: E402  (tests/ is on sys.path: rootdir conftest)


def _rel(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def _build(mode, K=1024, in_dim=64, B=4, C=8, G=2, norm_last_layer=True, seed=0, **kw):
    torch.manual_seed(seed)
    kw = dict(dict(hidden_dim=128, bottleneck_dim=64), **kw)
    student = D.DINOHead(in_dim, K, norm_last_layer=norm_last_layer, **kw)
    teacher = D.DINOHead(in_dim, K, norm_last_layer=norm_last_layer, **kw)
    for p in teacher.parameters():
        p.requires_grad = False
    student.precision = teacher.precision = mode
    with torch.no_grad():
        for p in list(student.parameters()) + list(teacher.parameters()):
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
        if not norm_last_layer:
            student.last_layer.weight_g.mul_(1.0 + 0.1 * torch.randn_like(student.last_layer.weight_g))
    loss_mod = D.DINOLoss(K, C, 0.04, 0.04, 0, 10, teacher_crops_number=G)
    with torch.no_grad():
        loss_mod.center.normal_(0, 0.3)
    xs = torch.randn(C * B, in_dim, requires_grad=True)
    xt = torch.randn(G * B, in_dim)
    return student, teacher, loss_mod, xs, xt


def _oracle(student, teacher, center0, xs, xt, C, G):
    ssd = {k: v.detach().numpy() for k, v in student.state_dict().items()}
    tsd = {k: v.detach().numpy() for k, v in teacher.state_dict().items()}
    return O.full_step(xs.detach().numpy(), xt.numpy(), ssd, tsd, center0.numpy(), 0.04, C, G, ema_m=0.996)


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("norm_last_layer", [True, False])
def test_step_wiring_against_the_oracle(monkeypatch, mode, tol, norm_last_layer):
    dbl.install(monkeypatch)
    C, G = 8, 2
    student, teacher, loss_mod, xs, xt = _build(mode, norm_last_layer=norm_last_layer)
    center0 = loss_mod.center.clone()
    ref = _oracle(student, teacher, center0, xs, xt, C, G)
    with torch.no_grad():
        t_out = teacher(xt)
    s_out = student(xs)
    assert getattr(s_out, "_dmc_stats", None) is not None and s_out._dmc_stats["kind"] == "student"      # fused statistics attached
    assert getattr(t_out, "_dmc_stats", None) is None
    loss = loss_mod(s_out, t_out, 0)
    assert "ce_fused" in dbl.calls and "ce_fwd" not in dbl.calls            # the fused route ran ...
    loss.backward()
    assert "ce_bwd" not in dbl.calls                                        # ... and its gradient was reused, not recomputed
    errs = {"loss": abs(float(loss.detach()) - ref["loss"]) / abs(ref["loss"]), "x": _rel(xs.grad.numpy(), ref["grads"]["x"])}
    for name, p in student.named_parameters():
        if p.requires_grad:
            assert p.grad is not None, name
            errs[name] = _rel(p.grad.numpy(), ref["grads"][name])
        else:
            assert p.grad is None, name
    assert all(e < tol for e in errs.values()), errs
    assert _rel(loss_mod.center.numpy(), ref["center"]) < (1e-6 if mode == "fp32" else 2e-3)
    assert loss_mod.center.data_ptr() != center0.data_ptr()                 # rebound like the reference (main_dino_mc.py:473)
    # the teacher built no autograd graph and saved nothing: its MLP GEMMs were launched without the second (gelu') output
    assert not t_out.requires_grad


def test_plain_route_when_the_statistics_do_not_match(monkeypatch):
    """Logits that did not come from a head bound to this loss (here: a clone) take the separate-pass route and give the same
    numbers; a second DINOLoss with another student temperature invalidates the head's statistics instead of using them."""
    dbl.install(monkeypatch)
    C, G = 8, 2
    student, teacher, loss_mod, xs, xt = _build("fp32")
    center0 = loss_mod.center.clone()
    ref = _oracle(student, teacher, center0, xs, xt, C, G)
    with torch.no_grad():
        t_out = teacher(xt)
    s_out = student(xs)
    plain = s_out.clone()                                                    # no _dmc_stats attribute
    loss = loss_mod(plain, t_out, 0)
    assert "ce_fwd" in dbl.calls and "ce_fused" not in dbl.calls
    loss.backward()
    assert "ce_bwd" in dbl.calls
    assert abs(float(loss.detach()) - ref["loss"]) / abs(ref["loss"]) < 2e-5
    assert _rel(xs.grad.numpy(), ref["grads"]["x"]) < 2e-5
    # another loss module, other temperature: the statistics the head attached (computed for `loss_mod`) must be rejected
    dbl.calls.clear()
    other = D.DINOLoss(1024, C, 0.04, 0.04, 0, 10, teacher_crops_number=G, student_temp=0.2)
    student.bind_loss(loss_mod)
    s2 = student(xs)
    assert s2._dmc_stats["scale"] == pytest.approx(1.0 / loss_mod.student_temp)
    other(s2, t_out, 0)
    assert "ce_fwd" in dbl.calls and "ce_fused" not in dbl.calls


def test_explicit_binding_survives_a_later_loss_module(monkeypatch):
    """Two DINOLoss modules alive: an unbound head follows the most recent one, a bound head keeps its own."""
    dbl.install(monkeypatch)
    C, G = 8, 2
    student, teacher, loss_a, xs, xt = _build("fp32")
    loss_b = D.DINOLoss(1024, C, 0.04, 0.04, 0, 10, teacher_crops_number=G, student_temp=0.25)     # constructed later: the default
    assert student(xs)._dmc_stats["scale"] == pytest.approx(4.0)
    student.bind_loss(loss_a)
    s_out = student(xs)
    assert s_out._dmc_stats["scale"] == pytest.approx(10.0)
    with torch.no_grad():
        t_out = teacher(xt)
    dbl.calls.clear()
    loss_a(s_out, t_out, 0)
    assert "ce_fused" in dbl.calls
    assert loss_b is not None


def test_nlayers_one_and_upstream_gradient_scale(monkeypatch):
    """nlayers = 1 (a bare Linear as `mlp`) and a non-unit upstream gradient (what a GradScaler delivers): the fused route
    rescales its stored gradient."""
    dbl.install(monkeypatch)
    C, G = 8, 2
    student, teacher, loss_mod, xs, xt = _build("fp32", nlayers=1)
    center0 = loss_mod.center.clone()
    ref = _oracle(student, teacher, center0, xs, xt, C, G)
    with torch.no_grad():
        t_out = teacher(xt)
    loss = loss_mod(student(xs), t_out, 0)
    (loss * 8.0).backward()
    assert "scale_if" in dbl.calls
    assert _rel(xs.grad.numpy() / 8.0, ref["grads"]["x"]) < 2e-5
    assert _rel(student.mlp.weight.grad.numpy() / 8.0, ref["grads"]["mlp.weight"]) < 2e-5


def test_dropin_train_one_epoch_host_logic(monkeypatch):
    """`dropin.train_one_epoch` (the seam for main_dino_mc.py:356-416) executed on CPU over the kernel double: schedules written into
    the optimizer, MultiCropWrapper over a list of crops of two resolutions, loss, backward, per-parameter clip, optimizer step,
    EMA with the scheduled momentum -- against the same loop written with oracle/torch_port.py in float64."""
    import types
    import test_gpu_dropin as G
    from dinomc_b200 import dropin
    dbl.install(monkeypatch)
    torch.manual_seed(0)
    student = D.MultiCropWrapper(G.ToyBackbone(), D.DINOHead(G.D_FEAT, G.K, hidden_dim=G.HID, bottleneck_dim=G.BOT))
    teacher = D.MultiCropWrapper(G.ToyBackbone(), D.DINOHead(G.D_FEAT, G.K, hidden_dim=G.HID, bottleneck_dim=G.BOT))
    teacher.load_state_dict(student.state_dict())
    for p in teacher.parameters():
        p.requires_grad = False
    student.head.precision = teacher.head.precision = "fp32"
    student_sd = {k: v.detach().clone() for k, v in student.named_parameters()}
    teacher_sd = {k: v.detach().clone() for k, v in teacher.named_parameters()}
    loss_mod = D.DINOLoss(G.K, G.C, 0.04, 0.04, 0, 10, teacher_crops_number=G.G)
    opt = torch.optim.SGD([p for p in student.parameters() if p.requires_grad], lr=0.0)
    lr, wd, mom = G._schedules()
    args = types.SimpleNamespace(epochs=1, global_crops_number=G.G, clip_grad=0.3, freeze_last_layer=0)
    batches = G._loader(5)
    stats = dropin.train_one_epoch(student, teacher, teacher, loss_mod, batches, opt, lr, wd, mom, 0, None, args,
                                   meters=dropin._PlainMeters(), host_sync="reference")
    assert dbl.calls.count("ema") == G.STEPS and dbl.calls.count("clip_grads") == G.STEPS
    ref_losses, sp, tp, center = G._reference_loop(student_sd, teacher_sd, batches, 0.3, dev="cpu")
    tol = 2e-5
    assert abs(stats["loss"] - float(np.mean(ref_losses))) / abs(float(np.mean(ref_losses))) < tol
    assert stats["lr"] == pytest.approx(float(np.mean(lr)))
    for k, v in student.named_parameters():
        assert _rel(v.detach().numpy(), sp[k].detach().numpy()) < tol, k
    for k, v in teacher.named_parameters():
        assert _rel(v.detach().numpy(), tp[k].detach().numpy()) < tol, k
    assert _rel(loss_mod.center.numpy(), center.numpy()) < 1e-5


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_shapes_tma_cannot_express_take_the_ffma_route(monkeypatch, mode):
    """Row strides that are not multiples of 16 bytes (in_dim 63, out_dim 1001) cannot be described by a TMA tensor map: every GEMM of
    the head must then be requested from the FFMA kernel (`simt=True`), the normalize backward runs as its own launch, no statistics
    are fused -- and the numbers are the oracle's."""
    dbl.install(monkeypatch)
    seen = []
    real_gemm = dbl.gemm

    def spy(*a, **k):
        seen.append((k.get("tag"), bool(k.get("simt"))))
        return real_gemm(*a, **k)

    monkeypatch.setattr(D.ops, "gemm", spy)
    C, G = 8, 2
    student, teacher, loss_mod, xs, xt = _build(mode, K=1001, in_dim=63)
    center0 = loss_mod.center.clone()
    ref = _oracle(student, teacher, center0, xs, xt, C, G)
    with torch.no_grad():
        t_out = teacher(xt)
    s_out = student(xs)
    assert getattr(s_out, "_dmc_stats", None) is None and s_out.dtype == torch.float32
    loss = loss_mod(s_out, t_out, 0)
    loss.backward()
    assert seen and all(simt for _, simt in seen), seen
    assert "normalize_bwd" in dbl.calls and "ce_fwd" in dbl.calls and "ce_bwd" in dbl.calls
    assert abs(float(loss.detach()) - ref["loss"]) / abs(ref["loss"]) < 2e-5
    assert _rel(xs.grad.numpy(), ref["grads"]["x"]) < 2e-5
    assert _rel(student.last_layer.weight_v.grad.numpy(), ref["grads"]["last_layer.weight_v"]) < 2e-5


def test_use_bn_head_wiring(monkeypatch):
    """`use_bn_in_head` (utils/vision_transformer.py:268-274): every Linear goes through `LinearFn` (own GEMM), BatchNorm1d and GELU
    stay torch modules between them; forward, every gradient and the running statistics against the same head in float64 torch."""
    import copy
    import torch.nn.functional as F
    dbl.install(monkeypatch)
    torch.manual_seed(11)
    head = D.DINOHead(64, 1024, use_bn=True, norm_last_layer=False, hidden_dim=128, bottleneck_dim=64)
    assert [type(m).__name__ for m in head.mlp] == ["Linear", "BatchNorm1d", "GELU", "Linear", "BatchNorm1d", "GELU", "Linear"]
    head.precision = "fp32"
    with torch.no_grad():
        head.last_layer.weight_g.uniform_(0.5, 1.5)
    ref_mlp = copy.deepcopy(head.mlp).double()
    g64 = head.last_layer.weight_g.detach().double().requires_grad_(True)
    v64 = head.last_layer.weight_v.detach().double().requires_grad_(True)
    x = torch.randn(48, 64, generator=torch.Generator().manual_seed(12))
    up = torch.randn(48, 1024, generator=torch.Generator().manual_seed(13))
    xg = x.clone().requires_grad_(True)
    out = head(xg)
    assert [c for c in dbl.calls if c.startswith("gemm_mlp_fwd")] == ["gemm_mlp_fwd_64x128", "gemm_mlp_fwd_128x128", "gemm_mlp_fwd_128x64"]
    (out.float() * up).sum().backward()
    x64 = x.double().requires_grad_(True)
    z = F.normalize(ref_mlp(x64), dim=-1, p=2)
    ref = z @ (v64 * (g64 / v64.norm(dim=1, keepdim=True))).t()
    (ref * up.double()).sum().backward()
    tol = 2e-5
    assert _rel(out.detach().numpy(), ref.detach().numpy()) < tol
    assert _rel(xg.grad.numpy(), x64.grad.numpy()) < tol
    assert _rel(head.last_layer.weight_v.grad.numpy(), v64.grad.numpy()) < tol
    assert _rel(head.last_layer.weight_g.grad.numpy(), g64.grad.numpy()) < tol
    for (n, p), q in zip(head.mlp.named_parameters(), ref_mlp.parameters()):
        if n not in ("0.bias", "3.bias"):      # a bias in front of a BatchNorm has an exactly-zero gradient
            assert _rel(p.grad.numpy(), q.grad.numpy()) < tol, n
    assert _rel(head.mlp[1].running_var.numpy(), ref_mlp[1].running_var.numpy()) < 1e-5


def test_dino_tp_layout_and_warmup_temperature(monkeypatch):
    """DINO-TP (main_dino_tp.py layout: 3 teacher crops, 9 crops in all, trainable gain) at an epoch inside the teacher-temperature
    warm-up: DINOLoss must take schedule[epoch] (main_dino_mc.py:445) and pass the crop layout through unchanged."""
    dbl.install(monkeypatch)
    C, G, K, B = 9, 3, 1024, 3
    torch.manual_seed(5)
    student = D.DINOHead(64, K, norm_last_layer=False, hidden_dim=128, bottleneck_dim=64)
    teacher = D.DINOHead(64, K, norm_last_layer=False, hidden_dim=128, bottleneck_dim=64)
    for p in teacher.parameters():
        p.requires_grad = False
    student.precision = teacher.precision = "fp32"
    loss_mod = D.DINOLoss(K, C, 0.04, 0.07, 3, 10, teacher_crops_number=G)
    temps = [float(t) for t in loss_mod.teacher_temp_schedule[:4]]
    assert temps == pytest.approx([0.04, 0.055, 0.07, 0.07])
    with torch.no_grad():
        loss_mod.center.normal_(0, 0.3)
    center0 = loss_mod.center.clone()
    xs = torch.randn(C * B, 64, requires_grad=True)
    xt = torch.randn(G * B, 64)
    ssd = {k: v.detach().numpy() for k, v in student.state_dict().items()}
    tsd = {k: v.detach().numpy() for k, v in teacher.state_dict().items()}
    ref = O.full_step(xs.detach().numpy(), xt.numpy(), ssd, tsd, center0.numpy(), temps[1], C, G)
    with torch.no_grad():
        t_out = teacher(xt)
    loss = loss_mod(student(xs), t_out, 1)
    loss.backward()
    assert abs(float(loss.detach()) - ref["loss"]) / abs(ref["loss"]) < 2e-5
    assert _rel(xs.grad.numpy(), ref["grads"]["x"]) < 2e-5
    assert _rel(student.last_layer.weight_g.grad.numpy(), ref["grads"]["last_layer.weight_g"]) < 2e-5
    assert _rel(loss_mod.center.numpy(), ref["center"]) < 1e-6
    with pytest.raises(ValueError):
        loss_mod(student(xs)[:-1], t_out, 1)                 # rows must be ncrops * B
