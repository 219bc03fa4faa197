"""GPU: the drop-in modules (DINOHead, DINOLoss, ema_update_) end to end against the golden vectors
produced by the reference's own modules, plus size-independent properties at the headline size.

Tolerances are the north star's: loss and gradients 1e-5 relative in fp32 mode, 2e-2 in bf16-GEMM mode;
center and EMA parameters 1e-6 (EMA is in fact bit-exact).  "Relative" = max|a-b| / max|b| (conftest.rel_err).
"""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

MODES = [("fp32", 1e-5), ("fp32_simt", 1e-5), ("bf16", 2e-2)]


def _build(golden, mode):
    import dinomc_b200 as D
    c = golden.cfg
    kw = dict(use_bn=False, norm_last_layer=c["norm_last_layer"], nlayers=c["nlayers"], hidden_dim=c["hidden_dim"],
              bottleneck_dim=c["bottleneck_dim"])
    student = D.DINOHead(c["in_dim"], c["out_dim"], **kw).cuda()
    teacher = D.DINOHead(c["in_dim"], c["out_dim"], **kw).cuda()
    student.load_state_dict({k: torch.from_numpy(v) for k, v in golden.sd("student").items()})
    teacher.load_state_dict({k: torch.from_numpy(v) for k, v in golden.sd("teacher").items()})
    student.precision = teacher.precision = mode
    for p in teacher.parameters():
        p.requires_grad = False
    loss = D.DINOLoss(c["out_dim"], c["ncrops"], golden.cfg["warmup_tt"], golden.cfg["tt"], c["warmup_epochs"],
                      c["nepochs"], teacher_crops_number=c["G"]).cuda()
    loss.center.copy_(torch.from_numpy(golden.inputs["center0"]))
    return D, student, teacher, loss


@pytest.mark.parametrize("mode,tol", MODES)
def test_golden_step(golden, mode, tol):
    D, student, teacher, loss_mod = _build(golden, mode)
    c = golden.cfg
    ref = golden.ref64
    xs = torch.from_numpy(golden.inputs["x_student"]).cuda().requires_grad_(True)
    xt = torch.from_numpy(golden.inputs["x_teacher"]).cuda()
    with torch.no_grad():
        t_out = teacher(xt)
    s_out = student(xs)
    assert s_out.shape == (c["ncrops"] * c["B"], c["out_dim"])
    assert rel_err(s_out.detach().float().cpu().numpy(), ref["student_logits"]) < tol
    assert rel_err(t_out.float().cpu().numpy(), ref["teacher_logits"]) < tol
    loss = loss_mod(s_out, t_out, c["epoch"])
    assert loss.dim() == 0 and loss.dtype == torch.float32
    assert abs(float(loss) - float(ref["loss1"])) / abs(float(ref["loss1"])) < tol
    loss.backward()
    assert rel_err(xs.grad.cpu().numpy(), ref["grad.x"]) < tol
    n_checked = 0
    for name, p in student.named_parameters():
        key = "grad." + name
        if key in ref:
            assert p.grad is not None, name
            assert rel_err(p.grad.cpu().numpy().reshape(ref[key].shape), ref[key]) < tol, name
            n_checked += 1
        else:
            assert p.grad is None, name                       # frozen weight_g gets no gradient
    assert n_checked >= 3
    # center: updated AFTER the loss, from the teacher logits (1e-6 in fp32 modes)
    ctol = 1e-6 if mode != "bf16" else 2e-3
    assert loss_mod.center.shape == (1, c["out_dim"])
    assert rel_err(loss_mod.center.cpu().numpy(), ref["center1"]) < ctol
    with torch.no_grad():
        loss2 = loss_mod(s_out.detach(), t_out, c["epoch"])     # second call sees the NEW center
    assert abs(float(loss2) - float(ref["loss2"])) / abs(float(ref["loss2"])) < tol
    assert rel_err(loss_mod.center.cpu().numpy(), ref["center2"]) < ctol
    assert list(loss_mod.state_dict().keys()) == ["center"]
    # EMA over the head parameters, zip order = registration order (bit-exact in every mode)
    D.ema_update_(list(teacher.parameters()), list(student.parameters()), float(golden.inputs["ema_m"]))
    for name, p in teacher.named_parameters():
        assert np.array_equal(p.detach().cpu().numpy(), golden.ref32["ema." + name]), name


def test_auto_precision_follows_autocast(golden):
    if golden.name != "mc_wide":
        pytest.skip("one case is enough")
    D, student, teacher, loss_mod = _build(golden, None)
    xs = torch.from_numpy(golden.inputs["x_student"]).cuda()
    assert student(xs).dtype == torch.float32
    with torch.autocast("cuda", dtype=torch.float16):
        out = student(xs)
    assert out.dtype == torch.bfloat16


def test_grad_scaler_and_amp(golden):
    """The reference's default AMP path (main_dino_mc.py:372,393-400): autocast + GradScaler."""
    if golden.name != "mc_small":
        pytest.skip("one case is enough")
    D, student, teacher, loss_mod = _build(golden, None)
    c = golden.cfg
    xs = torch.from_numpy(golden.inputs["x_student"]).cuda()
    xt = torch.from_numpy(golden.inputs["x_teacher"]).cuda()
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    opt = torch.optim.SGD([p for p in student.parameters() if p.requires_grad], lr=0.0)
    with torch.autocast("cuda", dtype=torch.float16):
        with torch.no_grad():
            t_out = teacher(xt)
        s_out = student(xs)
        loss = loss_mod(s_out, t_out, c["epoch"])
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    ref = golden.ref64
    assert abs(float(loss) - float(ref["loss1"])) / abs(float(ref["loss1"])) < 2e-2
    g = student.last_layer.weight_v.grad
    assert torch.isfinite(g).all()
    assert rel_err(g.cpu().numpy(), ref["grad.last_layer.weight_v"]) < 2e-2


def test_state_dict_round_trip_and_names(golden):
    D, student, teacher, loss_mod = _build(golden, "fp32")
    names = [n for n, _ in student.named_parameters()]
    assert names[-2:] == ["last_layer.weight_g", "last_layer.weight_v"]
    assert student.last_layer.weight_g.shape == (golden.cfg["out_dim"], 1)
    assert student.last_layer.weight_g.requires_grad == (not golden.cfg["norm_last_layer"])
    assert any("last_layer" in n for n in names)              # cancel_gradients_last_layer keys on this substring
    teacher.load_state_dict(student.state_dict())             # main_dino_mc.py:262
    for a, b in zip(student.state_dict().values(), teacher.state_dict().values()):
        assert torch.equal(a, b)


def test_cpu_tensors_are_rejected():
    import dinomc_b200 as D
    head = D.DINOHead(32, 64, hidden_dim=32, bottleneck_dim=16)
    with pytest.raises(RuntimeError, match="CUDA"):
        head(torch.zeros(4, 32))
    loss = D.DINOLoss(64, 2, 0.04, 0.04, 0, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        loss(torch.zeros(4, 64), torch.zeros(4, 64), 0)


def test_ddp_wrapping_single_rank(golden):
    """DDP must see ordinary leaf parameters and get their gradients through the custom Functions."""
    if golden.name != "mc_small":
        pytest.skip("one case is enough")
    import os
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1)
        created = True
    try:
        D, student, teacher, loss_mod = _build(golden, "fp32")
        c = golden.cfg
        ddp = DDP(student, device_ids=[0])
        xs = torch.from_numpy(golden.inputs["x_student"]).cuda()
        xt = torch.from_numpy(golden.inputs["x_teacher"]).cuda()
        with torch.no_grad():
            t_out = teacher(xt)
        loss = loss_mod(ddp(xs), t_out, c["epoch"])
        loss.backward()
        ref = golden.ref64
        assert abs(float(loss) - float(ref["loss1"])) / abs(float(ref["loss1"])) < 1e-5
        assert rel_err(student.last_layer.weight_v.grad.cpu().numpy(), ref["grad.last_layer.weight_v"]) < 1e-5
        assert rel_err(loss_mod.center.cpu().numpy(), ref["center1"]) < 1e-6   # all_reduce over world 1
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_use_bn_head_matches_torch_restatement(mode, tol):
    """`use_bn_in_head` (utils/vision_transformer.py:268-274).  The BatchNorm1d MLP stays torch
    modules; F.normalize and the weight-normed last layer (forward and backward) run on our kernels.  Checked against the
    same head evaluated in float64 on the CPU (train-mode batch statistics), including the running statistics."""
    import copy
    import torch.nn.functional as F
    import dinomc_b200 as D
    torch.manual_seed(11)
    head = D.DINOHead(64, 1024, use_bn=True, norm_last_layer=False, hidden_dim=128, bottleneck_dim=64).cuda()
    assert [type(m).__name__ for m in head.mlp] == ["Linear", "BatchNorm1d", "GELU", "Linear", "BatchNorm1d", "GELU", "Linear"]
    head.precision = mode
    with torch.no_grad():
        head.last_layer.weight_g.uniform_(0.5, 1.5)
    ref_mlp = copy.deepcopy(head.mlp).cpu().double()
    g64 = head.last_layer.weight_g.detach().cpu().double().requires_grad_(True)
    v64 = head.last_layer.weight_v.detach().cpu().double().requires_grad_(True)
    x = torch.randn(48, 64, generator=torch.Generator().manual_seed(12))
    up = torch.randn(48, 1024, generator=torch.Generator().manual_seed(13))
    xg = x.cuda().requires_grad_(True)
    out = head(xg)
    (out.float() * up.cuda()).sum().backward()
    x64 = x.double().requires_grad_(True)
    z = F.normalize(ref_mlp(x64), dim=-1, p=2)
    ref = z @ (v64 * (g64 / v64.norm(dim=1, keepdim=True))).t()
    (ref * up.double()).sum().backward()
    assert rel_err(out.detach().float().cpu().numpy(), ref.detach().numpy()) < tol
    assert rel_err(xg.grad.cpu().numpy(), x64.grad.numpy()) < tol
    assert rel_err(head.last_layer.weight_v.grad.cpu().numpy(), v64.grad.numpy()) < tol
    assert rel_err(head.last_layer.weight_g.grad.cpu().numpy(), g64.grad.numpy()) < tol
    wmax = max(float(q.grad.abs().max()) for q in ref_mlp.parameters())
    for (n, p), q in zip(head.mlp.named_parameters(), ref_mlp.parameters()):
        if n in ("0.bias", "3.bias"):          # a bias in front of a BatchNorm has an exactly-zero gradient: only rounding noise
            assert float(p.grad.abs().max()) < 1e-4 * wmax and float(q.grad.abs().max()) < 1e-10, n
        else:
            assert rel_err(p.grad.cpu().numpy(), q.grad.numpy()) < tol, n
    assert rel_err(head.mlp[1].running_var.cpu().numpy(), ref_mlp[1].running_var.numpy()) < 1e-5
    assert "mlp.1.running_mean" in head.state_dict()          # buffers: in checkpoints, excluded from the EMA zip (parameters only)


@pytest.mark.skipif(not __import__("os").environ.get("DMC_TEST_EXPERIMENTAL"), reason="experimental switch, not part of the default path")
def test_experimental_bf16_wgrad_storage(golden, monkeypatch):
    """functional.wgrad_bf16: dW of the last layer stored in bf16 between the wgrad GEMM and the weight-norm backward."""
    import dinomc_b200.functional as Fn
    monkeypatch.setattr(Fn, "wgrad_bf16", True)
    D, student, teacher, loss_mod = _build(golden, "bf16")
    c, ref = golden.cfg, golden.ref64
    xs = torch.from_numpy(golden.inputs["x_student"]).cuda().requires_grad_(True)
    with torch.no_grad():
        t_out = teacher(torch.from_numpy(golden.inputs["x_teacher"]).cuda())
    loss_mod(student(xs), t_out, c["epoch"]).backward()
    for name, p in student.named_parameters():
        if "grad." + name in ref:
            assert rel_err(p.grad.cpu().numpy().reshape(ref["grad." + name].shape), ref["grad." + name]) < 2e-2, name


def test_bf16_gradient_exchange_single_rank(golden):
    """GradAllReduce(compress="bf16") on a 1-rank NCCL group: the last layer's dW leaves the wgrad GEMM in bf16, is
    'averaged', and the weight-norm backward runs on the communication stream and sets weight_v.grad (and weight_g.grad
    when the gain is trainable, golden tp_small) directly; the small gradients travel through one flat bf16 buffer.
    Everything must still meet the bf16-GEMM tolerance against the reference's gradients, eagerly and from a CUDA graph."""
    if golden.name not in ("mc_small", "tp_small"):
        pytest.skip("two cases are enough")
    import os
    import torch.distributed as dist
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1)
        created = True
    D, student, teacher, loss_mod = _build(golden, "bf16")
    red = None
    try:
        c, ref, tol = golden.cfg, golden.ref64, 2e-2
        red = D.GradAllReduce(student.parameters(), compress="bf16")
        xs = torch.from_numpy(golden.inputs["x_student"]).cuda().requires_grad_(True)
        xt = torch.from_numpy(golden.inputs["x_teacher"]).cuda()
        center0 = loss_mod.center.clone()

        def step():
            for p in student.parameters():
                p.grad = None
            xs.grad = None
            with torch.no_grad():
                t_out = teacher(xt)
            loss = loss_mod(student(xs), t_out, c["epoch"])
            loss.backward()
            red.wait()
            return loss

        def check(loss):
            torch.cuda.synchronize()
            assert abs(float(loss) - float(ref["loss1"])) / abs(float(ref["loss1"])) < tol
            assert rel_err(xs.grad.cpu().numpy(), ref["grad.x"]) < tol
            n = 0
            for name, p in student.named_parameters():
                key = "grad." + name
                if key in ref:
                    assert p.grad is not None and p.grad.dtype == torch.float32 and p.grad.shape == p.shape, name
                    assert rel_err(p.grad.cpu().numpy().reshape(ref[key].shape), ref[key]) < tol, name
                    n += 1
                else:
                    assert p.grad is None, name
            assert n >= 3

        def reset_center():
            loss_mod.center.copy_(center0)                         # in place: also the buffer a captured step reads

        check(step())
        eager_v = student.last_layer.weight_v.grad.clone()
        reset_center()
        check(step())                                              # a second step claims the route again (.grad was reset)
        assert torch.equal(student.last_layer.weight_v.grad, eager_v)
        g = D.StepGraph(step, warmup=2, capture_error_mode="thread_local")
        reset_center()
        check(g.replay())
        assert torch.equal(student.last_layer.weight_v.grad, eager_v)
        reset_center()
        # a gradient already sitting in .grad (accumulation) sends the last layer through the regular autograd route
        for p in student.parameters():
            p.grad = None
        student.last_layer.weight_v.grad = torch.zeros_like(student.last_layer.weight_v)
        with torch.no_grad():
            t_out = teacher(xt)
        loss_mod(student(xs), t_out, c["epoch"]).backward()
        red.wait()
        torch.cuda.synchronize()
        assert rel_err(student.last_layer.weight_v.grad.cpu().numpy(), ref["grad.last_layer.weight_v"]) < tol
    finally:
        if red is not None:
            red.remove()
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_headline_size_properties(mode):
    """ViT-S/8 head at BASELINE cfg2 size (D=384, K=65536, B=256/4 to bound memory in fp32, 2+6 crops):
    properties that need no oracle at this size."""
    import dinomc_b200 as D
    from oracle import np_oracle as O
    torch.manual_seed(0)
    B, C, G, K, Din = (256 if mode == "bf16" else 64), 8, 2, 65536, 384
    student = D.DINOHead(Din, K).cuda()
    teacher = D.DINOHead(Din, K).cuda()
    teacher.load_state_dict(student.state_dict())
    student.precision = teacher.precision = mode
    loss_mod = D.DINOLoss(K, C, 0.04, 0.04, 0, 10, teacher_crops_number=G).cuda()
    xs = torch.randn(C * B, Din, device="cuda", requires_grad=True)
    xt = torch.randn(G * B, Din, device="cuda")
    with torch.no_grad():
        t_out = teacher(xt)
    s_out = student(xs)
    s_out.retain_grad()
    center0 = loss_mod.center.clone()
    loss = loss_mod(s_out, t_out, 0)
    loss.backward()
    # (1) rows of logits are bounded by ||zhat|| * ||w_k|| = 1 * g_k = 1
    assert s_out.float().abs().max().item() <= 1.0 + 2e-2
    # (2) loss equals the oracle's closed form evaluated on OUR logits for a few samples' worth of rows
    sub = slice(0, 4)
    idx_s = torch.cat([torch.arange(v * B, v * B + 4) for v in range(C)])
    idx_t = torch.cat([torch.arange(i * B, i * B + 4) for i in range(G)])
    ref_sub = O.dino_loss_closed(s_out.detach()[idx_s].double().cpu().numpy(), t_out[idx_t].double().cpu().numpy(),
                                 center0.double().cpu().numpy(), 0.04, C, G)
    full_ref_scale = abs(ref_sub)
    assert abs(float(loss) - ref_sub) / full_ref_scale < 5e-2          # 4 samples estimate the batch mean
    # (3) every dlogits row sums to ~0 and the whole gradient is finite
    gl = s_out.grad.float()
    assert torch.isfinite(gl).all()
    assert gl.sum(-1).abs().max().item() < 1e-2 / B
    # (4) center = 0.9*0 + 0.1*mean(teacher logits)
    ref_c = 0.1 * t_out.float().mean(0, keepdim=True)
    assert rel_err(loss_mod.center.cpu().numpy(), ref_c.double().cpu().numpy()) < (1e-5 if mode == "fp32" else 1e-3)
    # (5) frozen gain: no grad; direction grads orthogonal to v (weight-norm property: dv . v = 0 per row)
    assert student.last_layer.weight_g.grad is None
    dv, v = student.last_layer.weight_v.grad, student.last_layer.weight_v.detach()
    ortho = (dv * v).sum(-1).abs().max().item()
    assert ortho < 1e-4 * dv.abs().max().item() + 1e-12
    # (6) EMA with m=1 is the identity, with m=0 copies the student
    tp, sp = list(teacher.parameters()), list(student.parameters())
    before = [p.detach().clone() for p in tp]
    D.ema_update_(tp, sp, 1.0)
    assert all(torch.equal(a, b) for a, b in zip(before, tp))
    D.ema_update_(tp, sp, 0.0)
    assert all(torch.equal(a.detach(), b.detach()) for a, b in zip(tp, sp))


def test_teacher_overlap_side_stream(golden):
    """Teacher head on the side stream (set_teacher_overlap): same numbers, DINOLoss waits for the event."""
    if golden.name != "mc_wide":
        pytest.skip("one case is enough")
    import dinomc_b200 as D
    D.set_teacher_overlap(True)
    try:
        _, student, teacher, loss_mod = _build(golden, "fp32")
        c, ref = golden.cfg, golden.ref64
        xs = torch.from_numpy(golden.inputs["x_student"]).cuda().requires_grad_(True)
        xt = torch.from_numpy(golden.inputs["x_teacher"]).cuda()
        for _ in range(3):                                     # repeated steps exercise stream/allocator reuse
            loss_mod.center.copy_(torch.from_numpy(golden.inputs["center0"]))
            with torch.no_grad():
                t_out = teacher(xt)
            assert getattr(t_out, "_dmc_ready_event", None) is not None
            s_out = student(xs)
            loss = loss_mod(s_out, t_out, c["epoch"])
            xs.grad = None
            loss.backward()
        assert abs(float(loss.detach()) - float(ref["loss1"])) / abs(float(ref["loss1"])) < 1e-5
        assert rel_err(xs.grad.cpu().numpy(), ref["grad.x"]) < 1e-5
        assert rel_err(loss_mod.center.cpu().numpy(), ref["center1"]) < 1e-6
        D.wait_ready(t_out)
        assert rel_err(t_out.float().cpu().numpy(), ref["teacher_logits"]) < 1e-5
    finally:
        D.set_teacher_overlap(False)


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_fused_statistics_route_matches_plain_route(mode, tol):
    """The GEMM-epilogue statistics + single-pass loss/gradient route against the separate-passes route, on a shape
    where the fused route is taken (out_dim > 128), incl. a ragged out_dim and an upstream gradient != 1."""
    import dinomc_b200 as D
    Fn = D.functional
    torch.manual_seed(3)
    Din, K, B, C, G = 64, 1000, 6, 8, 2
    student = D.DINOHead(Din, K, hidden_dim=128, bottleneck_dim=64).cuda()
    teacher = D.DINOHead(Din, K, hidden_dim=128, bottleneck_dim=64).cuda()
    student.precision = teacher.precision = mode
    xs = torch.randn(C * B, Din, device="cuda")
    xt = torch.randn(G * B, Din, device="cuda")
    res = {}
    for fused in (True, False):
        Fn.fused_stats_enabled = fused
        Fn.fused_teacher_stats = fused                    # also cover the (default-off) teacher statistics fusion
        try:
            loss_mod = D.DINOLoss(K, C, 0.04, 0.07, 3, 10, teacher_crops_number=G).cuda()
            loss_mod.center.normal_(0, 0.2, generator=torch.Generator(device="cuda").manual_seed(5))
            x = xs.clone().requires_grad_(True)
            for epoch in (0, 0, 1):                       # epoch 1 changes the teacher temperature -> one fallback step
                with torch.no_grad():
                    t_out = teacher(xt)
                s_out = student(x)
                assert (getattr(s_out, "_dmc_stats", None) is not None) == fused
                loss = loss_mod(s_out, t_out, epoch)
                for p in student.parameters():
                    p.grad = None
                x.grad = None
                (loss * 3.0).backward()                   # upstream gradient 3: exercises the rescale kernel
            res[fused] = (float(loss.detach()), x.grad.clone(), student.last_layer.weight_v.grad.clone(),
                          loss_mod.center.clone())
        finally:
            Fn.fused_stats_enabled = True
            Fn.fused_teacher_stats = False
    lf, lp = res[True][0], res[False][0]
    assert abs(lf - lp) / abs(lp) < tol
    for a, b in zip(res[True][1:], res[False][1:]):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < tol


def test_aux_stream_overlap_is_bit_identical(golden):
    """functional.aux_overlap only moves independent kernels (weight-norm materialisation / backward, bias-gradient
    column sums, later layers' operand casts) to a second stream: every result must be bit-identical."""
    from dinomc_b200 import functional as Fn
    results = []
    saved = Fn.aux_overlap
    try:
        for flag in (True, False):
            Fn.aux_overlap = flag
            D, student, teacher, loss_mod = _build(golden, "bf16")
            xs = torch.from_numpy(golden.inputs["x_student"]).cuda().requires_grad_(True)
            xt = torch.from_numpy(golden.inputs["x_teacher"]).cuda()
            with torch.no_grad():
                t_out = teacher(xt)
            loss = loss_mod(student(xs), t_out, golden.cfg["epoch"])
            loss.backward()
            torch.cuda.synchronize()
            results.append([loss.detach().clone(), xs.grad.clone(), loss_mod.center.clone()] +
                           [p.grad.clone() for p in student.parameters() if p.grad is not None])
    finally:
        Fn.aux_overlap = saved
    assert len(results[0]) == len(results[1]) > 5
    for a, b in zip(*results):
        assert torch.equal(a, b)


def test_step_graph_replay_matches_eager(golden):
    """StepGraph (the whole step captured once, auxiliary streams and programmatic launches included) replays to the
    same loss / gradients / center as the eager call sequence, step after step."""
    import dinomc_b200 as D2

    def make():
        D, student, teacher, loss_mod = _build(golden, "bf16")
        xs = torch.from_numpy(golden.inputs["x_student"]).cuda().requires_grad_(True)
        xt = torch.from_numpy(golden.inputs["x_teacher"]).cuda()

        def step():
            for p in student.parameters():
                p.grad = None
            xs.grad = None
            with torch.no_grad():
                t_out = teacher(xt)
            loss = loss_mod(student(xs), t_out, golden.cfg["epoch"])
            loss.backward()
            D.ema_update_(list(teacher.parameters()), list(student.parameters()), 0.99)
            return loss
        return step, student, teacher, loss_mod, xs

    step_e, st_e, te_e, lm_e, xs_e = make()
    step_g, st_g, te_g, lm_g, xs_g = make()
    graph = D2.StepGraph(step_g, warmup=2)          # the warm-up steps preserve teacher / center; the capture pass executes nothing
    for _ in range(3):
        le = step_e()
        lg = graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(le.detach(), lg.detach())
        assert torch.equal(xs_e.grad, xs_g.grad)
        assert torch.equal(lm_e.center, lm_g.center)
    for pe, pg in zip(te_e.parameters(), te_g.parameters()):
        assert torch.equal(pe, pg)


def test_step_graph_follows_schedules(golden):
    """A captured step must not freeze host scalars the reference changes while training (main_dino_mc.py:404, :445):
    replay(momentum=...) follows the EMA momentum schedule through device-resident scalars, replay(epoch=...) re-captures
    when the teacher temperature of that epoch differs.  Checked against eager steps with the same schedule."""
    if golden.name != "tp_small":
        pytest.skip("the case with a warm-up temperature schedule")
    import dinomc_b200 as D2
    c = golden.cfg
    moms = [0.99, 0.9925, 0.995, 0.9975]
    epochs = [0, 0, 1, min(2, c["nepochs"] - 1)]

    def make():
        D, student, teacher, loss_mod = _build(golden, "bf16")
        xs = torch.from_numpy(golden.inputs["x_student"]).cuda().requires_grad_(True)
        xt = torch.from_numpy(golden.inputs["x_teacher"]).cuda()

        def step(epoch=0, m=moms[0]):
            for p in student.parameters():
                p.grad = None
            xs.grad = None
            with torch.no_grad():
                t_out = teacher(xt)
            loss = loss_mod(student(xs), t_out, epoch)
            loss.backward()
            D.ema_update_(list(teacher.parameters()), list(student.parameters()), m)
            return loss
        return step, teacher, loss_mod, xs

    step_e, te_e, lm_e, xs_e = make()
    step_g, te_g, lm_g, xs_g = make()
    assert len(set(float(lm_e.teacher_temp_schedule[e]) for e in epochs)) > 1, "the case must exercise a temperature change"
    t0 = [p.detach().clone() for p in te_g.parameters()]
    c0 = lm_g.center.detach().clone()
    graph = D2.StepGraph(step_g, warmup=2)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(t0, te_g.parameters())), "warm-up must not advance the EMA teacher"
    assert torch.equal(c0, lm_g.center), "warm-up must not advance the center"
    for m, e in zip(moms, epochs):
        le = step_e(epoch=e, m=m)
        lg = graph.replay(momentum=m, epoch=e)
        torch.cuda.synchronize()
        assert torch.equal(le.detach(), lg.detach()), (m, e)
        assert torch.equal(xs_e.grad, xs_g.grad)
        assert torch.equal(lm_e.center, lm_g.center)
        for pe, pg in zip(te_e.parameters(), te_g.parameters()):
            assert torch.equal(pe, pg)


def test_step_graph_refuses_optimizer_and_fixed_epoch(golden):
    if golden.name != "tp_small":
        pytest.skip("one case is enough")
    import dinomc_b200 as D2
    D, student, teacher, loss_mod = _build(golden, "bf16")
    xs = torch.from_numpy(golden.inputs["x_student"]).cuda().requires_grad_(True)
    xt = torch.from_numpy(golden.inputs["x_teacher"]).cuda()
    opt = D2.FusedAdamW([p for p in student.parameters() if p.requires_grad], lr=1e-3)

    def step_with_opt():
        with torch.no_grad():
            t_out = teacher(xt)
        loss_mod(student(xs), t_out, 0).backward()
        opt.step()

    with pytest.raises(RuntimeError, match="StepGraph"):
        D2.StepGraph(step_with_opt, warmup=1)

    def step_fixed():
        for p in student.parameters():
            p.grad = None
        with torch.no_grad():
            t_out = teacher(xt)
        loss = loss_mod(student(xs), t_out, 0)
        loss.backward()
        return loss

    g = D2.StepGraph(step_fixed, warmup=1)
    g.replay()
    late = golden.cfg["nepochs"] - 1
    if float(loss_mod.teacher_temp_schedule[late]) != float(loss_mod.teacher_temp_schedule[0]):
        with pytest.raises(RuntimeError, match="epoch"):
            g.replay(epoch=late)


def test_teacher_operand_shadows_match_regular_path(golden):
    """The EMA pass refreshes the frozen teacher's bf16 GEMM operands (MLP weights, W = g v/||v||); a teacher forward
    from those shadows must equal the regular path (casts + weight-norm pass) bit for bit, and any other write to the
    teacher's parameters must switch the shadows off until the next EMA pass."""
    import dinomc_b200 as D2
    D, student, teacher, loss_mod = _build(golden, "bf16")
    xt = torch.from_numpy(golden.inputs["x_teacher"]).cuda()
    tp, sp = list(teacher.parameters()), list(student.parameters())
    with torch.no_grad():
        teacher(xt)                                             # first sight of a frozen no-grad head: allocates the shadows
    assert teacher._shadow is not None and teacher._shadow["state"] is None
    D2.ema_update_(tp, sp, 0.9)
    assert teacher._fresh_shadow("bf16", xt) is not None      # frozen parameters + an input without gradient: no no_grad needed
    with torch.no_grad():
        out_shadow = teacher(xt).clone()
        D2.head.set_operand_shadows(False)
        try:
            out_regular = teacher(xt).clone()
        finally:
            D2.head.set_operand_shadows(True)
    assert torch.equal(out_shadow, out_regular)
    # the shadow of the last layer equals the weight-norm kernel's own output on the updated parameters
    w, _, scale, inv = D2.ops.weightnorm_fwd(teacher.last_layer.weight_v.detach(), teacher.last_layer.weight_g.detach().reshape(-1), "bf16")
    assert torch.equal(w, teacher._shadow["what"]) and torch.equal(scale, teacher._shadow["scale"])
    assert torch.equal(inv, teacher._shadow["inv_norm"])
    # EMA stays bit-exact with the reference's op sequence (including weight_g, which travels with the rows)
    ref_t = [p.detach().clone() for p in tp]
    D2.ema_update_(tp, sp, 0.75)
    for r, q in zip(ref_t, sp):
        r.mul_(0.75).add_((1 - 0.75) * q.detach())
    assert all(torch.equal(a, b) for a, b in zip(ref_t, tp))
    # a foreign write invalidates
    with torch.no_grad():
        teacher.last_layer.weight_v.mul_(1.5)
        assert teacher._fresh_shadow("bf16", xt) is None
        out2 = teacher(xt)
    assert torch.isfinite(out2.float()).all()
