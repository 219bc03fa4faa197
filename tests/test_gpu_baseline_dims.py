"""GPU: parity of the WHOLE step at BASELINE.json's dimensions (out_dim 65536, hidden 2048, bottleneck 256).

The oracle is `oracle/torch_port.step` -- the reference's eager path restated op for op
(main_dino_mc.py:437-473 loss + center, :403-406 EMA, utils/vision_transformer.py:290-294 head), pinned to the
real reference by tests/golden/*.npz -- executed in float64.  It is plain torch code, so for these sizes it runs
on the same device as the kernels under test (fp64 cuBLAS / ATen as the checker; a few hundred ms per case).

Every case uses the default routes of the product path: fused student statistics in the last-layer GEMM epilogue
(EPI 2, bounded form), `ce_fused`, CTA-pair split-K dgrad, dual-M wgrad, auxiliary stream, teacher on its side
stream; one case replays the step from a `StepGraph`.

Tolerances (north star), rel = max|a-b| / max|b|: fp32 mode 1e-5 on the loss and every gradient, center 1e-6,
EMA bit-exact; bf16-GEMM mode 2e-2 (center 2e-3: it is a mean of bf16-stored logits).
"""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

H, BN, K = 2048, 256, 65536
TOL = {"fp32": 1e-5, "bf16": 2e-2}
CTOL = {"fp32": 1e-6, "bf16": 2e-3}


def _case(Din, B, C, G, *, norm_last_layer=True, warmup=0, epoch=0, seed=0, center_std=0.3):
    """Modules (ours) + fp64 oracle parameters with identical values, and seeded inputs."""
    import dinomc_b200 as D
    from oracle import torch_port as T
    torch.manual_seed(seed)
    student = D.DINOHead(Din, K, norm_last_layer=norm_last_layer).cuda()
    teacher = D.DINOHead(Din, K, norm_last_layer=norm_last_layer).cuda()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        if not norm_last_layer:                                  # DINO-TP style: trainable gain away from 1
            student.last_layer.weight_g.copy_(torch.rand(K, 1, generator=g) + 0.5)
        for p in student.parameters():
            if p.dim() == 1:                                     # the reference zero-inits biases; make them count
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
        teacher.load_state_dict(student.state_dict())
        for p in teacher.parameters():                           # the EMA teacher differs from the student
            p.mul_(1.0 + 0.05 * torch.randn(p.shape, generator=g).cuda().clamp_(-2, 2))
    for p in teacher.parameters():
        p.requires_grad = False
    loss_mod = D.DINOLoss(K, C, 0.02 if warmup else 0.04, 0.04, warmup, 10, teacher_crops_number=G).cuda()
    center0 = (torch.randn(1, K, generator=g) * center_std).cuda()
    loss_mod.center.copy_(center0)
    xs = torch.randn(C * B, Din, generator=g).cuda()
    xt = torch.randn(G * B, Din, generator=g).cuda()
    # oracle copies (fp64, same device) -- taken BEFORE anything runs
    sp = {k: v.detach().double().clone().requires_grad_(v.requires_grad) for k, v in student.named_parameters()}
    tp = {k: v.detach().double().clone() for k, v in teacher.named_parameters()}
    st = T.LossState(K, C, 0.02 if warmup else 0.04, 0.04, warmup, 10, teacher_crops_number=G, dtype=torch.float64)
    st.center = center0.double().clone()
    # EMA reference in fp32 with the reference's own op sequence (bit-exact target)
    ema_t32 = [v.detach().clone() for v in teacher.parameters()]
    ema_s32 = [v.detach().clone() for v in student.parameters()]
    return dict(D=D, T=T, student=student, teacher=teacher, loss_mod=loss_mod, xs=xs, xt=xt, sp=sp, tp=tp, st=st,
                ema_t32=ema_t32, ema_s32=ema_s32, epoch=epoch, B=B, C=C, G=G)


def _ours(c, mode, m, graph=False):
    D = c["D"]
    c["student"].precision = c["teacher"].precision = mode
    xs = c["xs"].clone().requires_grad_(True)
    tparams, sparams = list(c["teacher"].parameters()), list(c["student"].parameters())

    def run():
        for p in sparams:
            p.grad = None
        xs.grad = None
        with torch.no_grad():
            t_out = c["teacher"](c["xt"])
        s_out = c["student"](xs)
        loss = c["loss_mod"](s_out, t_out, c["epoch"])
        loss.backward()
        D.ema_update_(tparams, sparams, m)
        return loss

    if graph:
        # StepGraph's warm-up runs real steps: keep the state the parity check starts from
        center0 = c["loss_mod"].center.detach().clone()
        t0 = [p.detach().clone() for p in tparams]
        step = D.StepGraph(run, warmup=2)
        with torch.no_grad():
            c["loss_mod"].center.copy_(center0)
            for p, q in zip(tparams, t0):
                p.copy_(q)
        loss = step.replay()
    else:
        loss = run()
    torch.cuda.synchronize()
    return loss, xs


def _check(c, mode, loss, xs, m):
    T = c["T"]
    tol, ctol = TOL[mode], CTOL[mode]
    ref_loss, ref_g = T.step(c["xs"].double(), c["xt"].double(), c["sp"], c["tp"], c["st"], c["epoch"], m)
    assert abs(float(loss) - float(ref_loss)) / abs(float(ref_loss)) < tol, (float(loss), float(ref_loss))
    assert rel_err(xs.grad.cpu().numpy(), ref_g["x"].cpu().numpy()) < tol
    checked = 0
    for name, p in c["student"].named_parameters():
        if name in ref_g:
            assert p.grad is not None, name
            e = rel_err(p.grad.cpu().numpy(), ref_g[name].cpu().numpy())
            assert e < tol, (name, e)
            checked += 1
        else:
            assert p.grad is None, name
    assert checked >= 7
    assert rel_err(c["loss_mod"].center.cpu().numpy(), c["st"].center.cpu().numpy()) < ctol
    # EMA: the reference's three fp32 roundings, bit for bit (fp32 torch ops on the same device are IEEE-exact here)
    T.ema_update(c["ema_t32"], c["ema_s32"], m)
    for (name, p), r in zip(c["teacher"].named_parameters(), c["ema_t32"]):
        assert torch.equal(p.detach(), r), name


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("B", [32, 256])
def test_full_step_vit_small(mode, B):
    """BASELINE configs[0] (cfg1: batch 32) and configs[1] (cfg2, the headline: batch 256), D=384, 2+6 crops."""
    if mode == "fp32" and B == 256:
        B = 128                               # fp32 logits + fp64 oracle at 256 is only memory, not coverage
    c = _case(384, B, 8, 2)
    loss, xs = _ours(c, mode, 0.996)
    _check(c, mode, loss, xs, 0.996)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_full_step_graph_replay(mode):
    """Same step replayed from a captured StepGraph (what bench.py times), cfg1 dims."""
    c = _case(384, 32, 8, 2, seed=3)
    loss, xs = _ours(c, mode, 0.9995, graph=True)
    _check(c, mode, loss, xs, 0.9995)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("Din", [768, 2048])
def test_full_step_wide_features(mode, Din):
    """Swin-t (768) and ResNet-50 / WRN-50-2 (2048) feature widths (BASELINE configs[2..4]), B=8."""
    c = _case(Din, 8, 8, 2, seed=5)
    loss, xs = _ours(c, mode, 0.996)
    _check(c, mode, loss, xs, 0.996)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_full_step_dino_tp(mode):
    """DINO-TP shape: 3 teacher crops + 6 local (C=9, G=3), trainable gain, warm-up teacher temperature."""
    c = _case(384, 8, 9, 3, norm_last_layer=False, warmup=5, epoch=2, seed=7)
    loss, xs = _ours(c, mode, 0.99)
    _check(c, mode, loss, xs, 0.99)


# ---------------------------------------------------------------------------------------------------------
# components at headline size against float64
# ---------------------------------------------------------------------------------------------------------
def test_ce_fused_headline_columns():
    """ce_fused (default loss route) at K = 65536, B = 8, 2+6 crops, bf16 logits, against the fp64 pair loop."""
    import dinomc_b200 as D
    from oracle import np_oracle as O
    ops = D.ops
    B, C, G = 8, 8, 2
    g = torch.Generator().manual_seed(11)
    s = (torch.rand(C * B, K, generator=g) * 2 - 1).bfloat16().cuda()
    t = (torch.rand(G * B, K, generator=g) * 2 - 1).bfloat16().cuda()
    center = (torch.randn(K, generator=g) * 0.3).cuda()
    inv_ts, inv_tt = 10.0, 25.0
    t_stats, colsum = ops.teacher_stats_colsum(t, center, inv_tt)
    s_lse = torch.logsumexp(s.double() * inv_ts, dim=-1).float()
    loss, ds = ops.ce_fused(s, t, center, t_stats, s_lse, B, C, G, inv_ts, inv_tt)
    sn, tn, cn = s.double().cpu().numpy(), t.double().cpu().numpy(), center.double().cpu().numpy()[None]
    ref = O.dino_loss_loop(sn, tn, cn, 0.04, C, G)
    assert abs(float(loss) - ref) / abs(ref) < 1e-5
    ref_g = O.dino_loss_grad(sn, tn, cn, 0.04, C, G)
    assert rel_err(ds.float().cpu().numpy(), ref_g) < 6e-3           # ds is stored in bf16
    assert rel_err(colsum.cpu().numpy(), tn.sum(0)) < 1e-6
    # the plain route on the same inputs
    loss2, lse2 = ops.ce_fwd(s, t, center, t_stats, B, C, G, inv_ts, inv_tt)
    assert abs(float(loss2) - ref) / abs(ref) < 1e-5
    assert rel_err(lse2.cpu().numpy(), s_lse.cpu().numpy()) < 1e-6


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_last_layer_statistics_headline(mode):
    """EPI 2 (bounded statistics) of the 256 -> 65536 GEMM at M = 2048 (bf16) / 512 (fp32): the log-sum-exp merged
    from the epilogue partials equals logsumexp of exactly the stored logits, and the logits equal the fp64 product."""
    import dinomc_b200 as D
    from dinomc_b200 import functional as Fn
    ops = D.ops
    M = 2048 if mode == "bf16" else 512
    g = torch.Generator().manual_seed(13)
    z = torch.randn(M, BN, generator=g).cuda()
    v = ((torch.rand(K, BN, generator=g) * 2 - 1) / 16).cuda()
    gain = torch.ones(K).cuda()
    prepared = Fn.last_layer_weights(mode, gain, v, BN)
    zhat, zb, _ = ops.normalize_rows_fwd(z, want_bf16=(mode == "bf16"))
    zop = Fn.Operand(zb) if mode == "bf16" else Fn.prep(zhat, mode)
    if prepared["region"] is not None:
        prepared["region"].join()
    parts = ops.gemm_stats_parts(K)
    rp = torch.empty((M, parts, 2), dtype=torch.float32, device="cuda")
    stats = dict(kind="student", scale=10.0, center=None, row_partials=rp, bound=prepared["gmax"])
    logits = Fn.mm(mode, zop, prepared["wop"], M, K, BN, out_dtype=Fn.store_dtype(mode), stats=stats)
    lse = ops.lse_finalize(rp)
    ref_lse = torch.logsumexp(logits.double() * 10.0, dim=-1)
    assert rel_err(lse.cpu().numpy(), ref_lse.cpu().numpy()) < 2e-6
    w64 = v.double() * (1.0 / v.double().norm(dim=1, keepdim=True))
    z64 = z.double() / z.double().norm(dim=1, keepdim=True)
    ref = z64 @ w64.t()
    assert rel_err(logits.double().cpu().numpy(), ref.cpu().numpy()) < (1e-2 if mode == "bf16" else 2e-6)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_last_layer_backward_gemms_headline(mode):
    """dgrad 2048 x 256 x 65536 (CTA-pair split-K) and wgrad 65536 x 256 x 2048 (dual-M, MN-major operands) against
    fp64 matmuls of the very operands the kernels read."""
    from dinomc_b200 import functional as Fn
    M = 2048
    g = torch.Generator().manual_seed(17)
    d = (torch.randn(M, K, generator=g) * 1e-3).cuda()
    w = (torch.randn(K, BN, generator=g) / 16).cuda()
    zh = (torch.randn(M, BN, generator=g) / 16).cuda()
    if mode == "bf16":
        d, w, zh = d.bfloat16(), w.bfloat16(), zh.bfloat16()
        dop, wop, zop = Fn.Operand(d), Fn.Operand(w), Fn.Operand(zh)
        tol = 2e-5            # exact products of bf16 values, fp32 accumulation over 65536 / 2048 terms
    else:
        dop, wop, zop = Fn.prep(d, mode), Fn.prep(w, mode), Fn.prep(zh, mode)
        tol = 2e-5
    dz = Fn.mm(mode, dop, wop, M, BN, K, b_mn=True, out_dtype=torch.float32)
    dw = Fn.mm(mode, dop, zop, K, BN, M, a_mn=True, b_mn=True, out_dtype=torch.float32)
    torch.cuda.synchronize()
    ref_dz = d.double() @ w.double()
    ref_dw = d.double().t() @ zh.double()
    assert rel_err(dz.cpu().numpy(), ref_dz.cpu().numpy()) < tol
    assert rel_err(dw.cpu().numpy(), ref_dw.cpu().numpy()) < tol


def test_headline_loss_on_all_rows():
    """cfg2 size (B = 256, 2+6 crops, bf16): the loss the step reports equals the reference's pair loop (fp64, on the
    device) evaluated on the logits the heads produced -- every row, not a subsample."""
    import dinomc_b200 as D
    from oracle import torch_port as T
    torch.manual_seed(0)
    B, C, G, Din = 256, 8, 2, 384
    student = D.DINOHead(Din, K).cuda()
    teacher = D.DINOHead(Din, K).cuda()
    teacher.load_state_dict(student.state_dict())
    student.precision = teacher.precision = "bf16"
    loss_mod = D.DINOLoss(K, C, 0.04, 0.04, 0, 10, teacher_crops_number=G).cuda()
    loss_mod.center.normal_(0, 0.3)
    st = T.LossState(K, C, 0.04, 0.04, 0, 10, teacher_crops_number=G, dtype=torch.float64)
    st.center = loss_mod.center.detach().double().clone()
    xs = torch.randn(C * B, Din, device="cuda", requires_grad=True)
    xt = torch.randn(G * B, Din, device="cuda")
    with torch.no_grad():
        t_out = teacher(xt)
    s_out = student(xs)
    s_out.retain_grad()
    loss = loss_mod(s_out, t_out, 0)
    loss.backward()
    s64 = s_out.detach().double().requires_grad_(True)
    ref = T.loss_forward(st, s64, D.wait_ready(t_out).double(), 0)
    assert abs(float(loss) - float(ref)) / abs(float(ref)) < 1e-5
    ref.backward()
    assert rel_err(s_out.grad.float().cpu().numpy(), s64.grad.cpu().numpy()) < 6e-3      # gradient stored in bf16
    assert rel_err(loss_mod.center.cpu().numpy(), st.center.cpu().numpy()) < 1e-6
