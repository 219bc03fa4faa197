"""CPU: the restated oracles (oracle/np_oracle.py, oracle/torch_port.py) against the golden vectors
that the REAL reference modules produced (tests/golden, made by oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import np_oracle as O
from oracle import torch_port as T

TOL64 = 1e-10   # fp64 oracle vs fp64 reference run


def test_np_head_forward(golden):
    s = O.head_forward(golden.inputs["x_student"], golden.sd("student"))
    t = O.head_forward(golden.inputs["x_teacher"], golden.sd("teacher"))
    assert rel_err(s, golden.ref64["student_logits"]) < TOL64
    assert rel_err(t, golden.ref64["teacher_logits"]) < TOL64


def test_np_loss_loop_and_closed_form(golden):
    c = golden.cfg
    args = (golden.ref64["student_logits"], golden.ref64["teacher_logits"], golden.inputs["center0"],
            float(golden.inputs["temp"]), c["ncrops"], c["G"])
    assert abs(O.dino_loss_loop(*args) - float(golden.ref64["loss1"])) < 1e-11
    assert abs(O.dino_loss_closed(*args) - float(golden.ref64["loss1"])) < 1e-11
    # second evaluation uses the updated center (loss-before-center ordering, main_dino_mc.py:459-460)
    args2 = args[:2] + (golden.ref64["center1"],) + args[3:]
    assert abs(O.dino_loss_closed(*args2) - float(golden.ref64["loss2"])) < 1e-11


def test_np_loss_grad(golden):
    c = golden.cfg
    g = O.dino_loss_grad(golden.ref64["student_logits"], golden.ref64["teacher_logits"], golden.inputs["center0"],
                         float(golden.inputs["temp"]), c["ncrops"], c["G"])
    assert rel_err(g, golden.ref64["dlogits"]) < TOL64


def test_np_head_backward(golden):
    _, cache = O.head_forward(golden.inputs["x_student"], golden.sd("student"), return_cache=True)
    grads = O.head_backward(golden.ref64["dlogits"], cache)
    checked = 0
    for k, ref in golden.ref64.items():
        if k.startswith("grad."):
            assert rel_err(grads[k[5:]].reshape(ref.shape), ref) < 1e-9, k
            checked += 1
    assert checked >= 3
    if golden.cfg["norm_last_layer"]:
        assert "grad.last_layer.weight_g" not in golden.ref64     # frozen gain gets no grad in the reference
    else:
        assert "grad.last_layer.weight_g" in golden.ref64


def test_np_center(golden):
    c1 = O.update_center(golden.inputs["center0"], golden.ref64["teacher_logits"])
    assert rel_err(c1, golden.ref64["center1"]) < 1e-12
    c2 = O.update_center(c1, golden.ref64["teacher_logits"])
    assert rel_err(c2, golden.ref64["center2"]) < 1e-12


def test_np_ema_is_bit_exact(golden):
    names = [k[4:] for k in golden.ref32 if k.startswith("ema.")]
    ssd, tsd = golden.sd("student"), golden.sd("teacher")
    out = O.ema_update_fp32([tsd[n] for n in names], [ssd[n] for n in names], float(golden.inputs["ema_m"]))
    for n, o in zip(names, out):
        assert np.array_equal(o, golden.ref32["ema." + n]), n


def test_np_schedules():
    s = O.teacher_temp_schedule(0.04, 0.07, 5, 10)
    assert len(s) == 10 and s[0] == 0.04 and abs(s[4] - 0.07) < 1e-15 and s[9] == 0.07
    m = O.cosine_scheduler(0.996, 1, 10, 7)
    assert len(m) == 70 and m[0] == 0.996 and m[-1] < 1.0
    assert O.n_loss_terms(8, 2) == 14 and O.n_loss_terms(9, 3) == 24 and O.n_loss_terms(2, 2) == 2


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_torch_port_step(golden, dtype, tol):
    c = golden.cfg
    to = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)
    sp = {k: to(v) for k, v in golden.sd("student").items()}
    tp = {k: to(v) for k, v in golden.sd("teacher").items()}
    for k, v in sp.items():
        v.requires_grad_(not (k.endswith("weight_g") and c["norm_last_layer"]))
    st = T.LossState(c["out_dim"], c["ncrops"], golden.cfg["warmup_tt"], golden.cfg["tt"], c["warmup_epochs"],
                     c["nepochs"], teacher_crops_number=c["G"], dtype=dtype)
    st.center = to(golden.inputs["center0"])
    loss, grads = T.step(to(golden.inputs["x_student"]), to(golden.inputs["x_teacher"]), sp, tp, st, c["epoch"],
                         float(golden.inputs["ema_m"]))
    assert abs(float(loss) - float(golden.ref64["loss1"])) < max(tol, 1e-10) * 10
    assert rel_err(st.center.numpy(), golden.ref64["center1"]) < max(tol, 1e-6 if dtype == torch.float32 else tol)
    for k, ref in golden.ref64.items():
        if k.startswith("grad."):
            assert rel_err(grads[k[5:]].detach().numpy().reshape(ref.shape), ref) < tol * 5, k
    if dtype == torch.float32:
        for k, ref in golden.ref32.items():
            if k.startswith("ema."):
                assert np.array_equal(tp[k[4:]].detach().numpy(), ref), k


def test_torch_port_lars_and_clip_against_reference_outputs(optim_golden):
    """oracle/torch_port.lars_step / clip_gradients against outputs of the reference's own utils.LARS / utils.clip_gradients
    (tests/golden/optim_small.npz).  Same torch ops in the same order: 1e-6 leaves room for a different CPU vector width
    in torch.norm on another host (bit-identical in the build container, tests/test_dropin_reference.py)."""
    import torch
    from oracle import torch_port
    G = optim_golden
    params = [p.detach().clone() for p in G.model.parameters()]
    mus = [torch.zeros_like(p) for p in params]
    reg, noreg = G.groups(G.model)
    for it in range(G.steps):
        lr, wd = G.schedule(it)
        gr = G.grads(it)
        torch_port.lars_step([params[i] for i in reg], [gr[i] for i in reg], [mus[i] for i in reg], lr, wd)
        torch_port.lars_step([params[i] for i in noreg], [gr[i] for i in noreg], [mus[i] for i in noreg], lr, 0.0)
    for n, p, mu in zip(G.names, params, mus):
        assert rel_err(p.numpy(), G.z["lars.p." + n]) < 1e-6, n
        assert rel_err(mu.numpy(), G.z["lars.mu." + n]) < 1e-6, n
    for clip in (3.0, 0.05):
        gr = G.grads(0)
        norms = torch_port.clip_gradients(gr, clip)
        assert rel_err(np.array(norms), G.z[f"clip{clip}.norms"]) < 1e-6
        for n, g in zip(G.names, gr):
            assert rel_err(g.numpy(), G.z[f"clip{clip}.g." + n]) < 1e-6, n
    n3 = G.z["clip3.0.norms"]
    assert (G.z["clip0.05.norms"] > 0.05).all() and (n3 > 3.0).any() and (n3 < 3.0).any()     # one case clips everything, one some
