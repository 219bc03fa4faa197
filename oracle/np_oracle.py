"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- float64 numpy oracle.

A from-the-formulas restatement of the DINO-MC per-step hot path, written in numpy so it
travels to the GPU box (the reference itself cannot).  The forward follows the reference
line by line; the backward is derived by hand (no autograd) so that it is an independent
check of both the reference's autograd and of the CUDA kernels.

Reference lines followed (all relative to /root/reference):
  DINOHead.forward ............ utils/vision_transformer.py:290-294
  DINOHead.__init__ ........... utils/vision_transformer.py:261-282  (layer layout / names)
  weight_norm ................. utils/vision_transformer.py:279 (torch: w = g * v / ||v||_row)
  DINOLoss.__init__ ........... main_dino_mc.py:420-435  (teacher temperature schedule)
  DINOLoss.forward ............ main_dino_mc.py:437-461
  DINOLoss.update_center ...... main_dino_mc.py:463-473
  EMA loop .................... main_dino_mc.py:403-406
  cosine_scheduler ............ utils/utils.py:200-213

Parity pin: checked against tests/golden/*.npz, which were produced by running the real
reference modules in the build container (oracle/gen_golden.py).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import erf

F64 = np.float64


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def gelu(x):
    """nn.GELU() default = exact erf form (utils/vision_transformer.py:270,275)."""
    return 0.5 * x * (1.0 + erf(x / math.sqrt(2.0)))


def gelu_grad(x):
    return 0.5 * (1.0 + erf(x / math.sqrt(2.0))) + x * np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def _mlp_linear_keys(sd):
    """Ordered (weight_key, bias_key) of the MLP Linears, following the Sequential indices
    of utils/vision_transformer.py:264-277 (nlayers==1 -> a bare nn.Linear named 'mlp')."""
    if "mlp.weight" in sd:
        return [("mlp.weight", "mlp.bias")]
    idx = sorted(int(k.split(".")[1]) for k in sd if k.startswith("mlp.") and k.endswith(".weight")
                 and np.asarray(sd[k]).ndim == 2)
    return [(f"mlp.{i}.weight", f"mlp.{i}.bias") for i in idx]


# --------------------------------------------------------------------------------------
# DINOHead  (utils/vision_transformer.py:260-294), use_bn=False
# --------------------------------------------------------------------------------------
def head_forward(x, sd, return_cache=False, dtype=F64):
    """x [N,in_dim]; sd: state_dict-like {name: array} with the reference's key names.

    mlp (Linear/GELU chain) -> F.normalize(dim=-1, p=2, eps=1e-12) -> weight-normed Linear.
    """
    x = np.asarray(x, dtype)
    sd = {k: np.asarray(v, dtype) for k, v in sd.items()}
    keys = _mlp_linear_keys(sd)
    acts = [x]          # inputs of every Linear
    pre = []            # pre-activations of every Linear
    h = x
    for li, (wk, bk) in enumerate(keys):
        z = h @ sd[wk].T + sd[bk]
        pre.append(z)
        h = gelu(z) if li < len(keys) - 1 else z
        if li < len(keys) - 1:
            acts.append(h)
    z_b = h                                                   # bottleneck [N, Bn]
    nrm = np.sqrt((z_b * z_b).sum(-1, keepdims=True))
    den = np.maximum(nrm, 1e-12)                               # F.normalize eps (:292)
    zhat = z_b / den
    v = sd["last_layer.weight_v"]
    g = sd["last_layer.weight_g"]
    vnorm = np.sqrt((v * v).sum(-1, keepdims=True))           # norm over dim 1 for every out row
    w = v * (g / vnorm)                                        # torch _weight_norm(v, g, dim=0)
    logits = zhat @ w.T
    if not return_cache:
        return logits
    cache = dict(sd=sd, keys=keys, acts=acts, pre=pre, z_b=z_b, nrm=nrm, den=den, zhat=zhat,
                 v=v, g=g, vnorm=vnorm, w=w)
    return logits, cache


def head_backward(dlogits, cache):
    """Hand-derived backward of head_forward.  Returns {param_name: grad, 'x': grad}."""
    dlogits = np.asarray(dlogits, cache["zhat"].dtype)
    sd, keys = cache["sd"], cache["keys"]
    grads = {}
    zhat, w, v, g, vnorm = cache["zhat"], cache["w"], cache["v"], cache["g"], cache["vnorm"]
    dzhat = dlogits @ w                                        # [N,Bn]
    dw = dlogits.T @ zhat                                      # [K,Bn]
    vhat = v / vnorm
    dot = (dw * vhat).sum(-1, keepdims=True)                   # dW . v_hat  (per out row)
    grads["last_layer.weight_g"] = dot                         # d/dg (g * vhat) . dw
    grads["last_layer.weight_v"] = (g / vnorm) * (dw - dot * vhat)
    # normalize backward: zhat = z / max(||z||, eps)
    den, nrm, z_b = cache["den"], cache["nrm"], cache["z_b"]
    active = (nrm >= 1e-12)                                    # clamp inactive -> den depends on z
    proj = (dzhat * zhat).sum(-1, keepdims=True)
    dz = np.where(active, (dzhat - proj * zhat) / den, dzhat / den)
    # MLP backward
    dh = dz
    for li in range(len(keys) - 1, -1, -1):
        wk, bk = keys[li]
        dzl = dh if li == len(keys) - 1 else dh * gelu_grad(cache["pre"][li])
        grads[wk] = dzl.T @ cache["acts"][li]
        grads[bk] = dzl.sum(0)
        dh = dzl @ sd[wk]
    grads["x"] = dh
    return grads


# --------------------------------------------------------------------------------------
# DINOLoss  (main_dino_mc.py:419-473)
# --------------------------------------------------------------------------------------
def teacher_temp_schedule(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs, nepochs):
    """main_dino_mc.py:431-435."""
    return np.concatenate((
        np.linspace(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs),
        np.ones(nepochs - warmup_teacher_temp_epochs) * teacher_temp,
    ))


def _log_softmax(x):
    m = x.max(-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))


def _softmax(x):
    m = x.max(-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(-1, keepdims=True)


def n_loss_terms(ncrops, teacher_crops):
    return teacher_crops * ncrops - min(teacher_crops, ncrops)


def dino_loss_loop(student_output, teacher_output, center, temp, ncrops, teacher_crops=2,
                   student_temp=0.1, dtype=F64):
    """Literal double loop of main_dino_mc.py:441-459 (uses the OLD center)."""
    s = np.asarray(student_output, dtype) / student_temp
    t = np.asarray(teacher_output, dtype)
    c = np.asarray(center, dtype).reshape(1, -1)
    s_chunks = np.split(s, ncrops, axis=0)
    q_chunks = np.split(_softmax((t - c) / temp), teacher_crops, axis=0)
    total, n = 0.0, 0
    for iq, q in enumerate(q_chunks):
        for v in range(len(s_chunks)):
            if v == iq:
                continue
            loss = (-q * _log_softmax(s_chunks[v])).sum(-1)
            total += loss.mean()
            n += 1
    return total / n


def dino_loss_closed(student_output, teacher_output, center, temp, ncrops, teacher_crops=2,
                     student_temp=0.1, dtype=F64):
    """Single-pass closed form: every logit is used exactly once.

    L = 1/(n B) sum_b sum_i [ sum_{v!=i} lse_v[b] - sum_k q_i[b,k] (S[b,k] - 1[i<C] x_i[b,k]) ]
    """
    C, G = ncrops, teacher_crops
    s = np.asarray(student_output, dtype) / student_temp
    t = np.asarray(teacher_output, dtype)
    c = np.asarray(center, dtype).reshape(1, -1)
    B = s.shape[0] // C
    x = s.reshape(C, B, -1)
    q = _softmax((t - c) / temp).reshape(G, B, -1)
    m = x.max(-1, keepdims=True)
    lse = (m + np.log(np.exp(x - m).sum(-1, keepdims=True)))[..., 0]          # [C,B]
    S = x.sum(0)                                                               # [B,K]
    total = 0.0
    for i in range(G):
        lse_sum = lse.sum(0) - (lse[i] if i < C else 0.0)
        cross = (q[i] * (S - (x[i] if i < C else 0.0))).sum(-1)
        total += (lse_sum - cross).sum()
    return total / (n_loss_terms(C, G) * B)


def dino_loss_grad(student_output, teacher_output, center, temp, ncrops, teacher_crops=2,
                   student_temp=0.1, dtype=F64):
    """dL/d student_output (hand-derived).

    dL/ds_v[b,k] = (n_v p_v[b,k] - sum_{i != v} q_i[b,k]) / (n B tau_s),  n_v = #{i<G, i!=v}.
    """
    C, G = ncrops, teacher_crops
    s = np.asarray(student_output, dtype) / student_temp
    t = np.asarray(teacher_output, dtype)
    c = np.asarray(center, dtype).reshape(1, -1)
    B = s.shape[0] // C
    x = s.reshape(C, B, -1)
    q = _softmax((t - c) / temp).reshape(G, B, -1)
    p = _softmax(x)
    Q = q.sum(0)
    n = n_loss_terms(C, G)
    grad = np.empty_like(x)
    for v in range(C):
        n_v = G - (1 if v < G else 0)
        qsum = Q - (q[v] if v < G else 0.0)
        grad[v] = (n_v * p[v] - qsum) / (n * B * student_temp)
    return grad.reshape(C * B, -1)


def update_center(center, teacher_output, center_momentum=0.9, world_size=1, all_rank_outputs=None,
                  dtype=F64):
    """main_dino_mc.py:463-473.  `all_rank_outputs`: list of every rank's teacher_output
    (emulates dist.all_reduce(SUM) at :469); default = this rank only."""
    outs = [teacher_output] if all_rank_outputs is None else all_rank_outputs
    bc = sum(np.asarray(o, dtype).sum(0, keepdims=True) for o in outs)
    bc = bc / (len(teacher_output) * world_size)
    return np.asarray(center, dtype).reshape(1, -1) * center_momentum + bc * (1 - center_momentum)


# --------------------------------------------------------------------------------------
# EMA teacher update (main_dino_mc.py:403-406) with the reference's exact fp32 roundings
# --------------------------------------------------------------------------------------
def ema_update_fp32(teacher_params, student_params, m):
    """param_k.data.mul_(m).add_((1 - m) * param_q):   three fp32 roundings, `m` is a python/
    numpy float64 (momentum_schedule[it]); both scalars are cast to fp32 by torch before use."""
    m32 = np.float32(m)
    c32 = np.float32(1.0 - float(m))
    out = []
    for pk, pq in zip(teacher_params, student_params):
        pk = np.asarray(pk, np.float32)
        pq = np.asarray(pq, np.float32)
        out.append((pk * m32).astype(np.float32) + (c32 * pq).astype(np.float32))
    return out


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0):
    """utils/utils.py:200-213 (used for the EMA momentum schedule, main_dino_mc.py:305)."""
    warmup_schedule = np.array([])
    warmup_iters = warmup_epochs * niter_per_ep
    if warmup_epochs > 0:
        warmup_schedule = np.linspace(start_warmup_value, base_value, warmup_iters)
    iters = np.arange(epochs * niter_per_ep - warmup_iters)
    schedule = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * iters / len(iters)))
    schedule = np.concatenate((warmup_schedule, schedule))
    assert len(schedule) == epochs * niter_per_ep
    return schedule


# --------------------------------------------------------------------------------------
# one whole step, for smoke()/tests
# --------------------------------------------------------------------------------------
def full_step(x_student, x_teacher, student_sd, teacher_sd, center, temp, ncrops, teacher_crops,
              student_temp=0.1, center_momentum=0.9, ema_m=0.996):
    """head(student) -> head(teacher) -> loss -> grads -> new center -> EMA'd teacher head."""
    s_logits, cache = head_forward(x_student, student_sd, return_cache=True)
    t_logits = head_forward(x_teacher, teacher_sd)
    loss = dino_loss_closed(s_logits, t_logits, center, temp, ncrops, teacher_crops, student_temp)
    dlog = dino_loss_grad(s_logits, t_logits, center, temp, ncrops, teacher_crops, student_temp)
    grads = head_backward(dlog, cache)
    new_center = update_center(center, t_logits, center_momentum)
    names = list(student_sd.keys())
    new_teacher = dict(zip(names, ema_update_fp32([teacher_sd[k] for k in names],
                                                  [student_sd[k] for k in names], ema_m)))
    return dict(student_logits=s_logits, teacher_logits=t_logits, loss=loss, dlogits=dlog, grads=grads,
                center=new_center, teacher=new_teacher)
