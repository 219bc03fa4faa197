"""TEST DOUBLE for the kernels behind `dinomc_b200.ops` -- test infrastructure, never imported by the product.

The container that runs the `-m "not gpu"` suite has no GPU, so the CUDA kernels cannot execute there.  What CAN be
checked on the host is everything ABOVE the C ABI: the autograd wiring of `functional.py` (which operand goes into which
GEMM in which layout, what each epilogue is asked to do, what is saved for backward, how many gradients each Function
returns), the statistics hand-over between `DINOHead` and `DINOLoss`, and the center update.  `install(monkeypatch)`
replaces the tensor-level wrappers of `dinomc_b200.ops` that this path calls by plain torch restatements of their
documented contracts (include/dinomc.h) and switches off the stream forks; the test then runs the real modules on CPU
tensors and compares the step with `oracle/np_oracle.py`.  A wiring mistake (a transposed operand, a missing gelu', a
statistics record that goes to the wrong loss) shows up as a numerical mismatch; kernel arithmetic is NOT under test
here -- that is what the `-m gpu` parity suite is for.
"""
from __future__ import annotations

import math

import torch

from dinomc_b200 import _lib as L

calls = []          # op names in call order (the tests assert on the launch sequence)


def _gelu(z):
    return 0.5 * z * (1.0 + torch.erf(z / math.sqrt(2.0)))


def _gelu_grad(z):
    return 0.5 * (1.0 + torch.erf(z / math.sqrt(2.0))) + z * torch.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)


def gemm(A, B, M, N, K, *, a_mn=False, b_mn=False, A_lo=None, B_lo=None, out=None, out_dtype=torch.float32, col_scale=None,
         bias=None, alpha=1.0, alpha_dev=None, act=L.ACT_NONE, aux=None, simt=False, split_k=0, tag="gemm", stats=None,
         row_scale=None, row_eps=0.0):
    """D[M,N] = epilogue(sum_k A(m,k) B(n,k)); A stored [M,K] or (a_mn) [K,M]; B stored [N,K] or (b_mn) [K,N]."""
    calls.append(tag)
    assert tuple(A.shape) == ((K, M) if a_mn else (M, K)), (tag, tuple(A.shape), M, K, a_mn)
    assert tuple(B.shape) == ((K, N) if b_mn else (N, K)), (tag, tuple(B.shape), N, K, b_mn)
    assert A.dtype == B.dtype and (A_lo is None) == (B_lo is None)
    a = A.double() + (A_lo.double() if A_lo is not None else 0.0)
    b = B.double() + (B_lo.double() if B_lo is not None else 0.0)
    a = a.t() if a_mn else a
    b = b if b_mn else b.t()
    acc = float(alpha) * (a @ b)
    if alpha_dev is not None:
        acc = acc * alpha_dev.double()
    if col_scale is not None:
        acc = acc * col_scale.double()[None, :]
    if bias is not None:
        acc = acc + bias.double()[None, :]
    if act == L.ACT_GELU:
        if aux is not None:
            aux.copy_(acc.to(aux.dtype))
        acc = _gelu(acc)
    elif act == L.ACT_GELU_DG:
        aux.copy_(_gelu_grad(acc).to(aux.dtype))
        acc = _gelu(acc)
    elif act == L.ACT_GELU_BWD:
        acc = acc * _gelu_grad(aux.double())
    elif act == L.ACT_MUL_AUX:
        acc = acc * aux.double()
    elif act == L.ACT_NORMALIZE_BWD:            # backward of F.normalize: aux = zhat, row_scale = 1 / max(||z||, eps)
        zh = aux.double()
        acc = (acc - zh * (zh * acc).sum(dim=1, keepdim=True)) * row_scale.double()[:, None]
    else:
        assert act == L.ACT_NONE, act
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype)
    out.copy_(acc.to(out.dtype))
    if stats is not None:                       # softmax partials of the STORED values, one (max, sum) pair per column part
        rp = stats["row_partials"]
        parts = rp.shape[1]
        width = (N + parts - 1) // parts
        y = out.double() * float(stats["scale"])
        if stats.get("center") is not None:
            y = y - stats["center"].double()[None, :] * float(stats["scale"])
        for p in range(parts):
            blk = y[:, p * width:(p + 1) * width]
            if blk.shape[1] == 0:
                rp[:, p, 0], rp[:, p, 1] = -1e30, 0.0
                continue
            mx = blk.max(dim=1).values
            rp[:, p, 0] = mx.float()
            rp[:, p, 1] = torch.exp(blk - mx[:, None]).sum(dim=1).float()
    return out


def lse_finalize(row_partials):
    calls.append("lse_finalize")
    mx, sm = row_partials[:, :, 0].double(), row_partials[:, :, 1].double()
    top = mx.max(dim=1).values
    return (top + torch.log((sm * torch.exp(mx - top[:, None])).sum(dim=1))).float()


def cast_bf16(x):
    calls.append("cast_bf16")
    return x if x.dtype == torch.bfloat16 else x.float().to(torch.bfloat16)


def cast_bf16_batch(tensors):
    calls.append("cast_bf16")
    return [t if t.dtype == torch.bfloat16 else t.float().to(torch.bfloat16) for t in tensors]


def split_tf32(x):
    calls.append("split_tf32")
    hi = x.float().view(torch.int32).bitwise_and(~0x1FFF).view(torch.float32)        # 10 mantissa bits
    return hi, (x.float() - hi)


def colsum(X):
    calls.append("colsum")
    return X.double().sum(dim=0).float()


def normalize_rows_fwd(z, eps=1e-12, want_bf16=False):
    calls.append("normalize_fwd")
    inv = 1.0 / z.double().norm(dim=1).clamp_min(eps)
    zh = (z.double() * inv[:, None]).float()
    return zh, (zh.to(torch.bfloat16) if want_bf16 else None), inv.float()


def normalize_rows_bwd(dzhat, zhat, inv_den, eps=1e-12, out_dtype=torch.float32):
    calls.append("normalize_bwd")
    zh, d = zhat.double(), dzhat.double()
    return ((d - zh * (zh * d).sum(dim=1, keepdim=True)) * inv_den.double()[:, None]).to(out_dtype)


last_gmax = [None]


def weightnorm_fwd(v, g, mode):
    calls.append("weightnorm_fwd")
    inv = 1.0 / v.double().norm(dim=1)
    scale = g.double().reshape(-1) * inv
    w = (v.double() * scale[:, None]).float()
    last_gmax[0] = g.abs().max().float().reshape(())
    if mode == "bf16":
        return w.to(torch.bfloat16), None, scale.float(), inv.float()
    if mode == "tf32x3":
        hi, lo = split_tf32(w)
        calls.pop()
        return hi, lo, scale.float(), inv.float()
    return w, None, scale.float(), inv.float()


def weightnorm_bwd(dw, v, scale, inv_vnorm, want_dg):
    """W = g v/||v||: dg = dW . v/||v||;  dv = (g/||v||) (dW - (dW . vhat) vhat)."""
    calls.append("weightnorm_bwd")
    vhat = v.double() * inv_vnorm.double()[:, None]
    dot = (dw.double() * vhat).sum(dim=1, keepdim=True)
    dv = (scale.double()[:, None] * (dw.double() - dot * vhat)).float()
    return dv, (dot.float() if want_dg else None)


def teacher_stats_colsum(t, center, inv_temp, bounds=None):
    calls.append("teacher_stats_colsum")
    y = (t.double() - center.double().reshape(1, -1)) * float(inv_temp)
    mx = y.max(dim=1).values
    inv_sum = 1.0 / torch.exp(y - mx[:, None]).sum(dim=1)
    return torch.stack([mx, inv_sum], dim=1).float(), t.double().sum(dim=0).float()


def _teacher_probs(t, center, t_stats, inv_tt):
    y = (t.double() - center.double().reshape(1, -1)) * float(inv_tt)
    return torch.exp(y - t_stats[:, 0].double()[:, None]) * t_stats[:, 1].double()[:, None]


def _loss_and_grad(s, t, center, t_stats, s_lse, B, C, G, inv_ts, inv_tt, want_grad):
    """main_dino_mc.py:437-459 written out pair by pair (fp64)."""
    q = _teacher_probs(t, center, t_stats, inv_tt).reshape(G, B, -1)
    x = (s.double() * float(inv_ts)).reshape(C, B, -1)
    logp = x - s_lse.double().reshape(C, B, 1)
    total, terms = 0.0, 0
    grad = torch.zeros_like(x)
    for iq in range(G):
        for v in range(C):
            if v == iq:
                continue
            total = total + (-(q[iq] * logp[v]).sum(dim=1)).mean()
            terms += 1
            if want_grad:
                grad[v] += (torch.exp(logp[v]) * q[iq].sum(dim=1, keepdim=True) - q[iq]) / B
    loss = (total / terms).float().reshape(())
    return loss, (grad * float(inv_ts) / terms).reshape(C * B, -1)


def ce_fused(s, t, center, t_stats, s_lse, B, C, G, inv_ts, inv_tt):
    calls.append("ce_fused")
    loss, ds = _loss_and_grad(s, t, center, t_stats, s_lse, B, C, G, inv_ts, inv_tt, True)
    return loss, ds.to(s.dtype)


def ce_fwd(s, t, center, t_stats, B, C, G, inv_ts, inv_tt):
    calls.append("ce_fwd")
    s_lse = torch.logsumexp(s.double() * float(inv_ts), dim=1).float()
    loss, _ = _loss_and_grad(s, t, center, t_stats, s_lse, B, C, G, inv_ts, inv_tt, False)
    return loss, s_lse


def ce_bwd(s, t, center, t_stats, s_lse, grad_out, B, C, G, inv_ts, inv_tt):
    calls.append("ce_bwd")
    _, ds = _loss_and_grad(s, t, center, t_stats, s_lse, B, C, G, inv_ts, inv_tt, True)
    return (ds * grad_out.double()).to(s.dtype)


def scale_inplace_if(x, scale_dev, expected=1.0):
    calls.append("scale_if")
    if float(scale_dev) != expected:
        x.mul_(float(scale_dev))
    return x


def center_update(center, colsum_, count, momentum):
    calls.append("center_update")
    batch_center = (colsum_.double() / float(count)).float().reshape(center.shape)
    return center * momentum + batch_center * (1 - momentum)


class ClipPlan:
    """utils/utils.py:145-154 per parameter: g *= clip / (||g|| + 1e-6) when that is below 1; returns the pre-clip norms."""

    def __init__(self, grads):
        self.grads = list(grads)

    def run(self, clip):
        calls.append("clip_grads")
        norms = torch.stack([g.double().norm() for g in self.grads])
        for g, n in zip(self.grads, norms):
            coef = float(clip) / (float(n) + 1e-6)
            if coef < 1:
                g.mul_(coef)
        return norms.float()


class EmaPlan:
    """main_dino_mc.py:403-406 with its three fp32 roundings (the shadow outputs of the real plan are not modelled)."""

    def __init__(self, teacher_params, student_params, shadows=None, wn=None):
        self.pairs = list(zip(teacher_params, student_params))

    def run(self, m):
        calls.append("ema")
        m = 1.0 if ops_preserve_state() else float(m)
        for pk, pq in self.pairs:
            pk.mul_(m).add_((1 - m) * pq)


def ops_preserve_state():
    import dinomc_b200 as D
    return D.ops.preserve_state


_REPLACED = ("ClipPlan", "EmaPlan", "gemm", "lse_finalize", "cast_bf16", "cast_bf16_batch", "split_tf32", "colsum", "normalize_rows_fwd", "normalize_rows_bwd",
             "weightnorm_fwd", "weightnorm_bwd", "teacher_stats_colsum", "ce_fused", "ce_fwd", "ce_bwd", "scale_inplace_if",
             "center_update", "last_gmax")


def install(monkeypatch):
    """Patch dinomc_b200 for a CPU run of its host logic (undone by pytest's monkeypatch at the end of the test)."""
    import dinomc_b200 as D
    from dinomc_b200 import head as H
    import sys
    me = sys.modules[__name__]
    for name in _REPLACED:
        monkeypatch.setattr(D.ops, name, getattr(me, name))
    monkeypatch.setattr(D.functional, "aux_overlap", False)              # no stream forks: there are no streams
    monkeypatch.setattr(H, "_teacher_overlap", False)
    monkeypatch.setattr(H, "_early_teacher_stats", False)
    monkeypatch.setattr(H, "_operand_shadows", False)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)   # the modules refuse CPU tensors
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self, raising=False)        # the drop-in loop moves its crops
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    from dinomc_b200 import ema, optim
    monkeypatch.setattr(ema, "_plans", {})                # plans are cached by data_ptr: never reuse one across tests
    monkeypatch.setattr(optim, "_plans", {})
    calls.clear()


# ----------------------------------------------------------------------------------------------------------------
# Data-parallel host logic (GradAllReduce) on CPU / gloo: stand-ins for CUDA streams and events (program order is the only
# order there is on the host), the bf16 narrow / widen launches, and the symmetric-memory exchange buffer.
# ----------------------------------------------------------------------------------------------------------------
class Patcher:
    """monkeypatch.setattr for processes that pytest did not start (spawned ranks): no undo needed, the process exits."""

    def setattr(self, obj, name, value, raising=True):
        setattr(obj, name, value)


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_event(self, ev):
        pass

    def wait_stream(self, st):
        pass

    def synchronize(self):
        pass


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass

    def wait(self, stream=None):
        pass


class _NullCtx:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def narrow_bf16_into(srcs, dsts):
    calls.append("narrow_bf16")
    for a, b in zip(srcs, dsts):
        assert a.dtype == torch.float32 and b.dtype == torch.bfloat16 and a.numel() == b.numel()
        b.copy_(a.to(torch.bfloat16))


def widen_bf16_batch(srcs, dsts):
    calls.append("widen_bf16")
    for a, b in zip(srcs, dsts):
        assert a.dtype == torch.bfloat16 and b.dtype == torch.float32 and a.numel() == b.numel()
        b.copy_(a.float())


class SymmetricBuffer:
    """xrank.SymmetricBuffer's contract over gloo: a flat buffer per rank; allreduce_ leaves scale * (sum over ranks) in it on
    every rank (bf16 buffers: summed in fp32, rounded once) and optionally widens element ranges into fp32 tensors."""
    instances = 0

    def __init__(self, numel, dtype, group=None, ctas=148):
        import torch.distributed as dist
        SymmetricBuffer.instances += 1
        per16 = 16 // (2 if dtype == torch.bfloat16 else 4)
        self.numel = (int(numel) + per16 - 1) // per16 * per16
        self.tensor = torch.zeros(self.numel, dtype=dtype)
        self.group, self.ctas, self.multicast = group, int(ctas), False
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def allreduce_(self, scale=1.0, widen_to=None, widen_offsets=None):
        import torch.distributed as dist
        calls.append("xrank_allreduce")
        acc = self.tensor.float()
        _real_all_reduce(acc, group=self.group)
        self.tensor.copy_((acc * float(scale)).to(self.tensor.dtype))
        for t, off in zip(widen_to or [], widen_offsets or []):
            assert t.dtype == torch.float32 and t.is_contiguous() and off % 8 == 0
            t.copy_(self.tensor[off:off + t.numel()].float())
        return self.tensor


_real_all_reduce = None


def install_data_parallel(patch):
    """On top of install(): what GradAllReduce touches.  gloo has no AVG: the stand-in all_reduce sums and divides."""
    global _real_all_reduce
    import contextlib
    import sys
    import torch.distributed as dist
    import dinomc_b200 as D
    from dinomc_b200 import xrank
    me = sys.modules[__name__]
    patch.setattr(torch.cuda, "Stream", _Stream)
    patch.setattr(torch.cuda, "Event", _Event)
    patch.setattr(torch.cuda, "stream", _NullCtx)
    patch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    patch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)
    patch.setattr(D.ops, "narrow_bf16_into", me.narrow_bf16_into)
    patch.setattr(D.ops, "widen_bf16_batch", me.widen_bf16_batch)
    patch.setattr(xrank, "SymmetricBuffer", me.SymmetricBuffer)
    if _real_all_reduce is None:
        _real_all_reduce = dist.all_reduce

    def all_reduce(t, op=dist.ReduceOp.SUM, group=None, async_op=False):
        if op == dist.ReduceOp.AVG:
            if t.dtype == torch.bfloat16:           # NCCL averages bf16 buffers natively; gloo: widen, sum, divide, round once
                acc = t.float()
                _real_all_reduce(acc, group=group)
                t.copy_((acc / dist.get_world_size(group)).to(torch.bfloat16))
            else:
                _real_all_reduce(t, group=group)
                t.div_(dist.get_world_size(group))
            return None
        return _real_all_reduce(t, op=op, group=group, async_op=async_op)

    patch.setattr(dist, "all_reduce", all_reduce)

    @contextlib.contextmanager
    def no_coalescing(*a, **k):                     # gloo cannot coalesce; one all-reduce per tensor is the same exchange
        yield

    patch.setattr(dist, "_coalescing_manager", no_coalescing)
