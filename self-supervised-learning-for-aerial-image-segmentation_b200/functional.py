"""autograd.Functions that stitch the libdinomc kernels into the DINOHead / DINOLoss graph.

Precision modes (how the GEMMs run; everything around them is fp32):
  "bf16"      bf16 operands on tcgen05 (kind::f16), fp32 accumulate in TMEM; logits and their gradient are
              stored as bf16.  The analogue of the reference's default fp16 autocast (main_dino_mc.py:89,372).
  "fp32"      fp32 operands split hi/lo and run as three TF32 tcgen05 passes (3xTF32, ~fp32 accuracy);
              logits stored fp32.  Parity mode: 1e-5 against the reference.
  "fp32_simt" fp32 FFMA kernel, no tensor cores (cross-check arm).
"""
from __future__ import annotations

import os
import weakref

import torch

from . import _lib as L
from . import ops

MODES = ("bf16", "fp32", "fp32_simt")

# ---------------------------------------------------------------------------------------------------------
# Fused logit statistics.  The last-layer GEMM can emit, from its epilogue, the softmax statistics DINOLoss
# needs (student log-sum-exp partials; teacher softmax partials + column sums), which removes the loss's own
# statistics passes over the logits.  The head does not know the temperatures or the center -- they belong to
# DINOLoss.  A head learns which DINOLoss its logits go to either explicitly (`DINOHead.bind_loss(loss)`) or, for code
# that constructs heads and loss independently and never connects them (main_dino_mc.py:236-270 -- the drop-in case),
# from the default below: the most recently constructed / used DINOLoss registers itself here (weak reference).  The
# head passes the module it resolved to NormLastLayerFn as an argument and gets the statistics record back through the
# same argument: the autograd Functions themselves read no module-level state for this.  Everything is validated again
# inside DINOLoss.forward (temperature, center identity and version, row count, tensor identity); on any mismatch the
# separate passes run instead, so a wrong guess costs time, never correctness.
# ---------------------------------------------------------------------------------------------------------
_os_environ_get = os.environ.get

fused_stats_enabled = True
fused_teacher_stats = False   # teacher row statistics + column sums from the GEMM epilogue (EPI 3, lean path for bf16 logits):
                              # correct and tested, but measured slower than the dedicated one-pass teacher kernel
                              # (step 0.810 ms against 0.796 ms: the center subtraction, the running maximum and the column
                              # sums weigh on an epilogue with two warps per scheduler), so off by default; the student's
                              # log-sum-exp partials are always fused
_loss_ref = None          # weakref: the DEFAULT DINOLoss of heads without an explicit bind_loss()


def register_loss(loss_module):
    global _loss_ref
    _loss_ref = weakref.ref(loss_module)


def _current_loss():
    return _loss_ref() if _loss_ref is not None else None


# ---------------------------------------------------------------------------------------------------------
# Auxiliary stream.  The step alternates tensor-bound kernels (the MLP GEMMs: one 320-thread CTA per SM, most
# of the register file and shared memory, little HBM traffic) with HBM-bound streaming kernels that do not
# depend on them: the weight-norm materialisation W = g v/||v|| of the last layer (reads 64 MiB, needed only by
# the last GEMM) and the bias-gradient column sums of the backward pass.  Those run on a second stream, forked
# from and joined back into the caller's stream with events, so that their 256-thread blocks (32 registers, no
# shared memory) share the SMs with the GEMM CTAs.  Works eagerly and inside CUDA-graph capture.
# ---------------------------------------------------------------------------------------------------------
aux_overlap = True
defer_joins = _os_environ_get("DMC_DEFER_JOINS", "1") != "0"       # see _AuxRegion.join_at_end_of_backward
# Grid cap of the auxiliary-stream weight-norm passes (0 = none, the default).  MEASURED (profiles/r02_scheduling_ab.md): capped
# at one CTA per SM these passes do run next to the GEMMs, but at 0.6 TB/s (8 warps per SM cannot keep enough loads in
# flight): weight-norm forward 22 -> 143 us, backward 28 -> 176 us, step 0.80 -> 0.97 ms.  Kept as a switch, off.
polite_ctas = int(_os_environ_get("DMC_POLITE_CTAS", "0"))
# Teacher row statistics from the GEMM epilogue (EPI 3) + column sums by linearity (rowdot).  MEASURED: step 0.873 -> 0.971 ms
# with it (the running-maximum epilogue more than doubles the teacher's last GEMM); kept as a switch, off.
teacher_epilogue_stats = _os_environ_get("DMC_TEACHER_EPILOGUE_STATS", "0") != "0"
fuse_normalize_bwd = _os_environ_get("DMC_FUSE_NORMALIZE_BWD", "1") != "0"
gelu_dg = _os_environ_get("DMC_GELU_DG", "1") != "0"      # MLP forward saves gelu'(z) for the backward instead of z
grad_exchange_active = False     # set by GradAllReduce: the last layer's dv (64 MiB, the largest all-reduce of the step) must
                                 # then be final as early as possible, so its weight-norm backward stays in front of the dgrad
grad_exchange = None             # the active GradAllReduce (or None); with compress="bf16" the last layer hands it a bf16 dW
wgrad_bf16 = _os_environ_get("DMC_WGRAD_BF16", "1") == "1"   # (measured: step 0.797 -> 0.785 ms) in the bf16-GEMM mode the last layer's
                                 # wgrad stores dW in bf16 and the weight-norm backward reads it (-64 MB of traffic per step at
                                 # K = 65536; same kernels as the bf16 gradient exchange, gradients stay within the 2e-2 tolerance)
_aux_streams = {}


def _aux_stream(device, cur):
    key = (torch.device(device).index, cur.cuda_stream)
    st = _aux_streams.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _aux_streams[key] = st
    return st


class _AuxRegion:
    """with _AuxRegion(dev) as r: kernels launched inside run on the auxiliary stream after everything already
    queued on the caller's stream; r.join(*tensors) makes the caller's stream wait for them."""

    def __init__(self, device, fork_now=False):
        self.cur = torch.cuda.current_stream(device)
        self.aux = _aux_stream(device, self.cur)
        self.ctx = None
        self.event = None
        self.fork_event = None
        if fork_now:             # depend on what is queued on the caller's stream NOW, even if entered later
            self.fork_event = torch.cuda.Event()
            self.fork_event.record(self.cur)

    def __enter__(self):
        if self.fork_event is not None:
            self.aux.wait_event(self.fork_event)
        else:
            self.aux.wait_stream(self.cur)
        self.ctx = torch.cuda.stream(self.aux)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self.event = torch.cuda.Event()
        self.event.record(self.aux)
        self.ctx.__exit__(*exc)
        return False

    def join(self, *tensors):
        self.cur.wait_event(self.event)
        for t in tensors:
            if t is not None:
                t.record_stream(self.cur)

    def join_at_end_of_backward(self, *tensors):
        """Inside an autograd backward: join when the whole backward pass has been queued (engine callback) instead of
        now.  For results nothing later in the backward pass reads (parameter gradients): the caller's stream then does
        not stall on the auxiliary kernel in the middle of the pass, and everything is joined before `backward()` returns."""
        if not defer_joins:
            return self.join(*tensors)
        for t in tensors:                   # allocator bookkeeping now; the callback must not keep the tensors alive
            if t is not None:               # (a second reference would make AccumulateGrad CLONE the gradient instead of taking it)
                t.record_stream(self.cur)
        cur, event = self.cur, self.event
        torch.autograd.Variable._execution_engine.queue_callback(lambda: cur.wait_event(event))


def last_layer_weights(mode, g, v, dim_in, after_current=True):
    """The weight-normed last-layer operand W = g v/||v|| (utils/vision_transformer.py:279) in the representation
    `mode` needs, plus scale = g/||v||, 1/||v|| and the device scalar max|g|.  With `aux_overlap` the kernel runs on
    the auxiliary stream; the returned record carries the region to join before the first use."""
    K = v.shape[0]
    mode = resolve_mode(mode, dim_in, K)
    kind = {"bf16": "bf16", "fp32": "tf32x3", "fp32_simt": "f32"}[mode]
    region = None
    if aux_overlap:
        region = _AuxRegion(v.device)
        with region, ops.polite(polite_ctas):          # one CTA per SM: runs NEXT TO the MLP GEMMs instead of in front of them
            w, w_lo, scale, inv_vnorm = ops.weightnorm_fwd(v.detach(), g.detach().reshape(-1), kind)
    else:
        w, w_lo, scale, inv_vnorm = ops.weightnorm_fwd(v.detach(), g.detach().reshape(-1), kind)
    return dict(mode=mode, wop=Operand(w, w_lo), scale=scale, inv_vnorm=inv_vnorm, gmax=ops.last_gmax[0], region=region)


class Operand:
    """A GEMM operand in the representation a mode needs: bf16 tensor, (hi, lo) TF32 pair, or plain fp32."""
    __slots__ = ("main", "lo")

    def __init__(self, main, lo=None):
        self.main, self.lo = main, lo


def prep(t: torch.Tensor, mode: str) -> Operand:
    if mode == "bf16":
        return Operand(ops.cast_bf16(t))
    t = t if t.dtype == torch.float32 else t.float()
    t = ops._rows2d(t)
    if mode == "fp32":
        hi, lo = ops.split_tf32(t)
        return Operand(hi, lo)
    return Operand(t)


def store_dtype(mode: str) -> torch.dtype:
    return torch.bfloat16 if mode == "bf16" else torch.float32


def resolve_mode(mode: str, *dims) -> str:
    """The tensor-core kernel reads operands through TMA descriptors, which need every stored row to be a
    multiple of 16 bytes.  Shapes that cannot satisfy this run on the (still on-device, hand-written) FFMA
    kernel instead -- never on the CPU or through torch."""
    if mode not in MODES:
        raise ValueError(f"unknown precision mode {mode!r}; expected one of {MODES}")
    if mode == "fp32_simt":
        return mode
    per16 = 8 if mode == "bf16" else 4
    return mode if all(int(d) % per16 == 0 for d in dims) else "fp32_simt"


def mm(mode, a: Operand, b: Operand, M, N, K, *, a_mn=False, b_mn=False, **kw):
    """Mode-aware GEMM (mode already resolved by resolve_mode)."""
    if mode == "fp32_simt":
        return ops.gemm(a.main, b.main, M, N, K, a_mn=a_mn, b_mn=b_mn, simt=True, **kw)
    return ops.gemm(a.main, b.main, M, N, K, a_mn=a_mn, b_mn=b_mn, A_lo=a.lo, B_lo=b.lo, **kw)


class LinearFn(torch.autograd.Function):
    """One Linear (+ GELU) of DINOHead.mlp (utils/vision_transformer.py:264-277, use_bn=False).

    forward(mode, h_in, z_in, W, b, h_op, w_op, apply_gelu) -> (h_out, z_out)
      h_in   input activations (the features for the first layer);
      z_in   the pre-activation h_in = gelu(z_in) came from, or None for the first layer;
      h_op / w_op  optional ready-made GEMM operands of h_in / W (batched casts done by the caller);
      h_out  = gelu(z_out) (store dtype) when apply_gelu, else z_out itself in fp32;  z_out is saved by the GEMM epilogue.

    One autograd node per layer (rather than one for the whole MLP) so that each layer's weight gradient reaches
    its accumulation hook -- and a data-parallel all-reduce -- as soon as that layer's wgrad has been issued.

    Private gradient convention inside the chain: the gradient a layer receives for its GELU output h_out is already
    dL/dz_out, because the consumer layer's dgrad epilogue multiplies by gelu'(z_out) (one fused pass instead of an
    elementwise kernel).  The chain is only ever built by `mlp_forward`, which keeps both ends consistent."""

    @staticmethod
    def forward(ctx, mode, h_in, z_in, W, b, h_op, w_op, apply_gelu):
        rows = h_in.shape[0]
        fo, fi = W.shape
        sd = store_dtype(mode)
        if h_op is None:
            h_op = Operand(ops._rows2d(h_in.detach())) if (mode == "bf16" and h_in.dtype == torch.bfloat16) else prep(h_in.detach(), mode)
        if w_op is None:
            w_op = prep(W.detach(), mode)
        bias = None if b is None else b.detach().float().contiguous()
        if apply_gelu and not any(ctx.needs_input_grad):
            # no autograd graph (the teacher): nothing is saved for a backward pass -- plain GELU epilogue, no second output
            z_out = h_in.new_empty(0)
            h_out = mm(mode, h_op, w_op, rows, fo, fi, out_dtype=sd, bias=bias, act=L.ACT_GELU, tag=f"gemm_mlp_fwd_{fi}x{fo}")
        elif apply_gelu:
            z_out = torch.empty((rows, fo), dtype=sd, device=h_in.device)
            # with `gelu_dg` the epilogue saves gelu'(z) instead of z (same bytes): the backward epilogue becomes a multiply
            h_out = mm(mode, h_op, w_op, rows, fo, fi, out_dtype=sd, bias=bias, act=L.ACT_GELU_DG if gelu_dg else L.ACT_GELU,
                       aux=z_out, tag=f"gemm_mlp_fwd_{fi}x{fo}")
        else:
            h_out = mm(mode, h_op, w_op, rows, fo, fi, out_dtype=torch.float32, bias=bias, tag=f"gemm_mlp_fwd_{fi}x{fo}")
            z_out = h_out.new_empty(0)
        ctx.mode, ctx.dims = mode, (rows, fo, fi)
        ctx.gelu_dg = gelu_dg               # what THIS layer's z_out holds; the consumer layer reads it from its own ctx.z_in_is_dg
        ctx.h_op, ctx.w_op, ctx.z_in = h_op, w_op, (None if z_in is None else z_in.detach())
        ctx.has_bias = b is not None
        ctx.bias_param = b
        ctx.mark_non_differentiable(z_out)
        ctx.set_materialize_grads(False)       # no zero-filled [rows, fo] gradient for the non-differentiable z_out
        return h_out, z_out

    @staticmethod
    def backward(ctx, dz, _unused):
        if dz is None:
            return (None,) * 8
        mode = ctx.mode
        rows, fo, fi = ctx.dims
        sd = store_dtype(mode)
        d_full = dz.contiguous()
        dW = db = d_in = None
        ctx.bias_grad_empty = ctx.bias_param is not None and ctx.bias_param.grad is None
        with ops.backward_cap():
            region = None
            if ctx.has_bias and ctx.needs_input_grad[4]:
                # bias gradient FIRST, forked before the wgrad is queued: it depends on dz only.  Forked after the wgrad it
                # waited for that GEMM and then for SM space under the dgrad, so the LAST layer's bias gradient was the last
                # gradient of the backward pass to become final -- and held back the final gradient exchange (8 GPUs: -30 us)
                if aux_overlap:          # HBM-bound column sum next to the tensor-bound wgrad / dgrad of this layer
                    region = _AuxRegion(d_full.device)
                    with region:
                        db = ops.colsum(d_full)
                    d_full.record_stream(region.aux)
                else:
                    db = ops.colsum(d_full)
                    ops.mark_ready(db)
            d = Operand(d_full) if (mode == "bf16" and d_full.dtype == torch.bfloat16) else prep(d_full, mode)
            if ctx.needs_input_grad[3]:
                # wgrad: dW[fo,fi] = dz^T . h_in   (both operands MN-major straight from their row-major storage)
                dW = mm(mode, d, ctx.h_op, fo, fi, rows, a_mn=True, b_mn=True, out_dtype=torch.float32, tag=f"gemm_mlp_wgrad_{fi}x{fo}")
                ops.mark_ready(dW)
            if ctx.z_in is not None:
                # dgrad with gelu'(z_in) fused: what flows upstream is already dL/dz_in
                d_in = mm(mode, d, ctx.w_op, rows, fi, fo, b_mn=True, out_dtype=sd, act=L.ACT_MUL_AUX if ctx.gelu_dg else L.ACT_GELU_BWD, aux=ctx.z_in,
                          tag=f"gemm_mlp_dgrad_{fi}x{fo}")
            elif ctx.needs_input_grad[1]:
                d_in = mm(mode, d, ctx.w_op, rows, fi, fo, b_mn=True, out_dtype=torch.float32, tag=f"gemm_mlp_dgrad_{fi}x{fo}")
            if region is not None:
                if ctx.bias_grad_empty:              # only AccumulateGrad consumes db, and only to store it
                    with torch.cuda.stream(region.aux):
                        ops.mark_ready(db)
                    region.join_at_end_of_backward(db)
                else:
                    region.join(db)
                    ops.mark_ready(db)
        return None, d_in, None, dW, db, None, None, None


def mlp_forward(mode, x, wb, w_ops=None, after_first_gemm=None):
    """The Linear/GELU chain: x -> z_last (fp32).  `wb` = [W0, b0, W1, b1, ...] (the nn.Linear parameters);
    `w_ops` = ready-made bf16 operands of the weights (a teacher's EMA-refreshed shadows), or None."""
    n = len(wb) // 2
    mode = resolve_mode(mode, x.shape[1], *[d for li in range(n) for d in wb[2 * li].shape])
    pre = None
    region = None
    if mode == "bf16" and w_ops is not None:
        pre = [Operand(ops.cast_bf16(x.detach()))] + list(w_ops)
    elif mode == "bf16":
        # bf16 operand copies of this forward: features + first weight in one launch on this stream; the later
        # layers' weights (not needed until the first GEMM has run) in a second launch on the auxiliary stream
        if aux_overlap and n > 1:
            first = ops.cast_bf16_batch([x.detach(), wb[0].detach()])
            region = _AuxRegion(x.device)
            with region:
                rest = ops.cast_bf16_batch([wb[2 * li].detach() for li in range(1, n)])
            pre = [Operand(t) for t in first + rest]
        else:
            pre = [Operand(t) for t in ops.cast_bf16_batch([x.detach()] + [wb[2 * li].detach() for li in range(n)])]
    h, z = x, None
    for li in range(n):
        if li == 1 and region is not None:
            region.join(*[o.main for o in pre[2:]])
        h, z = LinearFn.apply(mode, h, z, wb[2 * li], wb[2 * li + 1], pre[0] if (pre and li == 0) else None,
                              pre[1 + li] if pre else None, li < n - 1)
        if li == 0 and after_first_gemm is not None:
            after_first_gemm()          # e.g. fork the last layer's weight-norm pass here: it then shares the machine with the
                                        # remaining (tensor-bound) GEMMs instead of delaying the first one
    return h


class NormLastLayerFn(torch.autograd.Function):
    """F.normalize -> weight_norm(Linear(bottleneck, out_dim, bias=False))
    (utils/vision_transformer.py:292-293, :279).  forward(mode, z, weight_g, weight_v[, prepared[, link]]) -> logits.

    `link`: optional dict, the channel between the calling head and the GEMM epilogue's fused statistics: on entry
    link["loss"] is the DINOLoss the logits will go to (or None: no fused statistics), on return link["stats"] holds the
    statistics record for `DINOLoss.forward` (or None)."""

    @staticmethod
    def forward(ctx, mode, z, g, v, prepared=None, link=None):
        rows, dim = z.shape
        K = v.shape[0]
        mode = resolve_mode(mode, dim, K)
        if prepared is None or prepared["mode"] != mode:
            prepared = last_layer_weights(mode, g, v, dim)
        zhat, zhat_bf16, inv_den = ops.normalize_rows_fwd(z.detach(), want_bf16=(mode == "bf16"))
        if mode == "bf16":
            zop = Operand(zhat_bf16)
        elif mode == "fp32":
            zop = prep(zhat, mode)
        else:
            zop = Operand(zhat)
        wop, scale, inv_vnorm, gmax = prepared["wop"], prepared["scale"], prepared["inv_vnorm"], prepared["gmax"]
        if prepared["region"] is not None:           # weights were materialised on the auxiliary stream: join it
            prepared["region"].join(wop.main, wop.lo, scale, inv_vnorm, gmax)
            prepared["region"] = None
        who = "student" if any(ctx.needs_input_grad) else "teacher"
        stats = None
        loss_mod = link.get("loss") if (link is not None and fused_stats_enabled and mode != "fp32_simt" and K > 128) else None
        if loss_mod is not None:
            parts = ops.gemm_stats_parts(K)
            rp = torch.empty((rows, parts, 2), dtype=torch.float32, device=z.device) if (who == "student" or fused_teacher_stats) else None
            if who == "teacher" and teacher_epilogue_stats and mode == "bf16" and not fused_teacher_stats:
                # Teacher: row softmax statistics from the GEMM epilogue (EPI 3, no column sums); the per-GPU batch column sum
                # of the logits (main_dino_mc.py:468) by linearity, sum_rows(zhat_r . W_k) = (sum_rows zhat_r) . W_k: a
                # [K, dim] x [dim] product on a side stream next to the GEMM -- the [Nt, K] logits are never read back.
                loss_mod.sync_center()
                cen = loss_mod.center
                rp = torch.empty((rows, parts, 2), dtype=torch.float32, device=z.device)
                # bounds for the fixed-shift statistics: |logit| <= the largest gain (unit rows x weight-normed rows), plus max |center|
                bounds = torch.empty(2, dtype=torch.float32, device=z.device)
                ops.absmax_into(g.detach(), bounds[0:1])
                ops.absmax_into(cen.detach(), bounds[1:2])
                stats = dict(kind="teacher", scale=loss_mod._last_inv_tt, center=cen.reshape(-1), row_partials=rp,
                             colsum_partials=None, center_ptr=cen.data_ptr(), center_version=cen._version,
                             bound=bounds[0:1], bound2=bounds[1:2])
                side = _AuxRegion(z.device)
                with side, ops.polite(polite_ctas):
                    zbar = ops.colsum(zhat_bf16)
                    side_colsum = ops.rowdot(wop.main, zbar)
                zhat_bf16.record_stream(side.aux)
                wop.main.record_stream(side.aux)
            elif who == "teacher" and not fused_teacher_stats:
                stats = None
            elif who == "student":
                stats = dict(kind="student", scale=1.0 / loss_mod.student_temp, center=None, row_partials=rp,
                             bound=gmax)
            else:
                loss_mod.sync_center()
                cen = loss_mod.center
                stats = dict(kind="teacher", scale=loss_mod._last_inv_tt, center=cen.reshape(-1), row_partials=rp,
                             colsum_partials=torch.empty(((rows + 31) // 32, K), dtype=torch.float32, device=z.device),
                             center_ptr=cen.data_ptr(), center_version=cen._version)
        logits = mm(mode, zop, wop, rows, K, dim, out_dtype=store_dtype(mode), tag="gemm_last_fwd_" + who, stats=stats)
        if stats is not None and stats["kind"] == "teacher" and stats["colsum_partials"] is None:
            t_stats, _ = ops.teacher_finalize(stats["row_partials"], None, rows, K)
            stats = dict(kind="teacher_final", scale=stats["scale"], t_stats=t_stats, colsum=side_colsum, event=side.event,
                         center_ptr=stats["center_ptr"], center_version=stats["center_version"], rows=rows)
        if link is not None:
            link["stats"] = stats
        ctx.mode = mode
        ctx.zop, ctx.wop = zop, wop
        ctx.save_for_backward(zhat, inv_den, v.detach(), scale, inv_vnorm)
        ctx.dims = (rows, dim, K)
        ctx.g_ptr = g.data_ptr()
        ctx.v_param, ctx.g_param = v, g
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        mode = ctx.mode
        zhat, inv_den, v, scale, inv_vnorm = ctx.saved_tensors
        rows, dim, K = ctx.dims
        if dlogits.dtype != store_dtype(mode):
            dlogits = dlogits.to(store_dtype(mode))
        d = prep(dlogits, mode) if mode != "bf16" else Operand(ops._rows2d(dlogits))
        dz = dg = dv = None
        region = None
        with ops.backward_cap():
            claimed = None
            if grad_exchange is not None and mode == "bf16" and ctx.needs_input_grad[3]:
                claimed = grad_exchange.claim_last_layer(v.data_ptr(), ctx.g_ptr if ctx.needs_input_grad[2] else None)
            if claimed is not None:
                # bf16 exchange: dW leaves the GEMM in bf16, is averaged over the ranks, and the weight-norm backward runs
                # on the averaged dW on the communication stream; weight_v.grad / weight_g.grad are set there directly
                dw = mm(mode, d, ctx.zop, K, dim, rows, a_mn=True, b_mn=True, out_dtype=torch.bfloat16, tag="gemm_last_wgrad",
                        out=grad_exchange.last_layer_buffer(K, dim))       # peer transport: straight into symmetric memory
                ops.mark_ready(dw)
                want_dg = ctx.needs_input_grad[2]
                grad_exchange.exchange_last_layer(dw, lambda: ops.weightnorm_bwd(dw, v, scale, inv_vnorm, want_dg=want_dg), *claimed)
            elif ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
                # wgrad first: dW[K,dim] = dlogits^T . zhat (both MN-major).  dv is the largest gradient of the step
                # (K x 256 fp32); marking it ready here lets its all-reduce overlap the dgrad and the MLP backward.
                dw = mm(mode, d, ctx.zop, K, dim, rows, a_mn=True, b_mn=True,
                        out_dtype=torch.bfloat16 if (wgrad_bf16 and mode == "bf16") else torch.float32, tag="gemm_last_wgrad")
                if aux_overlap and not grad_exchange_active:   # streaming pass on the auxiliary stream: the dgrad GEMM does not wait for it
                    region = _AuxRegion(dw.device)
                    with region, ops.polite(polite_ctas):
                        dv, dg = ops.weightnorm_bwd(dw, v, scale, inv_vnorm, want_dg=ctx.needs_input_grad[2])
                        if ctx.needs_input_grad[3]:
                            ops.mark_ready(dv)
                    dw.record_stream(region.aux)
                else:
                    dv, dg = ops.weightnorm_bwd(dw, v, scale, inv_vnorm, want_dg=ctx.needs_input_grad[2])
                    if ctx.needs_input_grad[3]:
                        ops.mark_ready(dv)
            if ctx.needs_input_grad[1]:
                # dgrad: contraction over out_dim (split-K), W read MN-major from its [K,dim] storage.  The backward of
                # F.normalize rides in the split-K reduction (one launch instead of reduce + normalize_bwd).
                per_kb = 64 if mode == "bf16" else 32
                if fuse_normalize_bwd and mode != "fp32_simt" and dim % 4 == 0 and dim <= 1024 and K >= 2 * per_kb:
                    dz = mm(mode, d, ctx.wop, rows, dim, K, b_mn=True, out_dtype=torch.float32, act=L.ACT_NORMALIZE_BWD, aux=zhat,
                            row_scale=inv_den, row_eps=1e-12, tag="gemm_last_dgrad")
                else:
                    dzhat = mm(mode, d, ctx.wop, rows, dim, K, b_mn=True, out_dtype=torch.float32, tag="gemm_last_dgrad")
                    dz = ops.normalize_rows_bwd(dzhat, zhat, inv_den)
            if region is not None:
                if ctx.v_param.grad is None and (ctx.g_param.grad is None or not ctx.needs_input_grad[2]):
                    region.join_at_end_of_backward(dv, dg)       # dv / dg are only stored by AccumulateGrad
                else:
                    region.join(dv, dg)
            if not ctx.needs_input_grad[3]:
                dv = None
        return None, dz, dg, dv, None, None


class DinoLossFn(torch.autograd.Function):
    """DINOLoss.forward (main_dino_mc.py:437-459) on the OLD center.  forward(student, teacher, center,
    inv_ts, inv_tt, B, C, G, s_pre, t_pre) -> (loss, colsum); colsum = per-GPU batch column sum of the teacher
    logits (input of update_center).

    Two routes.  Plain: teacher pass (row statistics + column sums), forward pass (student log-sum-exp + loss),
    and in backward one pass that writes the gradient.  Fused (when the last-layer GEMM epilogue already produced
    the statistics, `s_pre` / `t_pre`): no teacher pass, and a single pass computes the loss AND the gradient for
    an upstream gradient of 1 -- backward only rescales if autograd delivers something else."""

    @staticmethod
    def forward(ctx, s, t, center, inv_ts, inv_tt, B, C, G, s_pre, t_pre):
        s_d, t_d = s.detach(), t.detach()
        center = center.detach().reshape(-1)
        K = s_d.shape[1]
        if t_pre is not None and t_pre["kind"] == "teacher_final":
            t_stats, colsum = t_pre["t_stats"], t_pre["colsum"]
        elif t_pre is not None:
            t_stats, colsum = ops.teacher_finalize(t_pre["row_partials"], t_pre["colsum_partials"], t_d.shape[0], K)
        else:
            t_stats, colsum = ops.teacher_stats_colsum(t_d, center, inv_tt)
        ctx.fused = bool(s_pre is not None and ctx.needs_input_grad[0])
        ctx.used = False
        if ctx.fused:
            s_lse = ops.lse_finalize(s_pre["row_partials"])
            loss, ds = ops.ce_fused(s_d, t_d, center, t_stats, s_lse, B, C, G, inv_ts, inv_tt)
            ctx.ds = ds
        else:
            loss, s_lse = ops.ce_fwd(s_d, t_d, center, t_stats, B, C, G, inv_ts, inv_tt)
        ctx.save_for_backward(s_d, t_d, center, t_stats, s_lse)
        ctx.cfg = (B, C, G, inv_ts, inv_tt)
        ctx.mark_non_differentiable(colsum)
        ctx.set_materialize_grads(False)         # no zero-filled [K] gradient for the non-differentiable column sum
        return loss, colsum

    @staticmethod
    def backward(ctx, gloss, _gcolsum):
        if gloss is None:
            return (None,) * 10
        s, t, center, t_stats, s_lse = ctx.saved_tensors
        B, C, G, inv_ts, inv_tt = ctx.cfg
        if ctx.fused and not ctx.used:
            ctx.used = True
            ds = ops.scale_inplace_if(ctx.ds, gloss, 1.0)     # no-op kernel when the upstream gradient is 1
            ctx.ds = None
        else:
            ds = ops.ce_bwd(s, t, center, t_stats, s_lse, gloss, B, C, G, inv_ts, inv_tt)
        return ds, None, None, None, None, None, None, None, None, None
