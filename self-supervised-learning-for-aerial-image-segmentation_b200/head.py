"""DINOHead -- drop-in for the reference's `utils.vision_transformer.DINOHead`
(utils/vision_transformer.py:260-294): same constructor signature, same sub-module / parameter names,
registration order, shapes, init and state_dict keys, so `main_dino_mc.py:236-246` can construct it
unchanged and checkpoints move both ways.  The forward/backward run on the libdinomc kernels.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn

from . import functional as Fn

_default_precision = os.environ.get("DINOMC_PRECISION", "auto")


def set_default_precision(mode: str):
    """'auto' (bf16 under torch autocast, fp32 otherwise), 'bf16', 'fp32' or 'fp32_simt'."""
    global _default_precision
    if mode != "auto" and mode not in Fn.MODES:
        raise ValueError(f"unknown precision {mode!r}")
    _default_precision = mode


def get_default_precision() -> str:
    return _default_precision


# ---------------------------------------------------------------------------------------------
# optional: run no-grad (teacher) head forwards on a side stream so they overlap the student's head
# ---------------------------------------------------------------------------------------------
_teacher_overlap = False
_side_streams = {}
_operand_shadows = os.environ.get("DMC_TEACHER_SHADOWS", "1") != "0"


def set_operand_shadows(enabled: bool):
    """When enabled (default), a DINOHead that runs under torch.no_grad() with all parameters frozen (the EMA teacher,
    main_dino_mc.py:264-265) in the bf16-GEMM mode keeps bf16 copies of its GEMM operands -- the MLP weights and the
    weight-normed W = g v/||v|| -- which `ema_update_` refreshes in the same pass that updates the parameters.  The
    teacher forward then skips its cast launches and its weight-norm pass.  The copies are used only while every
    parameter is exactly what that EMA pass wrote (`_version` and storage checks); otherwise the regular path runs."""
    global _operand_shadows
    _operand_shadows = bool(enabled)


_wn_after_first_gemm = os.environ.get("DMC_WN_AFTER_FIRST_GEMM", "1") != "0"       # env: timing experiments
# fixed-shift teacher pass (dmc_teacher_stats_colsum_bounded): correct and tested, but MEASURED no faster than the general pass
# (the pass is latency- not instruction-bound at 512 rows) and its two absmax launches add ~10 us: off by default
_bounded_teacher_stats = os.environ.get("DMC_BOUNDED_TEACHER_STATS", "0") != "0"
_early_teacher_stats = os.environ.get("DMC_EARLY_TEACHER_STATS", "1") != "0"


def set_early_teacher_stats(enabled: bool):
    """When enabled (default), a frozen no-grad head (the teacher) whose logits will go to a registered DINOLoss launches
    that loss's teacher statistics pass (row softmax statistics + column sums, main_dino_mc.py:446,468) itself, on a side
    stream, right after its last GEMM -- the HBM-bound pass then runs next to the student head's tensor-bound MLP GEMMs
    instead of after them.  DINOLoss re-validates the result (temperature, center identity and version, row count) and
    runs the pass itself on any mismatch."""
    global _early_teacher_stats
    _early_teacher_stats = bool(enabled)


def set_teacher_overlap(enabled: bool):
    """When enabled, a DINOHead called under torch.no_grad() (the teacher, main_dino_mc.py:264-265,373) enqueues
    its kernels on a per-device side stream and returns immediately; the logits carry a CUDA event that
    dinomc_b200.DINOLoss waits for before reading them.  The teacher head then runs concurrently with the
    student head's forward instead of in front of it.  Contract: such logits must only be consumed by
    dinomc_b200.DINOLoss (any other consumer must first call `dinomc_b200.wait_ready(logits)`)."""
    global _teacher_overlap
    _teacher_overlap = bool(enabled)


def wait_ready(t):
    """Make the current stream wait for a tensor produced by an overlapped (side-stream) head forward."""
    ev = getattr(t, "_dmc_ready_event", None)
    if ev is not None:
        torch.cuda.current_stream(t.device).wait_event(ev)
        t._dmc_ready_event = None
    return t


_stats_streams = {}


def _stats_stream(device, cur):
    key = (torch.device(device).index, cur.cuda_stream)
    st = _stats_streams.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _stats_streams[key] = st
    return st


def _side_stream(device):
    key = torch.device(device).index
    st = _side_streams.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _side_streams[key] = st
    return st


def _trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    """Same sampling procedure as utils/utils.py:529-567 (inverse-CDF of a truncated uniform), so that a
    given torch RNG state yields the same weights as the reference's trunc_normal_."""
    def norm_cdf(x):
        return (1. + math.erf(x / math.sqrt(2.))) / 2.
    with torch.no_grad():
        lo = norm_cdf((a - mean) / std)
        up = norm_cdf((b - mean) / std)
        tensor.uniform_(2 * lo - 1, 2 * up - 1)
        tensor.erfinv_()
        tensor.mul_(std * math.sqrt(2.))
        tensor.add_(mean)
        tensor.clamp_(min=a, max=b)
    return tensor


class _WeightNormLinear(nn.Module):
    """Parameter container equal to `nn.utils.weight_norm(nn.Linear(in, out, bias=False))`:
    parameters `weight_g` [out,1] then `weight_v` [out,in] (that registration order), no bias."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        lin = nn.Linear(in_features, out_features, bias=False)       # default nn.Linear init, as the reference
        w = lin.weight.detach()
        self.weight_g = nn.Parameter(w.norm(dim=1, keepdim=True))
        self.weight_v = nn.Parameter(w.clone())

    @property
    def weight(self):
        """The effective weight g * v / ||v|| (what the reference's pre-hook materialises); for inspection."""
        v = self.weight_v
        return v * (self.weight_g / v.norm(dim=1, keepdim=True))

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias=False, weight_norm"


class DINOHead(nn.Module):
    def __init__(self, in_dim, out_dim, use_bn=False, norm_last_layer=True, nlayers=3, hidden_dim=2048,
                 bottleneck_dim=256):
        super().__init__()
        nlayers = max(nlayers, 1)
        if nlayers == 1:
            self.mlp = nn.Linear(in_dim, bottleneck_dim)
        else:
            layers = [nn.Linear(in_dim, hidden_dim)]
            if use_bn:
                layers.append(nn.BatchNorm1d(hidden_dim))
            layers.append(nn.GELU())
            for _ in range(nlayers - 2):
                layers.append(nn.Linear(hidden_dim, hidden_dim))
                if use_bn:
                    layers.append(nn.BatchNorm1d(hidden_dim))
                layers.append(nn.GELU())
            layers.append(nn.Linear(hidden_dim, bottleneck_dim))
            self.mlp = nn.Sequential(*layers)
        self.apply(self._init_weights)          # before last_layer exists, exactly like the reference (:278)
        self.last_layer = _WeightNormLinear(bottleneck_dim, out_dim)
        self.last_layer.weight_g.data.fill_(1)
        if norm_last_layer:
            self.last_layer.weight_g.requires_grad = False
        self.use_bn = use_bn
        self.precision = None                   # None -> module default (set_default_precision / autocast)
        self._shadow = None                     # operand shadows of a frozen (teacher) head, see set_operand_shadows
        self._loss_ref = None                   # weakref to the DINOLoss bound with bind_loss(); None -> functional's default

    def bind_loss(self, loss_module):
        """Tell this head which `dinomc_b200.DINOLoss` consumes its logits (temperatures and center for the statistics its
        last GEMM / its teacher pass produce on the loss's behalf).  Optional: an unbound head uses the most recently
        constructed or used DINOLoss of the process, which is what the reference's single-loss training script needs; bind
        explicitly when several DINOLoss modules are alive.  `None` removes the binding.  The binding is a weak reference and
        not part of the state_dict.  Whatever the head assumes is re-validated by `DINOLoss.forward`."""
        import weakref
        self._loss_ref = None if loss_module is None else weakref.ref(loss_module)
        return self

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_loss_ref"] = None               # a weak reference cannot be pickled, and a copy is not bound to anything
        return state

    def _loss_module(self):
        if self._loss_ref is not None:
            bound = self._loss_ref()
            if bound is not None:
                return bound
        return Fn._current_loss()

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def _mode(self):
        mode = self.precision or _default_precision
        if mode == "auto":
            mode = "bf16" if torch.is_autocast_enabled() else "fp32"
        return mode

    def _is_inference(self, x):
        """True for a forward that builds no autograd graph: no gradient mode or an input without gradient, and every
        parameter frozen -- the EMA teacher as main_dino_mc.py:264-265,373 runs it (NOT under no_grad there)."""
        if torch.is_grad_enabled() and x.requires_grad:
            return False
        return not any(p.requires_grad for p in self.parameters())

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("dinomc_b200.DINOHead runs on CUDA (sm_100a) only; there is no CPU fallback")
        if _teacher_overlap and self._is_inference(x):
            cur = torch.cuda.current_stream(x.device)
            side = _side_stream(x.device)
            side.wait_stream(cur)                       # inputs (features, EMA'd weights) are ready on `cur`
            with torch.cuda.stream(side):
                out = self._forward(x)
            ev = torch.cuda.Event()
            ev.record(side)
            x.record_stream(side)
            out.record_stream(cur)
            out._dmc_ready_event = ev
            return out
        return self._forward(x)

    # ---- operand shadows (teacher) ----------------------------------------------------------------------------
    def _linears(self):
        return [self.mlp] if isinstance(self.mlp, nn.Linear) else [m for m in self.mlp if isinstance(m, nn.Linear)]

    def _shadow_state(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _mark_shadow_fresh(self):
        if self._shadow is not None:
            self._shadow["state"] = self._shadow_state()

    def _fresh_shadow(self, mode, x):
        """The shadow record if it may be used for this forward, else None (allocating it on first sight of a teacher)."""
        if not _operand_shadows or mode != "bf16" or self.use_bn or not self._is_inference(x):
            return None
        lins = self._linears()
        K, dim = self.last_layer.weight_v.shape
        if Fn.resolve_mode(mode, x.shape[1], dim, K, *[d for lin in lins for d in lin.weight.shape]) != "bf16":
            return None
        sh = self._shadow
        if sh is None or sh["device"] != x.device:
            dev = x.device
            sh = dict(device=dev, state=None, weights=[lin.weight for lin in lins],
                      mlp=[torch.empty(lin.weight.shape, dtype=torch.bfloat16, device=dev) for lin in lins],
                      what=torch.empty((K, dim), dtype=torch.bfloat16, device=dev),
                      scale=torch.empty(K, dtype=torch.float32, device=dev),
                      inv_norm=torch.empty(K, dtype=torch.float32, device=dev))
            self._shadow = sh
            from . import ema
            ema.register_shadow_head(self)
        return sh if (sh["state"] is not None and sh["state"] == self._shadow_state()) else None

    def _forward(self, x):
        mode = self._mode()
        with torch.autocast("cuda", enabled=False):
            if self.use_bn:
                # use_bn_in_head (utils/vision_transformer.py:268-274): every Linear runs on the tcgen05 GEMM (bias in the
                # epilogue, own wgrad / dgrad); BatchNorm1d -- or SyncBatchNorm after the reference's convert_sync_batchnorm
                # (main_dino_mc.py:250-252), with its cross-rank statistics -- and GELU stay torch modules between them.
                lins = self._linears()
                mode_r = Fn.resolve_mode(mode, x.shape[1], *[d for lin in lins for d in lin.weight.shape])
                h = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
                for m in (self.mlp if isinstance(self.mlp, nn.Sequential) else [self.mlp]):
                    if isinstance(m, nn.Linear):
                        h, _ = Fn.LinearFn.apply(mode_r, h, None, m.weight, m.bias, None, None, False)
                    else:
                        h = m(h)
                z = h
                prepared = None
            else:
                linears = self._linears()
                wb = []
                for lin in linears:
                    wb += [lin.weight, lin.bias]
                sh = self._fresh_shadow(mode, x)
                if sh is not None:
                    # operands written by the last EMA pass: no casts, no weight-norm pass
                    prepared = dict(mode="bf16", wop=Fn.Operand(sh["what"]), scale=sh["scale"], inv_vnorm=sh["inv_norm"],
                                    gmax=None, region=None)
                    z = Fn.mlp_forward(mode, x, wb, w_ops=[Fn.Operand(t) for t in sh["mlp"]])
                elif _wn_after_first_gemm:
                    # the last layer's weight-norm materialisation (HBM-bound, 64 MiB read at out_dim 65536) does not depend
                    # on the MLP: fork it onto the auxiliary stream once the first GEMM is queued, so that it runs next to
                    # the remaining tensor-bound GEMMs (forked before them it takes the whole machine and delays the chain)
                    box = {}

                    def fork():
                        box["p"] = Fn.last_layer_weights(mode, self.last_layer.weight_g, self.last_layer.weight_v,
                                                         self.last_layer.in_features)
                    z = Fn.mlp_forward(mode, x, wb, after_first_gemm=fork)
                    prepared = box["p"]
                else:
                    prepared = Fn.last_layer_weights(mode, self.last_layer.weight_g, self.last_layer.weight_v,
                                                     self.last_layer.in_features)
                    z = Fn.mlp_forward(mode, x, wb)
            link = {"loss": self._loss_module(), "stats": None}
            out = Fn.NormLastLayerFn.apply(mode, z, self.last_layer.weight_g, self.last_layer.weight_v, prepared, link)
            if link["stats"] is not None:       # statistics the GEMM epilogue produced for dinomc_b200.DINOLoss
                out._dmc_stats = link["stats"]
            elif _early_teacher_stats and out.dtype in (torch.bfloat16, torch.float32) and self._is_inference(x):
                self._launch_teacher_stats(out)
            return out

    def _launch_teacher_stats(self, out):
        """The registered DINOLoss's teacher pass over `out`, on a side stream (see set_early_teacher_stats)."""
        from . import ops
        loss_mod = self._loss_module()
        if loss_mod is None or loss_mod.center.shape[-1] != out.shape[1] or loss_mod.center.device != out.device:
            return
        if out.shape[0] % loss_mod.teacher_crops_number:
            return
        loss_mod.sync_center()
        cen = loss_mod.center
        cur = torch.cuda.current_stream(out.device)
        side = _stats_stream(out.device, cur)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            bounds = None
            if _bounded_teacher_stats:
                # |logit| <= the largest gain (unit rows against weight-normed rows) and max |center|, both as device scalars:
                # lets the pass use a fixed shift instead of a per-row maximum (about half the instructions)
                bounds = torch.empty(2, dtype=torch.float32, device=out.device)
                ops.absmax_into(self.last_layer.weight_g.detach(), bounds[0:1])
                ops.absmax_into(cen.detach(), bounds[1:2])
            t_stats, colsum = ops.teacher_stats_colsum(out, cen.reshape(-1), loss_mod._last_inv_tt, bounds=bounds)
        ev = torch.cuda.Event()
        ev.record(side)
        out.record_stream(side)
        cen.record_stream(side)
        out._dmc_stats = dict(kind="teacher_final", scale=loss_mod._last_inv_tt, t_stats=t_stats, colsum=colsum, event=ev,
                              stream=side, center_ptr=cen.data_ptr(), center_version=cen._version, rows=out.shape[0])
