"""Gradient all-reduce for the data-parallel step (the exchange DDP performs at main_dino_mc.py:260).

`GradAllReduce` averages each parameter's gradient over the ranks on a dedicated communication stream so the
NCCL transfers (89 MB of head gradients at K = 65536) overlap the rest of the backward pass and the EMA update:

  * a large gradient (the weight-normed last layer's dv, 64 MB) is reduced the moment the kernel that produced
    it has finished (`ops.mark_ready` event), not when its autograd node returns;
  * the small gradients are batched into one coalesced NCCL launch once all of them exist;
  * while a reducer is active the persistent GEMM grids of the backward pass leave `reserve_sms` SMs free, so the
    NCCL kernel neither queues behind a one-CTA-per-SM GEMM nor starves its tail CTAs.

`compress="bf16"` (the bf16-GEMM mode's exchange; same idea as torch's `bf16_compress_hook` for DDP) halves the bytes:

  * the weight-normed last layer is exchanged BEFORE its weight-norm backward: the wgrad GEMM writes dW in bf16, the
    ranks average dW (32 MB instead of the 64 MB fp32 dv) and every rank then runs the weight-norm backward on the
    averaged dW on the communication stream -- that pass is linear in dW, so the result is the averaged dv (and dg).
    No extra pass over memory: the GEMM writes half the bytes and the weight-norm backward reads half the bytes;
  * the small gradients are narrowed into ONE flat bf16 buffer (one launch), averaged with one all-reduce, and widened
    back into their fp32 `.grad` tensors (one launch).

Unlike `DistributedDataParallel` it keeps no Python-side reducer state between steps, which makes the whole step
(collectives included) capturable in one CUDA graph (`StepGraph`).  `torch.distributed` (NCCL over NVLink /
NVSwitch) does the transport.
"""
from __future__ import annotations

import os
import sys
import time

import torch
import torch.distributed as dist

from . import ops

_DEBUG = bool(os.environ.get("DMC_XRANK_DEBUG"))
_BIG = 32 << 20      # bytes: gradients at least this large get their own all-reduce
_FLUSH = 64 << 20    # bytes: pending small gradients are sent as one coalesced all-reduce once they reach this, i.e. normally ONE
                     # launch after the last wgrad, overlapping the EMA (2 GPUs: 0.9715 ms with a 1 MB threshold = 4 launches, 0.9463 ms with one)
if os.environ.get("DMC_REDUCER_FLUSH_MB"):      # env: timing experiments only
    _FLUSH = int(float(os.environ["DMC_REDUCER_FLUSH_MB"]) * (1 << 20))


class GradAllReduce:
    """One instance per model and process.  Not supported: gradient accumulation over several backward passes before `wait()`
    (a gradient already sitting in `.grad` would be averaged again) -- call `wait()` and consume / clear the gradients after
    every backward pass, as main_dino_mc.py does."""

    def __init__(self, params, group=None, reserve_sms: int = 16, compress=None, transport="nccl"):
        """transport = "nccl": torch.distributed all-reduces (fp32, or bf16 with compress="bf16").
        transport = "peer": the bf16 exchange runs on libdinomc's own NVLink / NVSwitch all-reduce kernel over symmetric
        memory (xrank.SymmetricBuffer; needs compress="bf16"): the last layer's wgrad GEMM writes dW straight into a
        symmetric buffer, ONE kernel averages it over the ranks (multimem.ld_reduce / multimem.st through the switch),
        the weight-norm backward runs on the average; the small gradients travel as one flat bf16 buffer whose all-reduce
        kernel also widens the result back into the fp32 .grad tensors.  The kernel is one 256-thread, 32-register
        CTA per SM: it is co-resident with the backward GEMMs, so no SMs are reserved."""
        if compress not in (None, "bf16"):
            raise ValueError(f"GradAllReduce: compress must be None or 'bf16', got {compress!r}")
        if transport not in ("nccl", "peer"):
            raise ValueError(f"GradAllReduce: transport must be 'nccl' or 'peer', got {transport!r}")
        if transport == "peer" and compress != "bf16":
            raise ValueError("GradAllReduce: transport='peer' exchanges bf16 buffers; pass compress='bf16'")
        self.group = group
        self.compress = compress
        self.transport = transport
        self._dw_buf = None             # peer transport: symmetric bf16 buffer of the last layer's dW
        self._small_bufs = {}           # peer transport: flat symmetric bf16 buffers of the small gradients, one per flush shape
        self._keep = []                 # tensors produced on the caller's stream and read on the communication stream: kept
                                        # alive until wait() (inside a graph capture record_stream() does not defer reuse)
        self._flush_bytes = _FLUSH
        self._xrank_ctas = int(os.environ.get("DMC_XRANK_CTAS", "148"))
        if transport == "peer":
            # One exchange CTA per SM (256 threads, 32 registers), co-resident with the backward GEMMs; no SMs reserved.
            # MEASURED at 2 GPUs (profiles/r02_exchange_ab.md): 148 CTAs / 0 reserved 0.909 ms; 48 CTAs on 24 reserved SMs
            # 0.939; 32 on 16 0.956; 16 on 8 1.018 -- NVLink wants the loads of many SMs in flight, and the GEMMs lose more
            # from 16-24 missing SMs than from sharing.  DMC_XRANK_CTAS / DMC_XRANK_RESERVE_SMS override.
            reserve_sms = int(os.environ.get("DMC_XRANK_RESERVE_SMS", "0"))
            # flush the small gradients in two batches: the later MLP layers' (19 of 25 MB) leave while the first layer's
            # backward still runs, only the first layer's 3 MB are exchanged after the last wgrad
            self._flush_bytes = int(float(os.environ.get("DMC_PEER_FLUSH_MB", "8")) * (1 << 20))
        self.params = [p for p in params if p.requires_grad]
        self._by_ptr = {p.data_ptr(): p for p in self.params}
        self.comm = torch.cuda.Stream(priority=-1)
        self.comm2 = torch.cuda.Stream(priority=-1)
        self._late = None
        self._late_used = False
        self._narrow_done = None
        self._late_wn = os.environ.get("DMC_LATE_WN_BWD", "1") != "0"
        self._pending = []
        self._pending_bytes = 0
        self._seen = 0
        self._claimed = set()
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.bytes_per_step = sum(p.numel() * (2 if compress == "bf16" else p.element_size()) for p in self.params)
        if reserve_sms > 0:
            ops.backward_max_ctas = 148 - reserve_sms
        from . import functional
        functional.grad_exchange_active = os.environ.get("DMC_REDUCER_AUXWN", "") != "1"      # env: timing experiments only
        functional.grad_exchange = self
        ops.track_ready = True
        # early-launched (PDL) GEMM CTAs hold SMs while they wait for their predecessor, which delays the NCCL kernels
        # that share the machine: measured 0.992 -> 0.969 ms per step at 2 GPUs without it
        from . import _lib
        keep_pdl = os.environ.get("DMC_REDUCER_PDL", "1" if transport == "peer" else "") == "1"   # env: timing experiments only
        self._pdl_prev = _lib.load().dmc_set_pdl(1 if keep_pdl else 0)

    def _hook(self, p):
        if id(p) in self._claimed:
            # exchange_last_layer already produced this parameter's averaged gradient.  torch calls the post-accumulate
            # hook of a parameter even when the backward returned None for it (measured on torch 2.11), so without this the
            # last layer was exchanged a second time as a "big" gradient.
            self._claimed.discard(id(p))
            return
        g = p.grad
        self._seen += 1
        if g is None:                   # the hook also fires for a parameter that received no gradient in this pass
            if self._seen == len(self.params):
                self._flush()
            return
        skip = os.environ.get("DMC_REDUCER_SKIP", "")       # timing experiments only: "big" / "small"
        big = g.numel() * g.element_size() >= _BIG
        if (skip == "big" and big) or (skip == "small" and not big) or skip == "all":
            if self._seen == len(self.params):
                self._flush()
            return
        if big:
            ev = ops.ready_events.pop(g.data_ptr(), None)
            if ev is not None:
                self.comm.wait_event(ev)                 # start as soon as the producing kernel is done
            else:
                self.comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm):
                if self._compressible(g):
                    self._exchange_bf16([g])
                else:
                    dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
            g.record_stream(self.comm)
        else:
            self._pending.append((g, ops.ready_events.pop(g.data_ptr(), None)))
            self._pending_bytes += g.numel() * g.element_size()
        if self._seen == len(self.params) or self._pending_bytes >= self._flush_bytes:
            self._flush()

    def _flush(self):
        final = self._seen >= len(self.params)
        if self._pending:
            if all(ev is not None for _, ev in self._pending):
                for _, ev in self._pending:              # start as soon as the producing kernels are done
                    self.comm.wait_event(ev)
            else:
                self.comm.wait_stream(torch.cuda.current_stream())
            grads = [g for g, _ in self._pending]
            narrow = [g for g in grads if self._compressible(g)]
            grads = [g for g in grads if not self._compressible(g)]
            with torch.cuda.stream(self.comm):
                if narrow:
                    self._exchange_bf16(narrow)
            for g in narrow:
                g.record_stream(self.comm)
            with torch.cuda.stream(self.comm):
                try:
                    if grads:
                        with dist._coalescing_manager(group=self.group, device=grads[0].device, async_ops=False):
                            for g in grads:
                                dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
                except (AttributeError, TypeError, RuntimeError):
                    for g in grads:
                        dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
            for g in grads:
                g.record_stream(self.comm)
            self._pending = []
            self._pending_bytes = 0
        if final:
            self._run_late()            # AFTER the last small exchange is queued: launched first, its CTAs would keep that
                                        # exchange's tiny narrow kernel off the SMs for the whole pass (measured: +20 us)
        if self._seen >= len(self.params):
            self._seen = 0

    # ---- bf16 exchange --------------------------------------------------------------------------------------
    def _compressible(self, g):
        return self.compress == "bf16" and g.dtype == torch.float32 and g.is_contiguous()

    def last_layer_buffer(self, K, dim):
        """peer transport: the [K, dim] bf16 view of the symmetric buffer the last layer's wgrad GEMM writes dW into
        (None for the NCCL transport).  Allocated (collectively) on first use -- outside any graph capture."""
        if self.transport != "peer":
            return None
        if self._dw_buf is None or self._dw_buf.numel != K * dim:
            from .xrank import SymmetricBuffer
            self._dw_buf = SymmetricBuffer(K * dim, torch.bfloat16, group=self.group, ctas=self._xrank_ctas)
        return self._dw_buf.tensor.view(K, dim)

    def _exchange_bf16_peer(self, grads):
        """peer transport, on the communication stream: fp32 gradients -> flat symmetric bf16 buffer (one launch) -> ONE
        all-reduce kernel that also widens the averaged values back into the fp32 gradients."""
        offs, total = [], 0
        for g in grads:
            offs.append(total)
            total += (g.numel() + 7) & ~7
        key = tuple((g.numel()) for g in grads)
        buf = self._small_bufs.get(key)
        if buf is None:
            from .xrank import SymmetricBuffer
            # few CTAs for small buffers: each CTA costs 2 x world remote signals per barrier, and a 1.6 MB exchange on 148
            # CTAs is all barrier (measured 50 us at 8 GPUs)
            ctas = max(8, min(self._xrank_ctas, (2 * total) >> int(os.environ.get("DMC_XRANK_BYTES_PER_CTA_LOG2", "16"))))
            buf = self._small_bufs[key] = SymmetricBuffer(total, torch.bfloat16, group=self.group, ctas=ctas)
        t0 = time.perf_counter()
        flat = buf.tensor
        views = [flat[o:o + g.numel()] for o, g in zip(offs, grads)]
        ops.narrow_bf16_into([g.view(-1) for g in grads], views)
        self._narrow_done = torch.cuda.Event()
        self._narrow_done.record()              # (current stream = the communication stream) see _run_late
        if _DEBUG:
            self._dbg("narrow launch", t0)
        world = dist.get_world_size(self.group)
        if len(grads) <= 8:             # the head has 6: the exchange kernel widens the result into the fp32 gradients itself
            buf.allreduce_(1.0 / world, widen_to=[g.view(-1) for g in grads], widen_offsets=offs)
        else:
            buf.allreduce_(1.0 / world)
            ops.widen_bf16_batch(views, [g.view(-1) for g in grads])

    def _exchange_bf16(self, grads):
        """On the current (communication) stream: fp32 gradients -> one flat bf16 buffer -> all-reduce(AVG) -> back."""
        if self.transport == "peer":
            return self._exchange_bf16_peer(grads)
        offs, total = [], 0
        for g in grads:
            offs.append(total)
            total += (g.numel() + 7) & ~7                      # keep every slice 16-byte aligned
        flat = torch.empty(total, dtype=torch.bfloat16, device=grads[0].device)
        views = [flat[o:o + g.numel()] for o, g in zip(offs, grads)]
        if total != sum(g.numel() for g in grads):
            flat.zero_()                                       # padding travels through the all-reduce: keep it finite
        ops.narrow_bf16_into([g.view(-1) for g in grads], views)
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        ops.widen_bf16_batch(views, [g.view(-1) for g in grads])

    def claim_last_layer(self, v_ptr, g_ptr=None):
        """For NormLastLayerFn.backward: the parameters (weight_v[, weight_g]) whose gradients may be produced from an
        averaged dW, or None when that route does not apply (no compression, foreign parameters, or a gradient already
        accumulated in `.grad` -- then the regular route through autograd runs)."""
        if self.compress != "bf16" or self.transport != "peer":
            # NCCL transport: the last layer's dv takes the regular route (narrowed to bf16 like the small gradients).  The
            # averaged-dW route over NCCL produced rank-dependent gradients at out_dim 65536 (bench.py's exchange check:
            # max diff 0.73, replicas not identical) -- masked in round 1 by a second exchange of the same parameter
            return None
        pv = self._by_ptr.get(v_ptr)
        if pv is None or pv.grad is not None:
            return None
        pg = None
        if g_ptr is not None:
            pg = self._by_ptr.get(g_ptr)
            if pg is None or pg.grad is not None:
                return None
        return pv, pg

    def _dbg(self, what, t0):
        dt = time.perf_counter() - t0
        if dt > 2e-3:
            print(f"[reducer] {what} took {dt * 1e3:.1f} ms on the host", file=sys.stderr, flush=True)

    def exchange_last_layer(self, dw, weightnorm_bwd, pv, pg):
        """Average the bf16 dW over the ranks, then run `weightnorm_bwd()` (-> dv, dg) on the averaged dW, all on the
        communication stream; the results become `pv.grad` / `pg.grad` directly (they are complete after `wait()`)."""
        cur = torch.cuda.current_stream()
        ev = ops.ready_events.pop(dw.data_ptr(), None)
        if ev is not None:
            self.comm.wait_event(ev)
        else:
            self.comm.wait_stream(cur)
        t0 = time.perf_counter()
        peer = self.transport == "peer" and self._dw_buf is not None and dw.data_ptr() == self._dw_buf.tensor.data_ptr()
        with torch.cuda.stream(self.comm):
            if peer:
                self._dw_buf.allreduce_(1.0 / dist.get_world_size(self.group))      # in place, in the symmetric buffer
            else:
                dist.all_reduce(dw, op=dist.ReduceOp.AVG, group=self.group)
            if _DEBUG:
                self._dbg("last-layer all-reduce launch", t0)
                t0 = time.perf_counter()
            if peer and self._late_wn:
                # The weight-norm backward of the averaged dW (HBM-bound, ~30 us, thousands of CTAs) is NOT run here: behind
                # the exchange it lands in the middle of the MLP backward and its CTAs keep the next GEMMs off the SMs
                # (measured at 8 GPUs: MLP backward starts 20 us late).  It runs at the final flush instead, on a second
                # stream, next to the last (latency-bound) small exchange -- both are needed only by wait().
                done = torch.cuda.Event()
                done.record(self.comm)
                self._late = (weightnorm_bwd, pv, pg, done)
                dv = dg = None
            else:
                with ops.no_pdl():      # must not become resident before the exchange has finished (see ops.no_pdl)
                    dv, dg = weightnorm_bwd()
        if _DEBUG:
            self._dbg("weight-norm backward launch", t0)
        if not peer:                     # symmetric buffers are persistent: nothing for the caching allocator to track
            dw.record_stream(self.comm)
            self._keep.append(dw)        # ... and inside a graph capture only a live reference keeps the block from being reused
        if dv is not None:
            dv.record_stream(cur)
            pv.grad = dv.view_as(pv)
        self._claimed.add(id(pv))
        self._seen += 1
        if pg is not None:
            if dg is not None:
                dg.record_stream(cur)
                pg.grad = dg.view_as(pg)
            self._claimed.add(id(pg))
            self._seen += 1
        if self._seen == len(self.params):
            self._flush()

    def _run_late(self):
        """The deferred weight-norm backward (see exchange_last_layer), on the second stream, after everything the backward
        pass has queued on the caller's stream so far (i.e. after its last GEMM)."""
        if self._late is None:
            return
        fn, pv, pg, done = self._late
        self._late = None
        self._late_used = True
        cur = torch.cuda.current_stream()
        end = torch.cuda.Event()
        end.record(cur)
        self.comm2.wait_event(done)
        self.comm2.wait_event(end)
        if self._narrow_done is not None:       # let the last exchange's tiny narrow kernel onto the SMs first
            self.comm2.wait_event(self._narrow_done)
            self._narrow_done = None
        with torch.cuda.stream(self.comm2), ops.no_pdl():
            dv, dg = fn()
        dv.record_stream(cur)
        pv.grad = dv.view_as(pv)
        if pg is not None and dg is not None:
            dg.record_stream(cur)
            pg.grad = dg.view_as(pg)

    def wait(self):
        """Join: later work on the current stream sees the averaged gradients."""
        self._flush()
        self._run_late()
        self._seen = 0
        torch.cuda.current_stream().wait_stream(self.comm)
        if self._late_used:
            torch.cuda.current_stream().wait_stream(self.comm2)
            self._late_used = False
        self._keep.clear()
        ops.ready_events.clear()         # every gradient of this step has been taken: nothing recorded so far is still needed

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
        ops.backward_max_ctas = 0
        from . import functional, _lib
        functional.grad_exchange_active = False
        if functional.grad_exchange is self:
            functional.grad_exchange = None
            ops.track_ready = False
            ops.ready_events.clear()
        _lib.load().dmc_set_pdl(self._pdl_prev)
