#!/usr/bin/env python
"""bench.py -- the DINO-MC head + loss + center + EMA step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2] [--mode bf16|fp32]

One "step"  = teacher-head forward (no grad) on [G*B, D] + student-head forward on [C*B, D]
+ DINOLoss.forward (teacher statistics, column sum, center all-reduce over ranks, center EMA) + backward to
the head parameters AND the feature tensor (+ DDP/NCCL gradient all-reduce when N > 1) + EMA of the whole
backbone+head parameter list.  Synthetic features, reference-shaped random-init weights.

Prints ONE JSON line (rank 0).  `value` = samples/s with inputs resident in HBM; `e2e` = the same step
through the public modules with HOST (pinned) feature buffers copied in and the loss read back each step;
`roofline` = the dominant kernel, timed live with CUDA events; `cpu_baseline` = oracle/torch_port.py (the
reference's eager-PyTorch path restated) timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# ----------------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs; cfg2 is the headline the metric is quoted on)
# ----------------------------------------------------------------------------------------------------
WORKLOADS = {
    "cfg1": dict(arch="vit_small_p8", D=384, K=65536, B=32, C=8, G=2, note="reference CPU-runnable case"),
    "cfg2": dict(arch="vit_small_p8", D=384, K=65536, B=256, C=8, G=2, note="ViT-S/8 headline, per GPU"),
    "cfg3": dict(arch="resnet50", D=2048, K=65536, B=512, C=8, G=2, note="ResNet-50 features"),
    "cfg4": dict(arch="swin_t", D=768, K=65536, B=256, C=8, G=2, note="Swin-tiny features"),
    "cfg5": dict(arch="wide_resnet50_2", D=2048, K=65536, B=256, C=8, G=2, note="WRN-50-2, out_dim sweep via --out-dim"),
}
H, BN = 2048, 256   # DINOHead hidden / bottleneck (utils/vision_transformer.py:261)


def backbone_param_shapes(arch: str):
    """Shapes of the backbone parameters the EMA runs over (no backbone forward is ever executed)."""
    if arch == "vit_small_p8":   # utils/vision_transformer.py:134-256, vit_small(patch_size=8)
        E, depth, patch, img = 384, 12, 8, 224
        shapes = [(1, 1, E), (1, 1 + (img // patch) ** 2, E), (E, 3, patch, patch), (E,)]
        for _ in range(depth):
            shapes += [(E,), (E,), (3 * E, E), (3 * E,), (E, E), (E,), (E,), (E,), (4 * E, E), (4 * E,), (E, 4 * E), (E,)]
        shapes += [(E,), (E,)]
        return shapes
    import torchvision
    with torch.device("meta"):
        m = getattr(torchvision.models, arch)()
    for attr in ("fc", "head"):                # MultiCropWrapper replaces these by Identity (utils/utils.py:623)
        if hasattr(m, attr):
            setattr(m, attr, torch.nn.Identity())
    return [tuple(p.shape) for p in m.parameters()]


def roofline_model(w, P, logit_bytes):
    """Algorithmic FLOPs and bytes per step per GPU (SURVEY.md 8d formulas)."""
    D, K, B, C, G = w["D"], w["K"], w["B"], w["C"], w["G"]
    Ns, Nt = C * B, G * B
    f_mlp = lambda n: 2 * n * (D * H + H * H + H * BN)
    f_last = lambda n: 2 * n * BN * K
    flops = 3 * (f_mlp(Ns) + f_last(Ns)) + f_mlp(Nt) + f_last(Nt)
    e = logit_bytes
    nbytes = (e * K * (6 * Ns + 4 * Nt) + 7 * 4 * K * BN + 12 * P + 4 * (D + 2 * H + BN) * (3 * Ns + Nt)
              + 16 * (D * H + H * H + H * BN) + 12 * K)
    return flops, nbytes


def op_models(w, P, e):
    """{op tag: (algorithmic bytes, algorithmic flops)} of ALL launches of that op in one step;
    e = bytes per stored logit.  Tags are the ones ops.py times its C-ABI calls under."""
    D, K, B, C, G = w["D"], w["K"], w["B"], w["C"], w["G"]
    Ns, Nt = C * B, G * B
    mlp_w = D * H + H * H + H * BN                      # MLP weights of one head
    f_mlp = lambda n: 2 * n * mlp_w
    f_last = lambda n: 2 * n * BN * K
    act = lambda n: n * (D + 2 * H + BN)                # activation elements of one head forward
    m = {
        # loss
        "ce_fused": (e * K * (2 * Ns + Nt), 0), "ce_bwd": (e * K * (2 * Ns + Nt), 0), "ce_fwd": (e * K * (Ns + Nt), 0),
        "teacher_stats_colsum": (e * K * Nt, 0),
        # EMA (+ the teacher's operand shadows it emits)
        "ema": (12 * P + (2 * (mlp_w + K * BN) + 8 * K if e == 2 else 0), 0),   # + the teacher's bf16 operand shadows (bf16 mode)
        # last layer
        "gemm_last_fwd_student": (e * K * Ns + 2 * BN * (K + Ns), f_last(Ns)),
        "gemm_last_fwd_teacher": (e * K * Nt + 2 * BN * (K + Nt), f_last(Nt)),
        "gemm_last_wgrad": (e * K * Ns + 2 * BN * Ns + (2 if e == 2 else 4) * K * BN, f_last(Ns)),     # dW stored bf16 in the bf16 mode
        "gemm_last_dgrad": (e * K * Ns + 2 * K * BN + 4 * BN * Ns, f_last(Ns)),
        "weightnorm_fwd": ((4 + e) * K * BN, 0),                 # student only: the teacher's operand comes out of the EMA pass
        "weightnorm_bwd": ((8 + (2 if e == 2 else 4)) * K * BN, 0),
        "xrank_allreduce": (2 * (K * BN + mlp_w), 0),           # N > 1: bf16 dW + the small gradients, once over NVLink each way
        "cast_bf16": (6 * (Ns * D + Nt * D + mlp_w), 0),
        "colsum": (2 * Ns * (2 * H + BN), 0),
        "normalize_fwd": (10 * BN * (Ns + Nt), 0), "normalize_bwd": (12 * BN * Ns, 0),
    }

    def add(tag, nbytes, flops):            # D == H (ResNet-50 features): the first two Linears share a tag -- their models add up
        b0, f0 = m.get(tag, (0, 0))
        m[tag] = (b0 + nbytes, f0 + flops)

    # MLP, one op per Linear (in x out) -- bf16 operands / activations, fp32 weight gradients; student + teacher forward; the
    # hidden layers' forward also writes gelu'(z) for the backward (student rows), their dgrad reads it
    for li, (fi, fo) in enumerate(((D, H), (H, H), (H, BN))):
        out_b = 2 if li < 2 else 4                     # the bottleneck output (input of F.normalize) is stored fp32
        add(f"gemm_mlp_fwd_{fi}x{fo}", (Ns + Nt) * (2 * fi + out_b * fo) + (2 * Ns * fo if li < 2 else 0) + 2 * 2 * fi * fo, 2 * (Ns + Nt) * fi * fo)
        din_b = 2 if li > 0 else 4                     # the gradient of the features leaves in fp32
        add(f"gemm_mlp_dgrad_{fi}x{fo}", Ns * (2 * fo + din_b * fi) + (2 * Ns * fi if li > 0 else 0) + 2 * fi * fo, 2 * Ns * fi * fo)
        add(f"gemm_mlp_wgrad_{fi}x{fo}", 2 * Ns * (fi + fo) + 4 * fi * fo, 2 * Ns * fi * fo)
    return m


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source="fallback")


# ----------------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed region
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._t = [], set(), None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
class Step:
    """Owns the modules/buffers of one rank and runs one whole step on the current stream."""

    def __init__(self, w, mode, rank, world, device, ddp=False, reserve_sms=16, force_reducer=False, compress=None, transport="nccl"):
        import dinomc_b200 as D
        self.D, self.w, self.world, self.device = D, w, world, device
        torch.manual_seed(0)                                          # identical weights on every rank
        Din, K, B, C, G = w["D"], w["K"], w["B"], w["C"], w["G"]
        self.student = D.DINOHead(Din, K).to(device)
        self.teacher = D.DINOHead(Din, K).to(device)
        self.teacher.load_state_dict(self.student.state_dict())      # main_dino_mc.py:262
        for p in self.teacher.parameters():
            p.requires_grad = False
        self.student.precision = self.teacher.precision = mode
        self.loss_mod = D.DINOLoss(K, C, 0.04, 0.04, 0, 100, teacher_crops_number=G).to(device)
        g = torch.Generator(device="cpu").manual_seed(99)
        shapes = backbone_param_shapes(w["arch"])
        self.bb_student = [(torch.randn(s, generator=g) * 0.02).to(device) for s in shapes]
        self.bb_teacher = [t.clone() for t in self.bb_student]
        # zip order of main_dino_mc.py:405: MultiCropWrapper registers backbone first, then head
        self.ema_student = self.bb_student + list(self.student.parameters())
        self.ema_teacher = self.bb_teacher + list(self.teacher.parameters())
        self.P = sum(t.numel() for t in self.ema_student)
        self.n_tensors = len(self.ema_student)
        self.model = self.student
        self.reducer = None
        if world > 1 or force_reducer:
            if ddp:
                from torch.nn.parallel import DistributedDataParallel as DDP
                self.model = DDP(self.student, device_ids=[device.index])  # main_dino_mc.py:260
            else:
                # same exchange (mean of the head gradients over ranks), graph-capturable, overlapped with bwd + EMA
                self.reducer = D.GradAllReduce(self.student.parameters(), reserve_sms=reserve_sms, compress=compress, transport=transport)
        gs = torch.Generator(device="cpu").manual_seed(1234 + rank)
        gt = torch.Generator(device="cpu").manual_seed(4321 + rank)
        self.x_student_host = torch.randn(C * B, Din, generator=gs).pin_memory()
        self.x_teacher_host = torch.randn(G * B, Din, generator=gt).pin_memory()
        self.x_student = self.x_student_host.to(device).requires_grad_(True)
        self.x_teacher = self.x_teacher_host.to(device)
        self.loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        self.m = 0.996
        # e2e pipeline: the NEXT step's features travel host->device on a copy stream while this step computes
        self.copy_stream = torch.cuda.Stream()
        self.stage = [(torch.empty_like(self.x_student), torch.empty_like(self.x_teacher)) for _ in range(2)]
        self.stage_ready = [None, None]
        self.stage_consumed = [None, None]      # event: the step that read stage[i] has copied it into the graph's inputs
        self.stage_idx = 0
        # e2e loss read-back: one pinned slot per step parity; the host reads step i-1's loss while step i runs
        self.loss_slots = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.loss_events = [None, None]
        self.e2e_count = 0

    def run(self, x_student=None, x_teacher=None):
        xs = self.x_student if x_student is None else x_student
        xt = self.x_teacher if x_teacher is None else x_teacher
        for p in self.student.parameters():
            p.grad = None
        xs.grad = None
        dbg = os.environ.get("DMC_XRANK_DEBUG") and not torch.cuda.is_current_stream_capturing()
        t0 = time.perf_counter()
        with torch.no_grad():
            t_out = self.teacher(xt)
        s_out = self.model(xs)
        loss = self.loss_mod(s_out, t_out, 0)
        t1 = time.perf_counter()
        loss.backward()
        if dbg:
            print(f"[step host ms] forward+loss {1e3 * (t1 - t0):.1f}  backward {1e3 * (time.perf_counter() - t1):.1f}", file=sys.stderr, flush=True)
        if self.reducer is not None:
            # main_dino_mc.py:383-406: the optimizer consumes the averaged gradients BEFORE the EMA reads the student, so
            # the EMA must not be used to hide the gradient exchange: join the exchange first
            self.reducer.wait()
        self.D.ema_update_(self.ema_teacher, self.ema_student, self.m)
        self.loss_mod.sync_center()         # the all-reduced center is complete too (no-op unless it ran asynchronously)
        return loss

    def verify_exchange(self):
        """N > 1, after the timed region: numbers evidence for the gradient exchange.  One eager step WITH the reducer,
        one WITHOUT (local gradients, averaged here by a plain fp32 NCCL all-reduce); reports whether every rank holds
        bit-identical exchanged gradients and how far they are from the plain mean of the local gradients."""
        import torch.distributed as dist
        params = [p for p in self.student.parameters() if p.requires_grad]
        state = (self.loss_mod.center.detach().clone(), [t.detach().clone() for t in self.ema_teacher])
        self.run()
        torch.cuda.synchronize()
        got = [p.grad.detach().clone() for p in params]
        with torch.no_grad():                                   # same inputs and state for the second pass
            self.loss_mod.sync_center()
            self.loss_mod.center.copy_(state[0])
            for t, q in zip(self.ema_teacher, state[1]):
                t.copy_(q)
        red, self.reducer = self.reducer, None
        red.remove()
        try:
            self.run()
            torch.cuda.synchronize()
            ref = [p.grad.detach().clone() for p in params]
        finally:
            self.reducer = red
        worst, per = 0.0, {}
        names = [n for n, p in self.student.named_parameters() if p.requires_grad]
        for n, g, r in zip(names, got, ref):
            dist.all_reduce(r, op=dist.ReduceOp.AVG)
            e = float((g.double() - r.double()).abs().max()) / max(float(r.double().abs().max()), 1e-30)
            per[n] = e
            worst = max(worst, e)
        # bit-identity across ranks: every rank's byte-level checksum of the exchanged gradients must agree
        acc = torch.zeros((), dtype=torch.int64, device=self.device)
        for g in got:
            acc += g.contiguous().view(torch.int32).to(torch.int64).sum()
        lo, hi = acc.clone(), acc.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        w_t = torch.tensor([worst], device=self.device, dtype=torch.float64)
        dist.all_reduce(w_t, op=dist.ReduceOp.MAX)
        return {"grads_bit_identical_across_ranks": bool(int(lo) == int(hi)), "checksum": int(acc),
                "max_rel_diff_vs_mean_of_local_grads": float(w_t), "tensors": len(got),
                "rel_diff_per_tensor_rank0": {k: float("%.3g" % v) for k, v in per.items()}}

    def _prefetch(self, idx):
        if self.stage_consumed[idx] is not None:             # do not overwrite a stage its consumer has not copied out yet
            self.copy_stream.wait_event(self.stage_consumed[idx])
        with torch.cuda.stream(self.copy_stream), torch.no_grad():
            self.stage[idx][0].copy_(self.x_student_host, non_blocking=True)
            self.stage[idx][1].copy_(self.x_teacher_host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.stage_ready[idx] = ev

    def run_e2e(self, graph=None):
        """Public-API step from HOST buffers: H2D of this step's features, D2H of the loss, every step.
        With a StepGraph the features are copied into the graph's static input tensors and the captured
        step is replayed; otherwise the modules are called eagerly."""
        if graph is not None:
            main = torch.cuda.current_stream()
            cur, nxt = self.stage_idx, self.stage_idx ^ 1
            if self.stage_ready[cur] is None:
                self._prefetch(cur)                      # very first step: nothing was prefetched yet
            main.wait_event(self.stage_ready[cur])
            with torch.no_grad():                        # device-to-device into the graph's static inputs
                self.x_student.copy_(self.stage[cur][0], non_blocking=True)
                self.x_teacher.copy_(self.stage[cur][1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(main)
            self.stage_consumed[cur] = ev
            loss = graph.replay()
            slot = self.e2e_count & 1
            self.loss_slots[slot].copy_(loss.detach(), non_blocking=True)       # D2H of THIS step's loss
            lev = torch.cuda.Event()
            lev.record(main)
            self.loss_events[slot] = lev
            self._prefetch(nxt)                          # H2D of the next step's features overlaps this step
            self.stage_idx = nxt
            self.e2e_count += 1
            # the host waits for (and reads) the PREVIOUS step's loss: every step's loss crosses to the host inside the timed
            # region, but the GPU never idles behind a host round trip (the last one is drained by the final synchronize)
            prev = slot ^ 1
            if self.loss_events[prev] is not None:
                self.loss_events[prev].synchronize()
                return float(self.loss_slots[prev])
            return None
        else:
            xs = self.x_student_host.to(self.device, non_blocking=True).requires_grad_(True)
            xt = self.x_teacher_host.to(self.device, non_blocking=True)
            loss = self.run(xs, xt)
        self.loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.loss_host)


def timed(fn, steps, world, device):
    """barrier + sync, K steps between CUDA events on the current stream, sync + barrier; max over ranks."""
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms


def per_kernel_times(step, steps):
    """Per-op device time, measured live with CUDA events around every libdinomc call (ops.profile)."""
    ops = step.D.ops
    ops.profile_begin()
    for _ in range(steps):
        step.run()
    torch.cuda.synchronize()
    return ops.profile_end()


def cpu_baseline(w, sample_B, steps, warmup):
    """oracle/torch_port.py (the reference's eager path restated op for op) on the host cores.  `sample_B` = samples
    per step; by default the workload's own per-GPU batch, i.e. the same step the GPU arm runs."""
    from oracle import torch_port as T
    torch.set_num_threads(os.cpu_count() or 1)
    Din, K, C, G = w["D"], w["K"], w["C"], w["G"]
    sp = T.make_head_params(Din, K, seed=0)
    tp = {k: v.detach().clone() for k, v in sp.items()}
    g = torch.Generator().manual_seed(99)
    shapes = backbone_param_shapes(w["arch"])
    bb_s = [torch.randn(s, generator=g) * 0.02 for s in shapes]
    bb_t = [t.clone() for t in bb_s]
    st = T.LossState(K, C, 0.04, 0.04, 0, 100, teacher_crops_number=G)
    xs = torch.randn(C * sample_B, Din, generator=g)
    xt = torch.randn(G * sample_B, Din, generator=g)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        T.step(xs, xt, sp, tp, st, 0, 0.996, ema_extra=(bb_t, bb_s))
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    full = (sample_B == w["B"])
    return dict(value=sample_B / med, unit="samples/s", cores=torch.get_num_threads(), kind="port",
                sample=(f"{'the full step' if full else 'REDUCED batch'}: B={sample_B} samples per step (workload B={w['B']}; D={Din}, K={K}, "
                        f"{C} crops, {sum(t.numel() for t in bb_s) / 1e6 + sum(v.numel() for v in sp.values()) / 1e6:.1f}M-param EMA), "
                        f"{warmup} warm-up + {steps} timed steps, oracle/torch_port.py on torch CPU fp32, median"),
                ms_per_step=med * 1e3, sample_batch=sample_B)


_JSON_FD = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner on stdout) write to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32", "fp32_simt"])
    ap.add_argument("--out-dim", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--graph", type=int, default=1, help="replay the N=1 step from a CUDA graph (0 = eager launches)")
    ap.add_argument("--ddp", type=int, default=0, help="N>1: wrap the student head in torch DDP (eager) instead of GradAllReduce")
    ap.add_argument("--reserve-sms", type=int, default=16, help="N>1: SMs the backward GEMM grids leave to NCCL")
    ap.add_argument("--grad-compress", default="none", choices=["none", "bf16"],
                    help="N>1 gradient exchange: fp32 all-reduce (default; what DDP does in the reference), or bf16 (dW of the "
                         "last layer averaged before its weight-norm backward + the small gradients as one flat bf16 buffer). "
                         "bf16 measured SLOWER at 2 GPUs (1.068 vs 0.980 ms), so it stays opt-in")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1 gradient exchange transport: 'peer' = libdinomc's own all-reduce kernel over NVLink/NVSwitch symmetric "
                         "memory (bf16 exchange; falls back to 'nccl' if symmetric memory is unavailable), 'nccl' = torch.distributed")
    ap.add_argument("--overlap", type=int, default=1, help="teacher head forward on a side stream, overlapping the student's")
    ap.add_argument("--hiprio", type=int, default=0, help="run (and capture) the step on a high-priority stream: its kernels get SM "
                                                         "slots before the auxiliary-stream streaming kernels (experiment)")
    ap.add_argument("--cpu-sample-batch", type=int, default=0,
                    help="samples per CPU step for cpu_baseline / --impl reference (0 = the workload's own per-GPU batch)")
    ap.add_argument("--verify", type=int, default=1, help="N>1: after timing, check the exchanged gradients (bit-identical across "
                                                          "ranks, equal to the mean of the ranks' local gradients)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    w = dict(WORKLOADS[args.workload])
    if args.out_dim:
        w["K"] = args.out_dim
    if args.batch:
        w["B"] = args.batch
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    warmup = max(args.warmup, 3)
    compress = None if args.grad_compress == "none" else "bf16"
    if args.exchange == "peer" and args.mode == "bf16" and "--grad-compress" not in sys.argv:
        compress = "bf16"               # the peer transport exchanges bf16 buffers (gradients stay within the bf16-mode tolerance)
    if os.environ.get("DMC_BENCH_GRAD_COMPRESS"):                        # A/B runs under torchrun without changing the command line
        compress = None if os.environ["DMC_BENCH_GRAD_COMPRESS"] == "none" else "bf16"
    cfg = {"workload": f"{args.workload}: {w['note']}; D={w['D']} out_dim={w['K']} batch/GPU={w['B']} "
                       f"crops={w['G']}+{w['C'] - w['G']}; EMA over {w['arch']} backbone + head",
           "global_batch": w["B"] * world, "parallelism": f"dp{world}",
           "grad_allreduce": ("none (1 GPU)" if world == 1 else ("torch DDP" if args.ddp else
                                                                        "dinomc_b200.GradAllReduce (NCCL, side stream; "
                                                                        + ("bf16 exchange: last-layer dW averaged before its weight-norm backward, small gradients in one flat bf16 buffer)"
                                                                           if compress else "fp32)"))),
           "l2": "per-step working set (logits + gradients > 1 GiB) exceeds the 126 MB L2; no explicit flush"}

    cpu_B = args.cpu_sample_batch or w["B"]
    if args.impl == "reference":
        # The reference's own (eager PyTorch) path on this box's host cores: ONE host process with every core, at the
        # workload's per-GPU batch, whatever N the launch names (the other ranks exit without work) -- so the line says
        # n_gpus = 1 and global_batch = B: it is the same number at every N and must not be scaled by N.
        if rank != 0:
            return
        ref_warm = min(warmup, 3)
        cb = cpu_baseline(w, cpu_B, args.steps, ref_warm)
        cfg_ref = dict(cfg)
        cfg_ref.update({"global_batch": cpu_B, "parallelism": "host cores, one process",
                        "grad_allreduce": "none (one host process)", "launched_with_gpus": args.gpus})
        if cpu_B != w["B"]:
            cfg_ref["workload"] += f"; CPU arm runs a reduced batch of {cpu_B}"
        line = {"impl": "reference", "metric": "DINO head+loss+center+EMA step samples/sec", "value": cb["value"],
                "unit": "samples/s", "n_gpus": 1, "steps": args.steps, "warmup": ref_warm,
                "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg_ref,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "CPU arm (a GPU-over-CPU ratio against it says nothing about kernel quality; see roofline)"}
        _emit(line)
        return

    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device; there is no CPU fallback"
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    force_dp = bool(int(os.environ.get("DMC_BENCH_FORCE_DP", "0")))      # diagnostic: 1-rank NCCL group, reducer path on
    if force_dp and world == 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=device)
    if world > 1:
        opts = None
        try:        # NCCL kernels on a high-priority stream: they grab SMs as soon as compute CTAs retire
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:       # noqa: BLE001
            opts = None
        dist.init_process_group("nccl", device_id=device, pg_options=opts)
    import dinomc_b200 as D
    D._lib.check(D._lib.load().dmc_device_check(local_rank), "dmc_device_check")
    transport = "nccl"
    if world > 1 and args.exchange == "peer" and compress == "bf16" and not args.ddp:
        # probe symmetric memory + the exchange kernel on every rank; all ranks must agree before relying on it
        ok = 1
        try:
            from dinomc_b200.xrank import SymmetricBuffer
            probe = SymmetricBuffer(4096, torch.bfloat16)
            probe.tensor.fill_(1.0)
            probe.allreduce_(1.0 / world)
            torch.cuda.synchronize()
            ok = int(bool((probe.tensor.float() == 1.0).all()))
            multicast = probe.multicast
            del probe
        except Exception as e:          # noqa: BLE001
            print(f"[rank {rank}] symmetric-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
            ok, multicast = 0, False
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 1:
            transport = "peer"
            cfg["grad_allreduce"] = ("dinomc_b200.GradAllReduce over libdinomc's own all-reduce kernel (dmc_xrank_allreduce: symmetric memory, "
                                     + ("multimem.ld_reduce / multimem.st through NVSwitch" if multicast else "peer loads / stores")
                                     + "; bf16 exchange: last-layer dW averaged before its weight-norm backward, small gradients in one flat bf16 buffer)")
        elif "--grad-compress" not in sys.argv:
            compress = None

    D.set_teacher_overlap(bool(args.overlap))
    if (world > 1 or force_dp) and not args.ddp:
        D.set_async_center(True)
    step = Step(w, args.mode, rank, world, device, ddp=bool(args.ddp), reserve_sms=args.reserve_sms, force_reducer=force_dp,
                compress=compress, transport=transport)
    if transport == "peer":
        from dinomc_b200.xrank import SymmetricBuffer
        step.loss_mod.center_exchange = SymmetricBuffer(w["K"], torch.float32, ctas=16)      # the center exchange leaves NCCL too
        cfg["center_allreduce"] = "dmc_xrank_allreduce (fp32, 16 CTAs), asynchronous: consumed by the next step"
    ops = D.ops
    for _ in range(warmup):
        step.run()
    torch.cuda.synchronize()

    use_graph = bool(args.graph) and not (world > 1 and args.ddp)
    graph = None
    run_value = step.run
    if use_graph:
        # the whole step is stream-ordered libdinomc launches (+ NCCL all-reduces when N > 1) on fixed buffers:
        # capture once, replay K times
        try:
            graph = D.StepGraph(step.run, warmup=3, capture_error_mode="thread_local" if (world > 1 or force_dp) else "global",
                                stream=torch.cuda.Stream(priority=-1) if args.hiprio else None)
            run_value = graph.replay
        except Exception as e:          # noqa: BLE001 -- e.g. a collective that refuses capture: fall back to eager
            if world == 1:
                raise
            import traceback
            print(f"[rank {rank}] CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly\n"
                  + "".join(traceback.format_exception(type(e), e, e.__traceback__))[-6000:], file=sys.stderr)
            use_graph, graph = False, None
            torch.cuda.synchronize()
            from dinomc_b200.loss import drop_pending_events
            drop_pending_events()        # events recorded inside the failed capture are unusable

    l0 = ops.launch_count
    step.run()
    launches_per_step = ops.launch_count - l0
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = timed(run_value, args.steps, world, device)
    clocks = sampler.stop()
    ms_step = ms_total / args.steps
    value = w["B"] * world / (ms_step * 1e-3)

    # end-to-end: pinned host features in, loss out, through the public API (StepGraph replay when N == 1)
    e2e_fn = (lambda: step.run_e2e(graph)) if use_graph else step.run_e2e
    for _ in range(3):
        e2e_fn()
    ms_e2e = timed(e2e_fn, args.steps, world, device) / args.steps
    h2d = step.x_student_host.numel() * 4 + step.x_teacher_host.numel() * 4
    e2e = {"value": w["B"] * world / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
           "path": ("StepGraph.replay (public API); every step: H2D of its features from pinned memory (prefetched on a copy "
                    "stream during the previous step) and D2H of its loss; the host reads the loss of step i-1 while step i runs") if use_graph
                   else "eager module calls + per-step H2D/D2H"}

    # per-op device time for the roofline object: CUPTI kernel durations attributed to the libdinomc op that launched them
    # (ops.profile_kernels); CUDA-event pairs around the ops only if the profiler attributes nothing
    peaks = load_peaks()
    prof_steps = min(args.steps, 10)
    prof_us, timing = {}, "cupti"
    # Profiled with the auxiliary / side streams OFF: every kernel then runs alone on one stream, so its CUPTI duration is its
    # own (with the overlaps on, co-running kernels inflate each other's durations and the split would be meaningless).
    Fn = D.functional
    saved = (Fn.aux_overlap, D.head._teacher_overlap, D.head._early_teacher_stats)
    Fn.aux_overlap = False
    D.set_teacher_overlap(False)
    D.head.set_early_teacher_stats(False)
    # ... and WITHOUT programmatic dependent launch: a PDL kernel becomes resident while its predecessor still runs and sits in
    # griddepcontrol.wait, and CUPTI counts that wait as part of its duration (profiles/r02_timeline_1gpu.txt: the 3 us
    # center_update shows as 17 us, the last layer's dgrad as 91 us for 62 us of work)
    pdl_prev = D._lib.load().dmc_set_pdl(0)
    try:
        step.run()
        torch.cuda.synchronize()
        prof_us = ops.profile_kernels(step.run, prof_steps)                      # {op: (us per step, kernels per step)}
    except Exception as e:              # noqa: BLE001
        print(f"[rank {rank}] profile_kernels failed ({type(e).__name__}: {e}); using CUDA events", file=sys.stderr)
    finally:
        D._lib.load().dmc_set_pdl(pdl_prev)
        Fn.aux_overlap = saved[0]
        D.set_teacher_overlap(saved[1])
        D.head.set_early_teacher_stats(saved[2])
    if not prof_us:
        timing = "cuda_events"
        ev = per_kernel_times(step, prof_steps)
        prof_us = {k: (ms * 1e3 / prof_steps, n / prof_steps) for k, (ms, n) in ev.items()}
    logit_bytes = 2 if args.mode == "bf16" else 4
    flops, nbytes = roofline_model(w, step.P, logit_bytes)
    t_roof_ms = max(flops / (peaks["bf16_tflops"] * 1e12), nbytes / (peaks["hbm_gbs"] * 1e9)) * 1e3
    models = op_models(w, step.P, logit_bytes)
    traffic_db = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath):
        traffic_db = json.load(open(tpath)).get(args.workload + ":" + args.mode, {})
    kernels = []
    for name, (us, calls) in sorted(prof_us.items(), key=lambda kv: -kv[1][0]):
        k = {"kernel": name, "us_per_step": us, "launches_per_step": calls}
        if name in models:                    # algorithmic bytes / flops of ALL of that op's launches in one step
            b, f = models[name]
            t_hbm, t_tc = b / (peaks["hbm_gbs"] * 1e9), f / (peaks["bf16_tflops"] * 1e12)
            k["bound"] = "tensor" if t_tc > t_hbm else "hbm"
            k["alg_GB"], k["alg_GFLOP"] = b / 1e9, f / 1e9
            if k["bound"] == "hbm":
                k["achieved"], k["peak"], k["unit"] = b / 1e9 / (us * 1e-6), peaks["hbm_gbs"], "GB/s"
            else:
                k["achieved"], k["peak"], k["unit"] = f / 1e12 / (us * 1e-6), peaks["bf16_tflops"], "TFLOP/s"
            k["frac"] = k["achieved"] / k["peak"]
            k["traffic"] = traffic_db.get(name)
            if name == "xrank_allreduce":
                # its duration contains two cross-rank barriers, i.e. the wait for the slowest rank's gradients (in this eager,
                # profiler-instrumented pass: milliseconds of host skew) -- not a bandwidth figure, never the roofline kernel
                k["note"] = "duration includes waiting for the other ranks at its barriers; not a bandwidth measurement"
        kernels.append(k)
    dom = next((k for k in kernels if "frac" in k and "note" not in k), None)           # largest time share
    roofline = None
    if dom is not None:
        roofline = {"bound": dom["bound"], "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": dom["peak"],
                    "unit": dom["unit"], "frac": dom["frac"], "traffic": dom["traffic"], "peak_source": peaks["source"],
                    "timing": timing + " kernel durations (CUPTI activity records, kernels serialised on one stream: auxiliary-stream overlaps off "
                                       "and programmatic dependent launch off for this pass), all launches of the op per step; the timed `value` runs with both on",
                    "step": {"t_roof_ms": t_roof_ms, "alg_GB": nbytes / 1e9, "alg_GFLOP": flops / 1e9,
                             "frac_of_step_roofline": t_roof_ms / ms_step},
                    "kernels": kernels}

    verify = None
    if world > 1 and args.verify and step.reducer is not None:
        verify = step.verify_exchange()

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c = cpu_baseline(w, cpu_B, 3, 1)
        cb = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "DINO head+loss+center+EMA step samples/sec", "value": value, "unit": "samples/s",
                "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic", "config": cfg,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "launch_mode": "cuda_graph" if use_graph else "eager", "teacher_overlap": bool(args.overlap), "roofline": roofline, "cpu_baseline": cb,
                "ema_params": step.P, "ema_tensors": step.n_tensors}
        if verify is not None:
            line["exchange_check"] = verify
        _emit(line)
    if world > 1 or force_dp:
        # leave without tearing NCCL down: destroying communicators that CUDA graphs still reference can hang
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
