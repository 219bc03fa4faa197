"""Drop-in launcher: make the reference's own trainer (`main_dino_mc.py`) run on the libdinomc kernels.

    import dinomc_b200.dropin as dropin
    dropin.install()                      # before main_dino_mc.train_dino(args)

`install()` rebinds these names inside the reference's modules:

  utils.vision_transformer.DINOHead  -> dinomc_b200.DINOHead   (constructed at main_dino_mc.py:236-246)
  main_dino_mc.DINOLoss              -> dinomc_b200.DINOLoss   (constructed at main_dino_mc.py:269-277)
  main_dino_mc.train_one_epoch       -> train_one_epoch below  (the EMA loop at :403-406 is inline in the
                                         reference and has no seam of its own; this is the same step with the
                                         per-parameter loop replaced by one ema_update_ call)

  utils.MultiCropWrapper             -> dinomc_b200.MultiCropWrapper   (utils/utils.py:611-646; one feature concatenation)
  utils.LARS                         -> dinomc_b200.FusedLARS  (constructed at main_dino_mc.py:286 for `--optimizer lars`)

Everything else -- argument parsing, data loading, backbones, MultiCropWrapper, DDP, the optimizer, gradient
clipping, checkpointing, logging -- stays the reference's code, looked up at run time.
"""
from __future__ import annotations

import math
import sys

import torch

from .ema import ema_update_
from .head import DINOHead
from .loss import DINOLoss
from .optim import FusedLARS, cancel_gradients_last_layer, clip_gradients
from .wrapper import MultiCropWrapper


class _PlainMeters:
    """Minimal stand-in for the reference's utils.MetricLogger (utils/utils.py:287-363) for callers that run
    `train_one_epoch` without the reference importable: same four entry points, running averages only."""

    class _Avg:
        def __init__(self):
            self.total, self.count = 0.0, 0

        def update(self, v):
            self.total += float(v)
            self.count += 1

        @property
        def global_avg(self):
            return self.total / max(self.count, 1)

    def __init__(self):
        self.meters = {}

    def update(self, **kw):
        for k, v in kw.items():
            self.meters.setdefault(k, self._Avg()).update(v)

    def log_every(self, iterable, print_freq, header=None):
        yield from iterable

    def synchronize_between_processes(self):
        pass

    def __str__(self):
        return "  ".join(f"{k}: {m.global_avg:.6f}" for k, m in self.meters.items())


def train_one_epoch(student, teacher, teacher_without_ddp, dino_loss, data_loader, optimizer, lr_schedule,
                    wd_schedule, momentum_schedule, epoch, fp16_scaler, args, *, meters=None, host_sync="reference"):
    """Same contract as main_dino_mc.train_one_epoch (main_dino_mc.py:356-416).

    `meters`: the metric logger (default: the reference's `utils.MetricLogger`; `_PlainMeters()` when the reference is not
    importable, e.g. in tests).  `host_sync`:
      "reference"  the reference's behaviour (default): `loss.item()` every step before backward (:378-380) and
                   `torch.cuda.synchronize()` after the EMA (:409);
      "deferred"   no per-step device synchronisation: the loss of step i is copied to pinned host memory asynchronously
                   and checked / logged while step i+1 is queued (a non-finite loss is detected one step later); the
                   final step is drained at the end of the epoch.
    """
    if host_sync not in ("reference", "deferred"):
        raise ValueError("host_sync must be 'reference' or 'deferred'")
    if meters is None:
        ref = sys.modules.get("main_dino_mc")
        meters = ref.utils.MetricLogger(delimiter="  ") if ref is not None else _PlainMeters()
    header = "Epoch: [{}/{}]".format(epoch, args.epochs)
    steps_per_epoch = len(data_loader)
    student_params = list((student.module if hasattr(student, "module") else student).parameters())
    teacher_params = list(teacher_without_ddp.parameters())
    pending = None                        # deferred mode: (pinned loss slot, event) of the previous step
    slots = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)] if host_sync == "deferred" else None

    def drain(p):
        slot, ev = p
        ev.synchronize()
        v = float(slot)
        if not math.isfinite(v):
            print("Loss is {}, stopping training".format(v))
            sys.exit(1)
        meters.update(loss=v)

    for i, (images, _) in enumerate(meters.log_every(data_loader, 10, header)):
        it = steps_per_epoch * epoch + i
        for gi, group in enumerate(optimizer.param_groups):
            group["lr"] = lr_schedule[it]
            if gi == 0:
                group["weight_decay"] = wd_schedule[it]
        images = [im.cuda(non_blocking=True) for im in images]
        with torch.autocast("cuda", dtype=torch.float16, enabled=fp16_scaler is not None):
            teacher_output = teacher(images[:args.global_crops_number])
            student_output = student(images)
            loss = dino_loss(student_output, teacher_output, epoch)
        if host_sync == "reference":
            loss_value = loss.item()
            if not math.isfinite(loss_value):
                print("Loss is {}, stopping training".format(loss_value))
                sys.exit(1)
        else:
            slot = slots[i & 1]
            slot.copy_(loss.detach(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                drain(pending)            # step i-1: its loss arrived while this step was being queued
            pending = (slot, ev)
        optimizer.zero_grad()
        if fp16_scaler is None:
            loss.backward()
            if args.clip_grad:
                clip_gradients(student, args.clip_grad)           # two multi-tensor launches, no per-parameter .item()
            cancel_gradients_last_layer(epoch, student, args.freeze_last_layer)
            optimizer.step()
        else:
            fp16_scaler.scale(loss).backward()
            if args.clip_grad:
                fp16_scaler.unscale_(optimizer)
                clip_gradients(student, args.clip_grad)
            cancel_gradients_last_layer(epoch, student, args.freeze_last_layer)
            fp16_scaler.step(optimizer)
            fp16_scaler.update()
        ema_update_(teacher_params, student_params, momentum_schedule[it])      # one launch instead of 3 x #tensors
        if host_sync == "reference":
            torch.cuda.synchronize()
            meters.update(loss=loss_value)
        meters.update(lr=optimizer.param_groups[0]["lr"])
        meters.update(wd=optimizer.param_groups[0]["weight_decay"])
    if pending is not None:
        drain(pending)
    meters.synchronize_between_processes()
    print("Averaged stats:", meters)
    return {k: meter.global_avg for k, meter in meters.meters.items()}


def install(patch_train_loop: bool = True):
    """Patch the already-importable reference modules (`main_dino_mc`, `utils.vision_transformer`)."""
    import main_dino_mc
    import utils.vision_transformer as vits
    vits.DINOHead = DINOHead
    main_dino_mc.DINOLoss = DINOLoss
    main_dino_mc.utils.MultiCropWrapper = MultiCropWrapper     # same contract, single feature concatenation
    main_dino_mc.utils.LARS = FusedLARS                        # `--optimizer lars` (main_dino_mc.py:285-286): same signature / state
    if patch_train_loop:
        main_dino_mc.train_one_epoch = train_one_epoch
    return main_dino_mc
