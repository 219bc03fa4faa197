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


def train_one_epoch(student, teacher, teacher_without_ddp, dino_loss, data_loader, optimizer, lr_schedule,
                    wd_schedule, momentum_schedule, epoch, fp16_scaler, args):
    """Same contract as main_dino_mc.train_one_epoch (main_dino_mc.py:356-416)."""
    ref = sys.modules["main_dino_mc"]
    utils = ref.utils
    meters = utils.MetricLogger(delimiter="  ")
    header = "Epoch: [{}/{}]".format(epoch, args.epochs)
    steps_per_epoch = len(data_loader)
    student_params = list(student.module.parameters())
    teacher_params = list(teacher_without_ddp.parameters())
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
        loss_value = loss.item()
        if not math.isfinite(loss_value):
            print("Loss is {}, stopping training".format(loss_value), force=True)
            sys.exit(1)
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
        torch.cuda.synchronize()
        meters.update(loss=loss_value)
        meters.update(lr=optimizer.param_groups[0]["lr"])
        meters.update(wd=optimizer.param_groups[0]["weight_decay"])
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
