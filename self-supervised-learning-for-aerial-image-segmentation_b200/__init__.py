"""dinomc_b200 -- B200-native (sm_100a) implementation of DINO-MC's per-step hot path:
DINOHead -> DINOLoss -> teacher-center update -> EMA teacher update.

Import it as `dinomc_b200` (the alias package at the repo root; this directory's name is not a valid
Python identifier).  Public surface mirrors the reference:

    DINOHead(in_dim, out_dim, use_bn=False, norm_last_layer=True, nlayers=3, hidden_dim=2048, bottleneck_dim=256)
    DINOLoss(out_dim, ncrops, warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs, nepochs,
             teacher_crops_number=2, student_temp=0.1, center_momentum=0.9)
    ema_update_(teacher_params, student_params, m)
    clip_gradients(model, clip), cancel_gradients_last_layer(epoch, model, freeze_last_layer)   # utils/utils.py:145-162
    FusedAdamW(params_groups)                                 # torch.optim.AdamW of main_dino_mc.py:282, one launch per group
    FusedLARS(params_groups)                                  # utils.LARS of main_dino_mc.py:286, two launches per group
    StepGraph(fn)      # capture a whole step into a CUDA graph and replay it
    dropin.install()   # patch the reference's modules in place

All compute goes through libdinomc.so (hand-written CUDA behind the C ABI in include/dinomc.h).
There is no CPU fallback and no alternative backend: if the library is missing, importing the
kernels raises.
"""
from . import _lib, functional, ops  # noqa: F401
from .ema import ema_update_  # noqa: F401
from .graph import StepGraph  # noqa: F401
from .head import (DINOHead, get_default_precision, set_default_precision, set_operand_shadows,  # noqa: F401
                   set_teacher_overlap, wait_ready)
from .loss import DINOLoss, set_async_center  # noqa: F401
from .optim import FusedAdamW, FusedLARS, cancel_gradients_last_layer, clip_gradients  # noqa: F401
from .reducer import GradAllReduce  # noqa: F401
from .wrapper import MultiCropWrapper  # noqa: F401

__all__ = ["DINOHead", "DINOLoss", "ema_update_", "clip_gradients", "cancel_gradients_last_layer", "FusedAdamW", "FusedLARS", "MultiCropWrapper", "StepGraph", "GradAllReduce", "set_teacher_overlap", "set_operand_shadows", "set_async_center", "wait_ready", "set_default_precision", "get_default_precision",
           "ops", "functional"]
