"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- import the REAL reference modules.

Only usable where /root/reference exists (the build container); never on the GPU box.
Used by oracle/gen_golden.py (fixture generation) and by the container-only tests that
compare the restated oracles with the reference itself.

The reference cannot be imported as is: `main_dino_mc.py:33` imports
`data_process/dino_dataset.py`, whose line 8 imports `rasterio` (not installed).  A stub module
satisfies the import; no rasterio symbol is touched by the hot path.  `DINOLoss.update_center`
(`main_dino_mc.py:469`) calls `dist.all_reduce` unconditionally, so a world-size-1 gloo group is
created on demand.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("DINOMC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "main_dino_mc.py"))


def load():
    """Returns (main_dino_mc module, utils.vision_transformer module, utils.utils module)."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.modules.setdefault("rasterio", types.ModuleType("rasterio"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import main_dino_mc  # noqa: E402
        import utils.vision_transformer as vits  # noqa: E402
        import utils.utils as ref_utils  # noqa: E402
    return main_dino_mc, vits, ref_utils


def ensure_process_group():
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29581")
        dist.init_process_group("gloo", rank=0, world_size=1)
