"""pytest configuration: registers the `gpu` marker and shared fixtures/helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["mc_small", "tp_small", "one_layer", "mc_wide"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible, so a bare `pytest tests/` works here."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """One tests/golden/<name>.npz: inputs + outputs of the real reference modules."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.cfg = {k: v for k, v in zip(z["cfg_keys"].tolist(), z["cfg_vals"].tolist())}
        for k in ("in_dim", "out_dim", "hidden_dim", "bottleneck_dim", "nlayers", "B", "ncrops", "G",
                  "warmup_epochs", "nepochs", "epoch"):
            self.cfg[k] = int(self.cfg[k])
        self.cfg["norm_last_layer"] = bool(self.cfg["norm_last_layer"])
        self.inputs = {k[3:]: z[k] for k in z.files if k.startswith("in.")}
        self.ref32 = {k[6:]: z[k] for k in z.files if k.startswith("ref32.")}
        self.ref64 = {k[6:]: z[k] for k in z.files if k.startswith("ref64.")}

    def sd(self, who):
        pre = who + "."
        return {k[len(pre):]: v for k, v in self.inputs.items() if k.startswith(pre)}


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return Golden(request.param)


def rel_err(a, b):
    """max |a-b| / max |b|  -- the relative error used for every tolerance in this suite."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / denom
