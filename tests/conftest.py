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


class OptimGolden:
    """tests/golden/optim_small.npz: the reference's own utils.LARS / utils.clip_gradients on a small model
    (oracle/gen_golden_optim.py)."""

    def __init__(self):
        import torch
        z = np.load(os.path.join(GOLDEN_DIR, "optim_small.npz"))
        self.z = z
        self.names = z["names"].tolist()
        self.steps = int(z["steps"])
        self.model = torch.nn.Sequential(torch.nn.Linear(130, 150), torch.nn.GELU(), torch.nn.Linear(150, 9), torch.nn.LayerNorm(9))
        assert [n for n, _ in self.model.named_parameters()] == self.names
        with torch.no_grad():
            for n, p in self.model.named_parameters():
                p.copy_(torch.from_numpy(z["p0." + n]))

    @staticmethod
    def schedule(it):
        return 0.3 * (1 + 0.5 * it), 1e-4 * (1 + it)

    def grads(self, it):
        import torch
        return [torch.from_numpy(self.z[f"g{it}." + n].copy()) for n in self.names]

    def groups(self, model):
        """utils/utils.py:649-660 get_params_groups -> index lists (regularized, not regularized)."""
        reg = [i for i, (n, p) in enumerate(model.named_parameters()) if not (n.endswith(".bias") or len(p.shape) == 1)]
        return reg, [i for i in range(len(self.names)) if i not in reg]


@pytest.fixture
def optim_golden():
    return OptimGolden()
