"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/optim_small.npz from the REAL reference's step utilities
(build container only: needs /root/reference):

  utils.LARS                (utils/utils.py:570-608)   4 steps, two parameter groups (utils.get_params_groups), scheduled
                                                        lr / weight decay as main_dino_mc.py:363-367 rewrites them
  utils.clip_gradients      (utils/utils.py:145-154)   one gradient set, two clip values

so that the restated oracle (oracle/torch_port.py) and the CUDA kernels (lars.cu, clip.cu) can be checked against the
reference's own outputs on the GPU box, where the reference itself is not available.

    python -m oracle.gen_golden_optim
"""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

from oracle import reference_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "tests", "golden", "optim_small.npz")
STEPS = 4


def make_model():
    """A weight larger than one 16384-element kernel chunk, a small weight, biases and a LayerNorm (1-D parameters)."""
    torch.manual_seed(21)
    return torch.nn.Sequential(torch.nn.Linear(130, 150), torch.nn.GELU(), torch.nn.Linear(150, 9), torch.nn.LayerNorm(9))


def schedule(it):
    return 0.3 * (1 + 0.5 * it), 1e-4 * (1 + it)          # lr, weight decay of group 0


def main():
    if not reference_loader.available():
        raise SystemExit("reference not available; fixtures can only be generated in the build container")
    _, _, utils = reference_loader.load()
    model = make_model()
    names = [n for n, _ in model.named_parameters()]
    blob = {"names": np.array(names), "steps": np.array(STEPS)}
    for n, p in model.named_parameters():
        blob["p0." + n] = p.detach().numpy().copy()
    opt = utils.LARS(utils.get_params_groups(model))
    g = torch.Generator().manual_seed(22)
    for it in range(STEPS):
        lr, wd = schedule(it)
        for gi, group in enumerate(opt.param_groups):
            group["lr"] = lr
            if gi == 0:
                group["weight_decay"] = wd
        for n, p in model.named_parameters():
            p.grad = torch.randn(p.shape, generator=g) * 0.1
            blob[f"g{it}." + n] = p.grad.numpy().copy()
        opt.step()
    for n, p in model.named_parameters():
        blob["lars.p." + n] = p.detach().numpy().copy()
        blob["lars.mu." + n] = opt.state[p]["mu"].numpy().copy()
    # clip_gradients on the step-0 gradients, clip values on both sides of the norms
    for clip in (3.0, 0.05):
        m2 = copy.deepcopy(model)
        for n, p in m2.named_parameters():
            p.grad = torch.from_numpy(blob["g0." + n].copy())
        norms = utils.clip_gradients(m2, clip)
        blob[f"clip{clip}.norms"] = np.array(norms, dtype=np.float64)
        for n, p in m2.named_parameters():
            blob[f"clip{clip}.g." + n] = p.grad.numpy().copy()
    np.savez_compressed(PATH, **blob)
    print(f"wrote {PATH} ({os.path.getsize(PATH) / 1e6:.2f} MB), {len(names)} parameters, {STEPS} LARS steps")


if __name__ == "__main__":
    main()
