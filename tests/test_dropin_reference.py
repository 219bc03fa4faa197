"""CPU, build container only: the drop-in launcher patches the real reference modules."""
import pytest

from oracle import reference_loader

pytestmark = pytest.mark.reference


def test_install_rebinds_reference_symbols():
    if not reference_loader.available():
        pytest.skip("reference not present")
    main_dino_mc, vits, _ = reference_loader.load()
    orig = (vits.DINOHead, main_dino_mc.DINOLoss, main_dino_mc.train_one_epoch)
    try:
        import dinomc_b200
        from dinomc_b200 import dropin
        dropin.install()
        assert vits.DINOHead is dinomc_b200.DINOHead
        assert main_dino_mc.DINOLoss is dinomc_b200.DINOLoss
        assert main_dino_mc.train_one_epoch is dropin.train_one_epoch
        # the constructor calls main_dino_mc.py makes (positional teacher form at :243-246) work unchanged
        h = vits.DINOHead(384, 1024, False)
        assert h.last_layer.weight_v.shape == (1024, 256)
        l = main_dino_mc.DINOLoss(1024, 8, 0.04, 0.04, 0, 10, teacher_crops_number=2)
        assert l.center.shape == (1, 1024)
    finally:
        vits.DINOHead, main_dino_mc.DINOLoss, main_dino_mc.train_one_epoch = orig


def test_oracle_clip_gradients_equals_reference_function():
    """The restated clip_gradients (oracle/torch_port.py) against the reference's own utils.clip_gradients."""
    if not reference_loader.available():
        pytest.skip("reference not present")
    import torch
    from oracle import torch_port
    _, _, utils = reference_loader.load()
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.GELU(), torch.nn.Linear(64, 5))
    for clip in (3.0, 0.05):
        for p in model.parameters():
            p.grad = torch.randn_like(p) * 0.3
        grads = [p.grad.detach().clone() for p in model.parameters()]
        ref_norms = utils.clip_gradients(model, clip)
        norms = torch_port.clip_gradients(grads, clip)
        assert norms == ref_norms
        for p, g in zip(model.parameters(), grads):
            assert torch.equal(p.grad, g)
