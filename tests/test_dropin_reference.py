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
