"""CPU, build container only: the drop-in launcher patches the real reference modules."""
import pytest

from oracle import reference_loader

pytestmark = pytest.mark.reference


def test_install_rebinds_reference_symbols():
    if not reference_loader.available():
        pytest.skip("reference not present")
    main_dino_mc, vits, _ = reference_loader.load()
    orig = (vits.DINOHead, main_dino_mc.DINOLoss, main_dino_mc.train_one_epoch)
    orig_wrapper = main_dino_mc.utils.MultiCropWrapper
    orig_lars = main_dino_mc.utils.LARS
    try:
        import dinomc_b200
        from dinomc_b200 import dropin
        dropin.install()
        assert main_dino_mc.utils.LARS is dinomc_b200.FusedLARS
        assert vits.DINOHead is dinomc_b200.DINOHead
        assert main_dino_mc.DINOLoss is dinomc_b200.DINOLoss
        assert main_dino_mc.train_one_epoch is dropin.train_one_epoch
        assert main_dino_mc.utils.MultiCropWrapper is dinomc_b200.MultiCropWrapper
        # the constructor calls main_dino_mc.py makes (positional teacher form at :243-246) work unchanged
        h = vits.DINOHead(384, 1024, False)
        assert h.last_layer.weight_v.shape == (1024, 256)
        l = main_dino_mc.DINOLoss(1024, 8, 0.04, 0.04, 0, 10, teacher_crops_number=2)
        assert l.center.shape == (1, 1024)
    finally:
        vits.DINOHead, main_dino_mc.DINOLoss, main_dino_mc.train_one_epoch = orig
        main_dino_mc.utils.MultiCropWrapper = orig_wrapper
        main_dino_mc.utils.LARS = orig_lars


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


def test_oracle_lars_equals_reference_optimizer():
    """The restated LARS step (oracle/torch_port.py) against the reference's own utils.LARS (utils/utils.py:570-608) under
    the reference's usage: two parameter groups, lr / weight decay rewritten per iteration; bit-identical on CPU."""
    if not reference_loader.available():
        pytest.skip("reference not present")
    import copy
    import torch
    from oracle import torch_port
    _, _, utils = reference_loader.load()
    torch.manual_seed(5)
    ma = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.GELU(), torch.nn.Linear(64, 5), torch.nn.LayerNorm(5))
    mb = copy.deepcopy(ma)
    opt = utils.LARS(utils.get_params_groups(ma))
    mus = [torch.zeros_like(p) for p in mb.parameters()]
    for it in range(4):
        lr, wd = 0.3 * (1 + it), 1e-4 * (1 + it)
        for gi, group in enumerate(opt.param_groups):
            group["lr"] = lr
            if gi == 0:
                group["weight_decay"] = wd
        grads = []
        for pa in ma.parameters():
            pa.grad = torch.randn_like(pa) * 0.2
            grads.append(pa.grad.clone())
        opt.step()
        torch_port.lars_step([p.data for p in mb.parameters()], grads, mus, lr, wd)
        for pa, pb, mu in zip(ma.parameters(), mb.parameters(), mus):
            assert torch.equal(pa, pb)
            assert torch.equal(opt.state[pa]["mu"], mu)


def test_reference_checkpoint_layout_loads_into_dropin_modules():
    """a checkpoint written by the reference (main_dino_mc.py:333-345: `student` / `teacher` are
    state dicts of MultiCropWrapper(backbone, DINOHead), `dino_loss` holds the center) loads into the same wrapper built
    around dinomc_b200.DINOHead / DINOLoss with strict=True, and back."""
    if not reference_loader.available():
        pytest.skip("reference not present")
    import torch
    import dinomc_b200
    main_dino_mc, vits, utils = reference_loader.load()
    reference_loader.ensure_process_group()
    torch.manual_seed(0)

    def backbone():
        m = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 8 * 8, 48))
        m.fc, m.head = torch.nn.Identity(), torch.nn.Identity()
        return m

    ref_student = utils.MultiCropWrapper(backbone(), vits.DINOHead(48, 640, use_bn=False, norm_last_layer=True))
    ref_loss = main_dino_mc.DINOLoss(640, 8, 0.04, 0.04, 0, 10, teacher_crops_number=2)
    ref_loss.center.normal_()
    ckpt = {"student": {"module." + k: v for k, v in ref_student.state_dict().items()},     # DDP prefix (main_dino_mc.py:260)
            "teacher": ref_student.state_dict(), "dino_loss": ref_loss.state_dict()}

    ours = utils.MultiCropWrapper(backbone(), dinomc_b200.DINOHead(48, 640, use_bn=False, norm_last_layer=True))
    our_loss = dinomc_b200.DINOLoss(640, 8, 0.04, 0.04, 0, 10, teacher_crops_number=2)
    missing = ours.load_state_dict({k[len("module."):]: v for k, v in ckpt["student"].items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    our_loss.load_state_dict(ckpt["dino_loss"], strict=True)
    assert torch.equal(our_loss.center, ref_loss.center)
    for (ka, va), (kb, vb) in zip(ours.state_dict().items(), ref_student.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    # and the other way round: what we save, the reference modules load
    ref2 = utils.MultiCropWrapper(backbone(), vits.DINOHead(48, 640, use_bn=False, norm_last_layer=True))
    ref2.load_state_dict(ours.state_dict(), strict=True)
    main_dino_mc.DINOLoss(640, 8, 0.04, 0.04, 0, 10, teacher_crops_number=2).load_state_dict(our_loss.state_dict(), strict=True)


def test_multicrop_wrapper_matches_reference_wrapper():
    """dinomc_b200.MultiCropWrapper == utils.utils.MultiCropWrapper (utils/utils.py:611-646) on DINO-MC's multi-sized crop
    lists: same grouping, same crop-major feature order, same state_dict keys."""
    if not reference_loader.available():
        pytest.skip("reference not present")
    import copy
    import torch
    import dinomc_b200
    _, _, utils = reference_loader.load()
    torch.manual_seed(0)

    class Backbone(torch.nn.Module):          # resolution-agnostic toy backbone
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 24, 3, padding=1)
            self.fc, self.head = torch.nn.Linear(24, 10), torch.nn.Linear(24, 10)

        def forward(self, x):
            return self.head(self.fc(self.conv(x).mean(dim=(2, 3))))

    head = torch.nn.Linear(24, 7)
    ref = utils.MultiCropWrapper(Backbone(), copy.deepcopy(head))
    ours = dinomc_b200.MultiCropWrapper(Backbone(), copy.deepcopy(head))
    ours.load_state_dict(ref.state_dict(), strict=True)
    assert list(ours.state_dict()) == list(ref.state_dict())
    B = 3
    for sizes in ([32, 32, 16, 16, 16, 12, 12, 16], [32, 32], [16], [32, 32, 24, 20, 16, 12, 8, 8]):
        crops = [torch.randn(B, 3, s, s) for s in sizes]
        a, b = ref(list(crops)), ours(list(crops))
        assert a.shape == (len(sizes) * B, 7)
        assert torch.allclose(a, b, atol=1e-6)
    assert torch.equal(ref(crops[0]), ours(crops[0]))            # a single tensor instead of a list


def test_restart_from_checkpoint_restores_dropin_modules(tmp_path):
    """The reference's own resume path (utils/utils.py:165-197, called at main_dino_mc.py:310-319) against a checkpoint with
    the reference's layout (main_dino_mc.py:333-343, incl. the `fp16_scaler` entry :341-342 and a DDP-prefixed student):
    the drop-in MultiCropWrapper / DINOHead / DINOLoss / FusedAdamW objects are restored through `restart_from_checkpoint`
    exactly like the reference's own objects."""
    if not reference_loader.available():
        pytest.skip("reference not present")
    import argparse
    import torch
    import dinomc_b200
    main_dino_mc, vits, utils = reference_loader.load()
    torch.manual_seed(1)

    def backbone():
        m = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(12, 48))
        m.fc, m.head = torch.nn.Identity(), torch.nn.Identity()
        return m

    # a checkpoint written by the REFERENCE's objects (student saved under DDP's `module.` prefix, like :334)
    ref_student = utils.MultiCropWrapper(backbone(), vits.DINOHead(48, 640, use_bn=False, norm_last_layer=True))
    ref_teacher = utils.MultiCropWrapper(backbone(), vits.DINOHead(48, 640, False))
    ref_loss = main_dino_mc.DINOLoss(640, 8, 0.04, 0.04, 0, 10, teacher_crops_number=2)
    ref_loss.center.normal_()
    ref_opt = torch.optim.AdamW(utils.get_params_groups(ref_student))
    for p in ref_student.parameters():
        if p.requires_grad:
            p.grad = torch.randn_like(p) * 0.01
    ref_opt.step()
    scaler_sd = {"scale": 4096.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000, "_growth_tracker": 7}
    ckpt = {"student": {"module." + k: v for k, v in ref_student.state_dict().items()},
            "teacher": ref_teacher.state_dict(), "optimizer": ref_opt.state_dict(), "epoch": 3,
            "args": argparse.Namespace(arch="vit_small"), "dino_loss": ref_loss.state_dict(), "fp16_scaler": scaler_sd}
    path = tmp_path / "checkpoint.pth"
    torch.save(ckpt, path)

    class DDPLike(torch.nn.Module):             # what `student` is at :260 -- parameters live under `module.`
        def __init__(self, m):
            super().__init__()
            self.module = m

    student = DDPLike(dinomc_b200.MultiCropWrapper(backbone(), dinomc_b200.DINOHead(48, 640, use_bn=False, norm_last_layer=True)))
    teacher = dinomc_b200.MultiCropWrapper(backbone(), dinomc_b200.DINOHead(48, 640, False))
    loss = dinomc_b200.DINOLoss(640, 8, 0.04, 0.04, 0, 10, teacher_crops_number=2)
    opt = dinomc_b200.FusedAdamW(utils.get_params_groups(student))

    class ScalerLike:                            # torch's GradScaler.load_state_dict refuses when CUDA is absent (disabled scaler)
        def __init__(self):
            self.sd = None

        def load_state_dict(self, sd):
            self.sd = dict(sd)

    fp16 = ScalerLike()
    to_restore = {"epoch": 0}
    orig_load = torch.load
    torch.load = lambda *a, **k: orig_load(*a, **{**k, "weights_only": False})     # the reference's torch 2.5 default
    try:
        utils.restart_from_checkpoint(str(path), run_variables=to_restore, student=student, teacher=teacher, optimizer=opt,
                                      fp16_scaler=fp16, dino_loss=loss)
    finally:
        torch.load = orig_load
    assert to_restore["epoch"] == 3
    for k, v in ref_student.state_dict().items():
        assert torch.equal(student.module.state_dict()[k], v), k
    for k, v in ref_teacher.state_dict().items():
        assert torch.equal(teacher.state_dict()[k], v), k
    assert torch.equal(loss.center, ref_loss.center)
    assert fp16.sd == scaler_sd
    # optimizer state: same per-parameter moments as torch.optim.AdamW saved
    ref_state = ref_opt.state_dict()["state"]
    got_state = opt.state_dict()["state"]
    assert set(ref_state) == set(got_state)
    for i in ref_state:
        assert torch.equal(ref_state[i]["exp_avg"], got_state[i]["exp_avg"])
        assert torch.equal(ref_state[i]["exp_avg_sq"], got_state[i]["exp_avg_sq"])
        assert float(ref_state[i]["step"]) == float(got_state[i]["step"])
    # and back: a checkpoint written from the drop-in objects loads into the reference's
    ref2 = utils.MultiCropWrapper(backbone(), vits.DINOHead(48, 640, use_bn=False, norm_last_layer=True))
    ref2.load_state_dict(student.module.state_dict(), strict=True)
