"""CPU: host-side logic added in round 2 that needs no device -- which teacher tensors get operand shadows in the EMA plan,
the stand-in metric logger of the drop-in training step, the reference arm's bookkeeping in bench.py."""
import types

import pytest
import torch


def _teacher():
    import dinomc_b200 as D
    t = D.DINOHead(24, 256, nlayers=3, hidden_dim=32, bottleneck_dim=16)
    for p in t.parameters():
        p.requires_grad = False
    return t


def _fake_shadow(t):
    lins = t._linears()
    K, dim = t.last_layer.weight_v.shape
    t._shadow = dict(device=torch.device("cpu"), state=None, weights=[lin.weight for lin in lins],
                     mlp=[torch.empty(lin.weight.shape, dtype=torch.bfloat16) for lin in lins],
                     what=torch.empty((K, dim), dtype=torch.bfloat16), scale=torch.empty(K), inv_norm=torch.empty(K))


def test_shadow_spec_maps_head_tensors_to_their_position_in_the_ema_list():
    from dinomc_b200 import ema
    t = _teacher()
    _fake_shadow(t)
    ema.register_shadow_head(t)
    backbone = [torch.zeros(7), torch.zeros(3, 3)]                    # EMA list = backbone parameters first, then the head
    tp = backbone + [p.data for p in t.parameters()]
    shadows, wn, heads = ema._shadow_spec(tp)
    assert heads == [t]
    names = [n for n, _ in t.named_parameters()]
    off = len(backbone)
    assert sorted(shadows) == [off + names.index(n) for n in ("mlp.0.weight", "mlp.2.weight", "mlp.4.weight")]
    assert wn[0] == off + names.index("last_layer.weight_v") and wn[1] == off + names.index("last_layer.weight_g")
    assert wn[2] == 16 and wn[3] is t._shadow["what"]
    # a list that lacks part of the head gets no shadows at all (the plan must never write a half-updated operand set)
    shadows, wn, heads = ema._shadow_spec(tp[:-1])
    assert shadows == {} and wn is None and heads == []


def test_shadow_freshness_follows_parameter_versions():
    t = _teacher()
    _fake_shadow(t)
    t._mark_shadow_fresh()
    assert t._shadow["state"] == t._shadow_state()
    with torch.no_grad():
        t.last_layer.weight_v.mul_(2.0)                               # any foreign write bumps _version
    assert t._shadow["state"] != t._shadow_state()
    x = torch.zeros(2, 24)
    assert t._is_inference(x)                                         # frozen parameters, input without gradient: no no_grad needed
    assert not t._is_inference(x.requires_grad_(True))
    with torch.no_grad():
        assert t._is_inference(x)


def test_plain_meters_and_train_loop_argument_checks():
    from dinomc_b200 import dropin
    m = dropin._PlainMeters()
    m.update(loss=1.0, lr=0.1)
    m.update(loss=3.0)
    assert m.meters["loss"].global_avg == 2.0 and m.meters["lr"].global_avg == pytest.approx(0.1)
    assert list(m.log_every([1, 2, 3], 10, "x")) == [1, 2, 3]
    assert "loss" in str(m)
    with pytest.raises(ValueError, match="host_sync"):
        dropin.train_one_epoch(None, None, None, None, [], None, None, None, None, 0, None, types.SimpleNamespace(epochs=1),
                               host_sync="sometimes")


def test_bench_op_models_cover_the_profiled_ops_and_the_step_model():
    import bench
    w = bench.WORKLOADS["cfg2"]
    models = bench.op_models(w, 44_020_000, 2)
    for tag in ("ce_fused", "ema", "gemm_last_fwd_student", "gemm_last_wgrad", "gemm_last_dgrad", "gemm_mlp_fwd_384x2048",
                "gemm_mlp_fwd_2048x2048", "gemm_mlp_dgrad_2048x256", "gemm_mlp_wgrad_2048x2048", "weightnorm_fwd", "weightnorm_bwd",
                "teacher_stats_colsum", "cast_bf16", "colsum", "xrank_allreduce"):
        b, f = models[tag]
        assert b > 0 and f >= 0, tag
    flops, nbytes = bench.roofline_model(w, 44_020_000, 2)
    assert abs(flops / 1e9 - 296.6) < 0.5 and abs(nbytes / 2 ** 30 - 2.88) < 0.01         # SURVEY 8d worked numbers for cfg2
    # the GEMM ops of the step add up to the step model's flops
    gemm = sum(models[t][1] for t in models if t.startswith("gemm_"))
    assert abs(gemm - flops) / flops < 1e-9


def test_bench_op_models_add_up_when_two_linears_share_a_shape():
    """D == H (ResNet-50 features, cfg3): the first two Linears carry the same tag; their models must add, and the GEMM ops
    must still account for every flop of the step model."""
    import bench
    w = bench.WORKLOADS["cfg3"]
    assert w["D"] == bench.H
    models = bench.op_models(w, 25_000_000, 2)
    Ns, Nt = w["C"] * w["B"], w["G"] * w["B"]
    assert models["gemm_mlp_fwd_2048x2048"][1] == 2 * (2 * (Ns + Nt) * bench.H * bench.H)
    assert models["gemm_mlp_wgrad_2048x2048"][1] == 2 * (2 * Ns * bench.H * bench.H)
    flops, _ = bench.roofline_model(w, 25_000_000, 2)
    assert abs(sum(models[t][1] for t in models if t.startswith("gemm_")) - flops) / flops < 1e-9


def test_head_loss_binding_is_explicit_weak_and_not_state():
    """DINOHead.bind_loss: an explicit binding wins over the process default (the most recently constructed / used DINOLoss), is
    a weak reference, is not part of the state_dict, and does not travel with copies or pickles of the head."""
    import copy
    import gc
    import pickle
    import dinomc_b200 as D
    from dinomc_b200 import functional as Fn
    head = D.DINOHead(16, 64, hidden_dim=32, bottleneck_dim=16)
    keys = set(head.state_dict())
    first = D.DINOLoss(64, 4, 0.04, 0.04, 0, 10)
    second = D.DINOLoss(64, 4, 0.04, 0.04, 0, 10)
    assert Fn._current_loss() is second and head._loss_module() is second          # unbound: the default
    assert head.bind_loss(first) is head and head._loss_module() is first          # bound: explicit wins
    assert set(head.state_dict()) == keys
    assert copy.deepcopy(head)._loss_module() is second                             # a copy is unbound
    assert pickle.loads(pickle.dumps(head))._loss_module() is second
    assert head._loss_module() is first                                             # ... and the original still is
    del first
    gc.collect()
    assert head._loss_module() is second                                            # weak: a dead binding falls back
    head.bind_loss(None)
    assert head._loss_ref is None


def _run_bench(args, env_extra=None):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(root, "bench.py")] + args, capture_output=True, text=True, env=env, timeout=600)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference`: ONE JSON line on stdout with the contract's keys, describing what it ran (one host process,
    n_gpus 1, its own batch), whatever --gpus says; ranks other than 0 of a torchrun launch print nothing and exit 0."""
    import json
    r = _run_bench(["--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "1", "--cpu-sample-batch", "2", "--gpus", "8"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["gpu_launches"] == 0 and d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["global_batch"] == 2 and "reduced batch of 2" in d["config"]["workload"] and d["config"]["launched_with_gpus"] == 8
    assert abs(d["value"] - 2 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    other = _run_bench(["--impl", "reference", "--workload", "cfg1", "--steps", "1", "--gpus", "8"], {"RANK": "3", "LOCAL_RANK": "3", "WORLD_SIZE": "8"})
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_bench_own_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm stops with an error instead of timing something else."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is visible")
    r = _run_bench(["--steps", "1", "--no-cpu-baseline"])
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr
