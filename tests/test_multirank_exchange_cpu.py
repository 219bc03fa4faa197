"""CPU, gloo, world_size 2: the HOST logic of the data-parallel gradient exchange -- `dinomc_b200.GradAllReduce` (what DDP does at
main_dino_mc.py:260) driven by the real autograd Functions over the kernel double of tests/_ops_double.py.  Checked for every
transport / compression the GPU path offers: after `wait()` every parameter's `.grad` is the mean of the ranks' local gradients
(exactly for the fp32 exchange, within the bf16-GEMM tolerance for the bf16 exchanges), replicas are bit-identical, the
weight-normed last layer is exchanged ONCE (as an averaged dW when the peer transport claims it), nothing is left pending for
the next step, and the all-reduced center equals the reference's.  The hardware twin is tests/test_gpu_multi.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, compress, transport, norm_last_layer):
    import torch.distributed as dist
    for p in (HERE, os.path.dirname(HERE)):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import _ops_double as dbl
        import dinomc_b200 as D
        patch = dbl.Patcher()
        dbl.install(patch)
        dbl.install_data_parallel(patch)
        torch.manual_seed(0)                                   # identical weights on both ranks
        Din, K, B, C, G = 64, 1024, 4, 8, 2
        head = D.DINOHead(Din, K, hidden_dim=128, bottleneck_dim=64, norm_last_layer=norm_last_layer)
        teacher = D.DINOHead(Din, K, hidden_dim=128, bottleneck_dim=64, norm_last_layer=norm_last_layer)
        for p in teacher.parameters():
            p.requires_grad = False
        head.precision = teacher.precision = "bf16" if compress else "fp32"
        loss_mod = D.DINOLoss(K, C, 0.04, 0.04, 0, 10, teacher_crops_number=G)
        g = torch.Generator().manual_seed(100 + rank)          # different data per rank
        xs = torch.randn(C * B, Din, generator=g)
        xt = torch.randn(G * B, Din, generator=g)
        with torch.no_grad():
            t_out = teacher(xt)
        loss = loss_mod(head(xs), t_out, 0)                     # local (un-reduced) gradients first
        loss.backward()
        local = {n: p.grad.clone() for n, p in head.named_parameters() if p.grad is not None}
        center_after = loss_mod.center.clone()
        for p in head.parameters():
            p.grad = None
        red = D.GradAllReduce(head.parameters(), compress=compress, transport=transport)
        assert D.ops.track_ready and D.functional.grad_exchange is red
        reduced, n_exchanges = [], []
        for step in range(2):                                   # two steps: the second must not see leftovers of the first
            with torch.no_grad():
                loss_mod.center.zero_()
            dbl.calls.clear()
            loss2 = loss_mod(head(xs), t_out, 0)
            loss2.backward()
            red.wait()
            assert red._pending == [] and red._seen == 0 and red._late is None and not red._claimed and not D.ops.ready_events
            reduced.append({n: p.grad.clone() for n, p in head.named_parameters() if p.grad is not None})
            n_exchanges.append(dbl.calls.count("xrank_allreduce"))
            wn_bwd = dbl.calls.count("weightnorm_bwd")
            for p in head.parameters():
                p.grad = None
        red.remove()
        assert not D.ops.track_ready and D.functional.grad_exchange is None
        torch.save({"local": local, "reduced": reduced, "center": center_after, "t_out": t_out.float(), "loss": float(loss.detach()),
                    "xrank": n_exchanges, "wn_bwd": wn_bwd, "symm": dbl.SymmetricBuffer.instances},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("compress,transport,tol", [(None, "nccl", 1e-6), ("bf16", "nccl", 2e-2), ("bf16", "peer", 2e-2)])
@pytest.mark.parametrize("norm_last_layer", [True, False])
def test_gradient_exchange_host_logic_two_ranks(tmp_path, compress, transport, tol, norm_last_layer):
    import torch.multiprocessing as mp
    from oracle import np_oracle as O
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), compress, transport, norm_last_layer), nprocs=2, join=True)
    r = [torch.load(os.path.join(tmp_path, f"rank{i}.pt")) for i in range(2)]
    assert r[0]["loss"] != r[1]["loss"]                                        # different data per rank
    expected = {"mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "mlp.4.weight", "mlp.4.bias", "last_layer.weight_v"}
    if not norm_last_layer:
        expected.add("last_layer.weight_g")                                    # DINO-TP style trainable gain
    assert set(r[0]["local"]) == expected
    for step in range(2):
        assert set(r[0]["reduced"][step]) == expected
        for name in expected:
            mean = (r[0]["local"][name].double() + r[1]["local"][name].double()) / 2
            for i in range(2):
                got = r[i]["reduced"][step][name]
                assert got.dtype == torch.float32 and got.shape == r[i]["local"][name].shape
                err = (got.double() - mean).abs().max() / mean.abs().max()
                assert err < tol, (step, name, float(err))
            assert torch.equal(r[0]["reduced"][step][name], r[1]["reduced"][step][name])     # replicas stay bit-identical
        assert all(torch.equal(r[0]["reduced"][0][n], r[0]["reduced"][1][n]) for n in expected)   # same inputs, same result
    if transport == "peer":
        # the averaged-dW route: dW + two flushes of small gradients per step at most, symmetric buffers allocated once
        assert r[0]["xrank"][0] == r[0]["xrank"][1] and 2 <= r[0]["xrank"][0] <= 3
        assert r[0]["symm"] == r[0]["xrank"][0]
        assert r[0]["wn_bwd"] == 1                                            # ONE weight-norm backward, on the averaged dW
    else:
        assert r[0]["xrank"] == [0, 0] and r[0]["symm"] == 0
    outs = [x["t_out"].double().numpy() for x in r]
    ref = O.update_center(np.zeros((1, outs[0].shape[1])), outs[0], 0.9, world_size=2, all_rank_outputs=outs)
    for i in range(2):
        assert np.abs(r[i]["center"].double().numpy() - ref).max() / np.abs(ref).max() < 1e-6
    assert torch.equal(r[0]["center"], r[1]["center"])


def _ddp_worker(rank, world, port, out_dir):
    """The reference's own data-parallel wrapper (main_dino_mc.py:260: DistributedDataParallel around the student) over the drop-in
    head: DDP must see ordinary leaf parameters, receive every gradient through the custom autograd Functions, and average them."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    for p in (HERE, os.path.dirname(HERE)):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import _ops_double as dbl
        import dinomc_b200 as D
        patch = dbl.Patcher()
        dbl.install(patch)
        torch.manual_seed(0)
        Din, K, B, C, G = 64, 1024, 4, 8, 2
        head = D.DINOHead(Din, K, hidden_dim=128, bottleneck_dim=64)
        teacher = D.DINOHead(Din, K, hidden_dim=128, bottleneck_dim=64)
        for p in teacher.parameters():
            p.requires_grad = False
        head.precision = teacher.precision = "fp32"
        loss_mod = D.DINOLoss(K, C, 0.04, 0.04, 0, 10, teacher_crops_number=G)
        g = torch.Generator().manual_seed(100 + rank)
        xs = torch.randn(C * B, Din, generator=g)
        xt = torch.randn(G * B, Din, generator=g)

        def step(model):
            with torch.no_grad():
                loss_mod.center.zero_()
                t_out = teacher(xt)
            loss = loss_mod(model(xs), t_out, 0)
            loss.backward()
            return loss

        step(head)
        local = {n: p.grad.clone() for n, p in head.named_parameters() if p.grad is not None}
        for p in head.parameters():
            p.grad = None
        ddp = DDP(head)
        dbl.calls.clear()
        step(ddp)
        fused = "ce_fused" in dbl.calls                        # did the statistics attached by the head survive the DDP wrapper?
        reduced = {n: p.grad.clone() for n, p in head.named_parameters() if p.grad is not None}
        torch.save({"local": local, "reduced": reduced, "fused": fused}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_reference_ddp_wrapper_over_dropin_head_two_ranks(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [torch.load(os.path.join(tmp_path, f"rank{i}.pt")) for i in range(2)]
    assert set(r[0]["reduced"]) == set(r[0]["local"]) and len(r[0]["local"]) == 7
    for name in r[0]["local"]:
        mean = (r[0]["local"][name].double() + r[1]["local"][name].double()) / 2
        for i in range(2):
            err = (r[i]["reduced"][name].double() - mean).abs().max() / mean.abs().max()
            assert err < 1e-6, (name, float(err))
        assert torch.equal(r[0]["reduced"][name], r[1]["reduced"][name])
    # DDP hands the module's output tensor through unchanged (no find_unused_parameters): the statistics record the head attached
    # to its logits reaches DINOLoss and the single-pass loss route runs.  (Were a wrapper to re-wrap the tensor, DINOLoss would
    # find no record and take the separate-pass route -- slower, same numbers: test_host_wiring_cpu.)
    assert r[0]["fused"] is True and r[1]["fused"] is True
