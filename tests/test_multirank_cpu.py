"""CPU, gloo, world_size 2: semantics of the only data-path exchange the path has -- the SUM all-reduce of
the per-rank teacher column sums and the division by Nt*world (main_dino_mc.py:468-470) -- plus the
per-rank seeding / aggregation rules bench.py uses.  The CUDA kernels cannot run here; what is checked is
the host-side contract the GPU path implements (DINOLoss._update_center_from_colsum)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import np_oracle as O
from oracle import torch_port as T


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K, Nt, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(4321 + rank)              # bench.py's per-rank teacher seed
        t = torch.randn(Nt, K, generator=g)
        s = torch.randn(4 * Nt, K, generator=torch.Generator().manual_seed(1234 + rank))
        st = T.LossState(K, 8, 0.04, 0.04, 0, 10, teacher_crops_number=2)
        st.center = torch.full((1, K), 0.25)
        loss = T.loss_forward(st, s, t, 0, world_size=world, all_reduce=dist.all_reduce)
        # the exchange as the GPU path performs it: local column sum -> all_reduce(SUM) -> /(Nt*world) -> EMA
        colsum = t.sum(0)
        dist.all_reduce(colsum)
        bc = colsum.reshape(1, K) / (Nt * world)
        center_gpu_path = torch.full((1, K), 0.25) * 0.9 + bc * (1 - 0.9)
        q.put((rank, t.numpy(), st.center.numpy(), center_gpu_path.numpy(), float(loss)))
    finally:
        dist.destroy_process_group()


def test_center_allreduce_two_ranks():
    world, K, Nt = 2, 96, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, Nt, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs = [r[1] for r in res]
    ref = O.update_center(np.full((1, K), 0.25), outs[0], 0.9, world_size=world, all_rank_outputs=outs)
    for _, _, c_port, c_gpu_path, _ in res:
        assert np.abs(c_port - ref).max() < 1e-6                    # both ranks hold the same, correct center
        assert np.abs(c_port - c_gpu_path).max() < 1e-7             # split colsum / all-reduce / update: same result
    assert np.array_equal(res[0][2], res[1][2])                     # replicated state stays replicated
    assert not np.array_equal(res[0][1], res[1][1])                 # ranks really saw different data
    assert res[0][4] != res[1][4]                                   # and different local losses


def test_bench_weak_scaling_bookkeeping():
    import bench
    w = bench.WORKLOADS["cfg2"]
    shapes = bench.backbone_param_shapes(w["arch"])
    assert len(shapes) == 150 and sum(int(np.prod(s)) for s in shapes) == 21_670_272     # ViT-S/8
    head = w["D"] * 2048 + 2048 + 2048 * 2048 + 2048 + 2048 * 256 + 256 + w["K"] + w["K"] * 256
    P = sum(int(np.prod(s)) for s in shapes) + head
    assert abs(P / 1e6 - 44.02) < 0.01 and len(shapes) + 8 == 158
    flops, nbytes = bench.roofline_model(w, P, 4)
    assert abs(flops / 1e9 - 296.6) < 0.5 and abs(nbytes / 2 ** 30 - 4.630) < 0.01       # SURVEY 8d worked numbers
    flops, nbytes = bench.roofline_model(w, P, 2)
    assert abs(nbytes / 2 ** 30 - 2.880) < 0.01
    for arch, n_params in (("resnet50", 23.5), ("swin_t", 27.5), ("wide_resnet50_2", 66.8)):
        sh = bench.backbone_param_shapes(arch)
        assert abs(sum(int(np.prod(s)) for s in sh) / 1e6 - n_params) < 0.2, arch
