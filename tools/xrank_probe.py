"""2+ GPUs: libdinomc's peer-memory all-reduce (dmc_xrank_allreduce) against NCCL -- numbers and timing.
   torchrun --nproc-per-node N tools/xrank_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import dinomc_b200 as D
from dinomc_b200.xrank import SymmetricBuffer


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e3


for ctas in (148, 64, 32):
    for name, numel, dtype in (("dW bf16 32MB", 65536 * 256, torch.bfloat16), ("small bf16 12.6MB", 6295040, torch.bfloat16),
                               ("center f32 256KB", 65536, torch.float32), ("dv f32 64MB", 65536 * 256, torch.float32)):
        buf = SymmetricBuffer(numel, dtype, ctas=ctas)
        g = torch.Generator(device="cpu").manual_seed(100 + rank)
        x = (torch.randn(buf.numel, generator=g) * 0.01).to(dtype).to(dev)
        buf.tensor.copy_(x)
        ref = x.float().clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        ref /= world
        buf.allreduce_(1.0 / world)
        torch.cuda.synchronize()
        err = float((buf.tensor.float() - ref).abs().max()) / float(ref.abs().max())
        # every rank must hold identical bits
        chk = buf.tensor.view(torch.int16 if dtype == torch.bfloat16 else torch.int32).to(torch.int64).sum()
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        t_x = timeit(lambda: buf.allreduce_(1.0))
        y = x.clone()
        t_n = timeit(lambda: dist.all_reduce(y, op=dist.ReduceOp.AVG))
        if rank == 0:
            gb = buf.numel * x.element_size() / 1e9
            print(f"ctas={ctas:4d} {name:20s} multicast={buf.multicast} rel_err={err:.2e} identical={int(lo) == int(hi)} "
                  f"xrank {t_x:8.1f} us ({gb / (t_x * 1e-6):7.1f} GB/s alg)  nccl {t_n:8.1f} us", flush=True)
        del buf
# widening epilogue
buf = SymmetricBuffer(4096 + 8, torch.bfloat16)
a, b = torch.zeros(4096, device=dev), torch.zeros(5, device=dev)
buf.tensor.copy_(torch.arange(buf.numel, device=dev).float().mul(0.001 * (rank + 1)).bfloat16())
expect = sum(torch.arange(buf.numel, device=dev).float().mul(0.001 * (r + 1)).bfloat16().float() for r in range(world)) / world
buf.allreduce_(1.0 / world, widen_to=[a, b], widen_offsets=[0, 4096])
torch.cuda.synchronize()
ok = float(((a - expect[:4096]).abs() / expect[:4096].abs().clamp_min(1e-3)).max()) < 1e-2 and float(((b - expect[4096:4101]).abs() / expect[4096:4101]).max()) < 1e-2
if rank == 0:
    print("widen epilogue ok:", ok, flush=True)
# the exchange kernel inside a CUDA graph (what StepGraph does), with and without PDL
for pdl in (1, 0):
    D._lib.load().dmc_set_pdl(pdl)
    buf = SymmetricBuffer(1 << 20, torch.bfloat16)
    buf.tensor.fill_(1.0)
    side = torch.cuda.Stream()
    y = torch.ones(1 << 20, device=dev)
    try:
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            y.mul_(2.0)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                buf.allreduce_(1.0 / world)
            torch.cuda.current_stream().wait_stream(side)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        good = bool((buf.tensor.float() == 1.0).all())
        if rank == 0:
            print(f"graph capture + replay with PDL={pdl}: ok={good}", flush=True)
    except Exception as e:      # noqa: BLE001
        print(f"[rank {rank}] graph capture with PDL={pdl} FAILED: {type(e).__name__}: {str(e)[:200]}", flush=True)
        break
torch.cuda.synchronize(); dist.barrier()
os._exit(0)
