"""Diagnostic (torchrun): raw NCCL time of the step's gradient all-reduces, alone, and the step with/without them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True) if os.environ.get("HIPRI", "1") == "1" else None
dist.init_process_group("nccl", device_id=dev, pg_options=opts)
import dinomc_b200 as D
D.set_teacher_overlap(True); D.set_async_center(True)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# (b) raw NCCL: the seven gradient tensors of the head, back to back
shapes = [(2048, 384), (2048,), (2048, 2048), (2048,), (256, 2048), (256,), (65536, 256)]
grads = [torch.randn(s, device=dev) for s in shapes]
def comm_only():
    for g in reversed(grads):
        dist.all_reduce(g, op=dist.ReduceOp.AVG)
t_comm = timeit(comm_only)
big = grads[-1]
t_big = timeit(lambda: dist.all_reduce(big, op=dist.ReduceOp.AVG))
flat = torch.randn(sum(g.numel() for g in grads), device=dev)
t_flat = timeit(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG))

w = dict(bench.WORKLOADS["cfg2"])
step = bench.Step(w, "bf16", rank, world, dev, ddp=False, reserve_sms=int(os.environ.get("RESERVE", "16")))
g1 = D.StepGraph(step.run, warmup=3, capture_error_mode="thread_local")
t_with = timeit(g1.replay)
step.reducer.remove(); step.reducer = None
g2 = D.StepGraph(step.run, warmup=3, capture_error_mode="thread_local")
t_without = timeit(g2.replay)
if rank == 0:
    print(f"world={world}  raw all-reduce of 7 grads: {t_comm*1e3:.1f} us   64MB dv alone: {t_big*1e3:.1f} us   one flat 89MB: {t_flat*1e3:.1f} us")
    print(f"step with grad all-reduce: {t_with*1e3:.1f} us   without (center exchange only): {t_without*1e3:.1f} us")
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
