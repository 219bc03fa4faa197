import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
import dinomc_b200 as D
transport = sys.argv[1] if len(sys.argv) > 1 else "peer"
compress = "bf16" if transport == "peer" or (len(sys.argv) > 2 and sys.argv[2] == "bf16") else None
D.set_teacher_overlap(True)
D.set_async_center(True)
w = dict(bench.WORKLOADS[os.environ.get("DMC_PROF_WORKLOAD", "cfg2")])
step = bench.Step(w, "bf16", rank, world, dev, compress=compress, transport=transport)
if transport == "peer":
    from dinomc_b200.xrank import SymmetricBuffer
    step.loss_mod.center_exchange = SymmetricBuffer(w["K"], torch.float32, ctas=16)
for _ in range(5):
    step.run()
torch.cuda.synchronize()
g = D.StepGraph(step.run, warmup=3, capture_error_mode="thread_local")
for _ in range(5):
    g.replay()
torch.cuda.synchronize(); dist.barrier()
N = 6
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        g.replay()
    torch.cuda.synchronize()
if rank == 0:
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memset" not in e.name], key=lambda e: e.time_range.start)
    per = len(evs) // N
    last = evs[-per:]
    t0 = last[0].time_range.start
    print(f"## {world}-GPU step, transport={transport} compress={compress}: rank 0, last replay ({per} kernels): start_us dur_us stream name")
    for e in last:
        nm = e.name.replace("void ", "").replace("dmc::(anonymous namespace)::", "").split("(")[0][:70]
        print(f"  {e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  s{getattr(e, 'stream', '?')}  {nm}")
    print(f"  step span: {last[-1].time_range.end - t0:.1f} us")
torch.cuda.synchronize(); dist.barrier()
sys.stdout.flush()
os._exit(0)
