"""Warm per-kernel device times of one step, from the CUPTI activity records torch.profiler collects while the
step runs back to back (no replay, no cache flush).  Diagnostic only: bench.py never reports these numbers."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
w = dict(bench.WORKLOADS["cfg2"])
dev = torch.device("cuda", 0)
import dinomc_b200 as D
D.set_teacher_overlap(os.environ.get('DMC_BENCH_OVERLAP', '1') != '0')
force_dp = bool(int(os.environ.get("DMC_BENCH_FORCE_DP", "0")))     # 1-rank NCCL group: the reducer path without transfers
if force_dp:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29578")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    D.set_async_center(True)
step = bench.Step(w, mode, 0, 1, dev, force_reducer=force_dp)
for _ in range(5):
    step.run()
torch.cuda.synchronize()
use_graph = "--eager" not in sys.argv
run = D.StepGraph(step.run, warmup=3, capture_error_mode="thread_local" if force_dp else "global").replay if use_graph else step.run
for _ in range(3):
    run()
torch.cuda.synchronize()
N = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        run()
    torch.cuda.synchronize()
agg, cnt = collections.Counter(), collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace("void ", "").replace("dmc::(anonymous namespace)::", "")
        name = name.split("(")[0][:70]
        agg[name] += ev.device_time / N if hasattr(ev, "device_time") else ev.cuda_time / N
        cnt[name] += 1
# launch order of the LAST profiled step (start time, duration, gap to the previous kernel's end on any stream)
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name],
             key=lambda e: e.time_range.start)
per_step = len(evs) // N
last = evs[-per_step:]
t0 = last[0].time_range.start
print(f"## launch order, last step ({per_step} kernels): start_us dur_us name")
for e in last:
    nm = e.name.replace("void ", "").replace("dmc::(anonymous namespace)::", "").split("(")[0][:60]
    print(f"  {e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  s{getattr(e, 'stream', getattr(e, 'device_resource_id', '?'))}  {nm}")
print(f"  step span: {last[-1].time_range.end - t0:.1f} us")
tot = sum(agg.values())
print(f"sum of kernel time per step: {tot:.1f} us")
for n, v in agg.most_common(40):
    print(f"{n:72s} x{cnt[n] / N:<4.1f} {v:8.1f} us {100 * v / tot:5.1f}%")

if force_dp:
    sys.stdout.flush()
    os._exit(0)
