"""Pure-write, pure-read and copy HBM bandwidth (CUDA events, graph-captured loops).  Context for the rooflines: the
last-layer forward is a 268 MB pure write, the loss pass a read-heavy mix."""
import torch
def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / reps)
    return best * 1e-3
for mb in (268, 604, 1024):
    n = mb * 1000 * 1000 // 2
    a = torch.empty(n, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
    tw = timed(lambda: a.fill_(1.0))
    tr = timed(lambda: a.sum())        # read-only reduction
    tc = timed(lambda: b.copy_(a))
    print(f"{mb:5d} MB  write {mb/1e3/tw:7.1f} GB/s ({tw*1e6:6.1f} us)   read {mb/1e3/tr:7.1f} GB/s ({tr*1e6:6.1f} us)   copy r+w {2*mb/1e3/tc:7.1f} GB/s ({tc*1e6:6.1f} us)", flush=True)
