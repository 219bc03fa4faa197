"""Stand-alone timing of the loss kernels at cfg2 size (bf16 logits).  Not a test."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dinomc_b200
ops = dinomc_b200.ops
B, C, G, K = 256, 8, 2, 65536
dt = torch.bfloat16
s = (torch.randn(C * B, K, device="cuda") * 0.3).to(dt)
t = (torch.randn(G * B, K, device="cuda") * 0.3).to(dt)
c = torch.randn(K, device="cuda") * 0.1
gout = torch.ones((), device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
stats, _ = ops.teacher_stats_colsum(t, c, 25.0)
loss, lse = ops.ce_fwd(s, t, c, stats, B, C, G, 10.0, 25.0)
MB = lambda x: x / 1e6
e = 2
print("teacher_pass  %7.1f us  (%.0f MB)" % (timeit(lambda: ops.teacher_stats_colsum(t, c, 25.0)), MB(e * K * G * B)))
print("ce_fwd        %7.1f us  (%.0f MB)" % (timeit(lambda: ops.ce_fwd(s, t, c, stats, B, C, G, 10.0, 25.0)), MB(e * K * (C + G) * B)))
print("ce_bwd        %7.1f us  (%.0f MB)" % (timeit(lambda: ops.ce_bwd(s, t, c, stats, lse, gout, B, C, G, 10.0, 25.0)), MB(e * K * (2 * C + G) * B)))
print("ce_fused      %7.1f us  (%.0f MB)" % (timeit(lambda: ops.ce_fused(s, t, c, stats, lse, B, C, G, 10.0, 25.0)), MB(e * K * (2 * C + G) * B)))
