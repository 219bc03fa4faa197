import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dinomc_b200
ops = dinomc_b200.ops
M, N = 2048, int(os.environ.get("N", 65536))
for K in (64, 128, 256, 512, 1024, 2048):
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    for _ in range(3): ops.gemm(A, B, M, N, K, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm(A, B, M, N, K, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tiles = (M // 128) * (N // 256) / 148
    print(f"K={K:5d}: {ms*1e3:8.1f} us   per tile {ms*1e3/tiles:6.2f} us   per k-block {ms*1e3/tiles/(K/64):6.3f} us", flush=True)
