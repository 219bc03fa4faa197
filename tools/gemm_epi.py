"""Graph-captured timing (no host in the loop) of dmc_gemm on the MLP shapes of the step, per epilogue variant.
Diagnostic only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dinomc_b200
from dinomc_b200 import _lib as L
ops = dinomc_b200.ops

CASES = [  # name, M, N, K, a_mn, b_mn, out dtype
    ("t_fwd1", 512, 2048, 384, False, False, torch.bfloat16),
    ("t_fwd2", 512, 2048, 2048, False, False, torch.bfloat16),
    ("s_fwd1", 2048, 2048, 384, False, False, torch.bfloat16),
    ("s_fwd2", 2048, 2048, 2048, False, False, torch.bfloat16),
    ("s_fwd3", 2048, 256, 2048, False, False, torch.float32),
    ("dgrad2", 2048, 2048, 2048, False, True, torch.bfloat16),
    ("dgrad1", 2048, 2048, 256, False, True, torch.bfloat16),
    ("wgrad2", 2048, 2048, 2048, True, True, torch.float32),
]
VARIANTS = sys.argv[1].split(",") if len(sys.argv) > 1 else ["plain", "bias", "gelu", "gelu_aux", "gelu_bwd"]
REPS = 20


def timed(fn):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(REPS):
                fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / REPS)
    return best * 1e3


for name, M, N, K, a_mn, b_mn, odt in CASES:
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
    B = (torch.randn((K, N) if b_mn else (N, K), device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(M, N, dtype=odt, device="cuda")
    bias = torch.randn(N, device="cuda")
    aux = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    res = []
    for v in VARIANTS:
        kw = {}
        if v == "bias": kw = dict(bias=bias)
        elif v == "gelu": kw = dict(bias=bias, act=L.ACT_GELU)
        elif v == "gelu_aux": kw = dict(bias=bias, act=L.ACT_GELU, aux=aux)
        elif v == "gelu_bwd": kw = dict(act=L.ACT_GELU_BWD, aux=aux)
        us = timed(lambda: ops.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out=out, **kw))
        res.append(f"{v}: {us:6.1f}")
    Am = A.t() if a_mn else A
    Bm = B if b_mn else B.t()
    lib = timed(lambda: torch.matmul(Am, Bm))
    print(f"{name:8s} M={M:5d} N={N:5d} K={K:5d}  " + "  ".join(res) + f"   cublas {lib:6.1f} us", flush=True)
